// Fused feature-propagation block for sm_100a: three_nn weights -> three_interpolate -> concat skip ->
// (conv1x1 -> GroupNorm(4) -> ReLU) x L over the n points of the finer level.
//
// Replaces the torch-level stack of the reference's PointnetFPModule (utils/pointnet2_util.py:91-120 with
// SharedMLP utils/nn_util.py:151-168): 1/(dist+1e-8), sum, divide, three_interpolate, cat, and per layer a
// conv1x1 (cuDNN/cuBLAS), native_group_norm and ReLU -- ~14 launches forward and ~3x that backward per FP level,
// each a pass over (B,C,n) tensors.
//
// Here:  fp_interp_concat_kernel   builds the layer-0 operand X = [interp(known_feats) ; skip] and the weights,
//        pw_fwd_kernel             one kernel per layer: Y_l = W_l act(Y_{l-1}) with GroupNorm+ReLU of the previous
//                                  layer applied in the operand loader, K (= C_in up to 448) streamed through
//                                  shared memory in chunks, GroupNorm sums (fp64) in the epilogue,
//        gn_relu_apply_kernel      the block's output relu(scale*y+shift),
//        gn_relu_bwd_stats_kernel  backward entry: dz_L = relu'(.) dOut and its GroupNorm-backward sums.
// The layer backward reuses the SIMT set-abstraction kernels of mlp_bwd.cu (dense dz, nsample = 1): mlp_dw_kernel
// (identity activation for layer 0) and mlp_dx_kernel (dense mode between layers, plain mode for the gradient
// of X).  fp32 FMA throughout (bit-level fp32 conv semantics; the FP stack is 5 % of the network's FLOPs).
#include "mlp_common.cuh"

namespace ogc {

constexpr int kFpThreads = 256;
constexpr int kFpChanChunk = 16;

// X (B, c2+c1, n): channels [0,c2) = three_interpolate(known_feats, idx, w), [c2,c2+c1) = skip.
// w_j = (1/(dist_j+1e-8)) / sum_j(1/(dist_j+1e-8)),  dist = sqrt(d2)   (utils/pointnet2_util.py:98-101)
__global__ void __launch_bounds__(kFpThreads)
fp_interp_concat_kernel(int c2, int m, int c1, int n, const float *__restrict__ known_feats, const int *__restrict__ idx,
                        const float *__restrict__ d2, const float *__restrict__ skip, float *__restrict__ X,
                        float *__restrict__ weight) {
    const int b = blockIdx.z;
    const int ch0 = blockIdx.y * kFpChanChunk;
    const int p = blockIdx.x * kFpThreads + threadIdx.x;
    if (p >= n) return;
    const int ctot = c2 + c1;
    const int ch1 = min(ctot, ch0 + kFpChanChunk);
    if (ch0 < c2) {
        const size_t o3 = (static_cast<size_t>(b) * n + p) * 3;
        const int i0 = __ldg(idx + o3), i1 = __ldg(idx + o3 + 1), i2 = __ldg(idx + o3 + 2);
        const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2 + o3)), 1e-8f));
        const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2 + o3 + 1)), 1e-8f));
        const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2 + o3 + 2)), 1e-8f));
        const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
        const float w0 = __fdiv_rn(r0, norm), w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm);
        if (blockIdx.y == 0) { weight[o3] = w0; weight[o3 + 1] = w1; weight[o3 + 2] = w2; }
        const int ce = min(ch1, c2);
        for (int c = ch0; c < ce; ++c) {
            const float *src = known_feats + (static_cast<size_t>(b) * c2 + c) * m;
            X[(static_cast<size_t>(b) * ctot + c) * n + p] =
                __fmaf_rn(w2, __ldg(src + i2), __fmaf_rn(w0, __ldg(src + i0), __fmul_rn(w1, __ldg(src + i1))));
        }
    }
    for (int c = max(ch0, c2); c < ch1; ++c)
        X[(static_cast<size_t>(b) * ctot + c) * n + p] = __ldg(skip + (static_cast<size_t>(b) * c1 + (c - c2)) * n + p);
}

struct PwFwdParams {
    int Cin, Cout, P;
    const float *x;    // (B,Cin,P): raw layer input (ss == NULL) or the previous layer's pre-norm output
    const float *ss;   // (B,Cin,2) GroupNorm scale/shift of the previous layer, or NULL (identity, no ReLU)
    const float *Wt;   // (Cin,Cout)
    float *y;          // (B,Cout,P)
    double *sums;      // (B,4,2)
};

// Bs[k][p] = act(x[b, kc+k, p_base+p]) for k < kn, p < P_T; zero beyond P.
template <int P_T>
__device__ __forceinline__ void pw_load_chunk(const PwFwdParams &q, int b, int kc, int kn, int p_base, float *Bs, int ldb) {
    constexpr int Q4 = P_T / 4;
    const int total = kn * Q4;
    const bool act = q.ss != nullptr;
    for (int e0 = threadIdx.x; e0 < total; e0 += kMlpThreads * 4) {
        float4 v[4];
        float sc[4], sh[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * kMlpThreads;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            sc[u] = 1.f; sh[u] = 0.f;
            if (e < total) {
                const int k = e / Q4, p = (e - k * Q4) * 4, c = kc + k;
                if (act) {
                    sc[u] = __ldg(q.ss + (static_cast<size_t>(b) * q.Cin + c) * 2);
                    sh[u] = __ldg(q.ss + (static_cast<size_t>(b) * q.Cin + c) * 2 + 1);
                }
                const float *src = q.x + (static_cast<size_t>(b) * q.Cin + c) * q.P + p_base + p;
                if (p_base + p + 3 < q.P && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
                    v[u] = __ldg(reinterpret_cast<const float4 *>(src));
                } else {
                    v[u].x = p_base + p + 0 < q.P ? __ldg(src + 0) : 0.f;
                    v[u].y = p_base + p + 1 < q.P ? __ldg(src + 1) : 0.f;
                    v[u].z = p_base + p + 2 < q.P ? __ldg(src + 2) : 0.f;
                    v[u].w = p_base + p + 3 < q.P ? __ldg(src + 3) : 0.f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * kMlpThreads;
            if (e >= total) continue;
            const int k = e / Q4, p = (e - k * Q4) * 4;
            float4 o = v[u];
            if (act) {
                o.x = fmaxf(fmaf(sc[u], o.x, sh[u]), 0.f);
                o.y = fmaxf(fmaf(sc[u], o.y, sh[u]), 0.f);
                o.z = fmaxf(fmaf(sc[u], o.z, sh[u]), 0.f);
                o.w = fmaxf(fmaf(sc[u], o.w, sh[u]), 0.f);
            }
            if (p_base + p + 3 >= q.P) {            // tail of the sample: padded positions contribute nothing
                if (p_base + p + 0 >= q.P) o.x = 0.f;
                if (p_base + p + 1 >= q.P) o.y = 0.f;
                if (p_base + p + 2 >= q.P) o.z = 0.f;
                o.w = 0.f;
            }
            *reinterpret_cast<float4 *>(Bs + k * ldb + p) = o;
        }
    }
}

template <int R_T, int P_T>
__global__ void __launch_bounds__(kMlpThreads, 2)
pw_fwd_kernel(PwFwdParams q) {
    constexpr int TX = P_T / 8, LDB = P_T + 4, KC = 32;
    extern __shared__ __align__(16) float smem[];
    __shared__ double gs[kGnGroups][2];
    float *As = smem;               // [KC][R_T]
    float *Bs = smem + KC * R_T;    // [KC][LDB]
    const int tid = threadIdx.x, ty = tid / TX, tx = tid % TX;
    const int b = blockIdx.y;
    const int Cout = q.Cout, cpg = Cout / kGnGroups, P = q.P;
    if (tid < kGnGroups * 2) (&gs[0][0])[tid] = 0.0;
    const int rows[2] = {ty * 4, R_T / 2 + ty * 4};
    const int cols[2] = {tx * 4, P_T / 2 + tx * 4};
    const int ntiles = (P + P_T - 1) / P_T;
    float ps[2] = {0.f, 0.f}, pq[2] = {0.f, 0.f};

    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int p_base = t * P_T;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int kc = 0; kc < q.Cin; kc += KC) {
            const int kn = min(KC, q.Cin - kc);
            __syncthreads();
            for (int e = tid; e < kn * R_T; e += kMlpThreads) {
                const int k = e / R_T, r = e - k * R_T;
                As[e] = r < Cout ? __ldg(q.Wt + static_cast<size_t>(kc + k) * Cout + r) : 0.f;
            }
            pw_load_chunk<P_T>(q, b, kc, kn, p_base, Bs, LDB);
            __syncthreads();
            tile_gemm<R_T, P_T>(As, Bs, LDB, kn, ty, tx, acc);
        }
#pragma unroll
        for (int rc = 0; rc < 2; ++rc) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = rows[rc] + i;
                if (r >= Cout) continue;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int p = p_base + cols[cc];
                    const float v[4] = {acc[rc * 4 + i][cc * 4 + 0], acc[rc * 4 + i][cc * 4 + 1],
                                        acc[rc * 4 + i][cc * 4 + 2], acc[rc * 4 + i][cc * 4 + 3]};
                    float *dst = q.y + (static_cast<size_t>(b) * Cout + r) * P + p;
                    if (p + 3 < P) {
                        ps[rc] += (v[0] + v[1]) + (v[2] + v[3]);
                        pq[rc] += (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
                        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                        else { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; dst[3] = v[3]; }
                    } else {
                        for (int j = 0; j < 4; ++j)
                            if (p + j < P) {
                                ps[rc] += v[j];
                                pq[rc] += v[j] * v[j];
                                dst[j] = v[j];
                            }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int rc = 0; rc < 2; ++rc) {
        float s = ps[rc], sq = pq[rc];
        constexpr int W = TX >= 32 ? 32 : TX;
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) {
            s += __shfl_xor_sync(OGC_FULL_MASK, s, o);
            sq += __shfl_xor_sync(OGC_FULL_MASK, sq, o);
        }
        if ((tid % W) == 0 && rows[rc] < Cout) {
            const int g = rows[rc] / cpg;
            atomicAdd(&gs[g][0], static_cast<double>(s));
            atomicAdd(&gs[g][1], static_cast<double>(sq));
        }
    }
    __syncthreads();
    if (tid < kGnGroups * 2) atomicAdd(q.sums + static_cast<size_t>(b) * kGnGroups * 2 + tid, (&gs[0][0])[tid]);
}

// out = relu(scale*y + shift), rows of P values, 4 per thread.
__global__ void __launch_bounds__(256)
gn_relu_apply_kernel(long long rows, int P, const float *__restrict__ y, const float *__restrict__ ss, float *__restrict__ out) {
    const int q4 = (P + 3) / 4;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows * q4) return;
    const long long r = i / q4;
    const int p = static_cast<int>(i - r * q4) * 4;
    const float sc = __ldg(ss + r * 2), sh = __ldg(ss + r * 2 + 1);
    const float *src = y + r * P + p;
    float *dst = out + r * P + p;
    if (p + 3 < P && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src));
        *reinterpret_cast<float4 *>(dst) = make_float4(fmaxf(fmaf(sc, v.x, sh), 0.f), fmaxf(fmaf(sc, v.y, sh), 0.f),
                                                       fmaxf(fmaf(sc, v.z, sh), 0.f), fmaxf(fmaf(sc, v.w, sh), 0.f));
    } else {
        for (int j = 0; j < 4 && p + j < P; ++j) dst[j] = fmaxf(fmaf(sc, __ldg(src + j), sh), 0.f);
    }
}

// Backward entry of a (GroupNorm, ReLU) output: dz = (scale*y+shift > 0) ? dout : 0 and the GroupNorm-backward
// sums  ab (B,4,2) += [sum gamma dz, sum gamma dz yhat], dgamma[c] += sum dz yhat, dbeta[c] += sum dz.
// One CTA per (channel, sample) row.
__global__ void __launch_bounds__(256)
gn_relu_bwd_stats_kernel(int C, int P, const float *__restrict__ dout, const float *__restrict__ y,
                         const float *__restrict__ ss, const float *__restrict__ mean_rstd,
                         const float *__restrict__ gamma, float *__restrict__ dz, double *__restrict__ ab,
                         float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int c = blockIdx.x, b = blockIdx.y, g = c / (C / kGnGroups);
    const size_t row = (static_cast<size_t>(b) * C + c);
    const float sc = __ldg(ss + row * 2), sh = __ldg(ss + row * 2 + 1);
    const float mean = mean_rstd[(b * kGnGroups + g) * 2], rstd = mean_rstd[(b * kGnGroups + g) * 2 + 1];
    const float *dp = dout + row * P, *yp = y + row * P;
    float *zp = dz + row * P;
    float s = 0.f, sy = 0.f;
    const bool vec = (P % 4 == 0) && ((reinterpret_cast<uintptr_t>(dp) | reinterpret_cast<uintptr_t>(yp) | reinterpret_cast<uintptr_t>(zp)) & 15u) == 0;
    if (vec) {
        for (int p = threadIdx.x * 4; p < P; p += blockDim.x * 4) {
            const float4 d = __ldg(reinterpret_cast<const float4 *>(dp + p));
            const float4 v = __ldg(reinterpret_cast<const float4 *>(yp + p));
            float4 o;
            o.x = fmaf(sc, v.x, sh) > 0.f ? d.x : 0.f;
            o.y = fmaf(sc, v.y, sh) > 0.f ? d.y : 0.f;
            o.z = fmaf(sc, v.z, sh) > 0.f ? d.z : 0.f;
            o.w = fmaf(sc, v.w, sh) > 0.f ? d.w : 0.f;
            *reinterpret_cast<float4 *>(zp + p) = o;
            s += (o.x + o.y) + (o.z + o.w);
            sy += o.x * ((v.x - mean) * rstd) + o.y * ((v.y - mean) * rstd) + o.z * ((v.z - mean) * rstd) + o.w * ((v.w - mean) * rstd);
        }
    } else {
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            const float v = __ldg(yp + p);
            const float o = fmaf(sc, v, sh) > 0.f ? __ldg(dp + p) : 0.f;
            zp[p] = o;
            s += o;
            sy += o * ((v - mean) * rstd);
        }
    }
    __shared__ float red[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(OGC_FULL_MASK, s, o);
        sy += __shfl_xor_sync(OGC_FULL_MASK, sy, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = sy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        s = sy = 0.f;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) { s += red[0][w]; sy += red[1][w]; }
        atomicAdd(dbeta + c, s);
        atomicAdd(dgamma + c, sy);
        atomicAdd(ab + (b * kGnGroups + g) * 2, static_cast<double>(gamma[c]) * s);
        atomicAdd(ab + (b * kGnGroups + g) * 2 + 1, static_cast<double>(gamma[c]) * sy);
    }
}

template <int R_T, int P_T>
static cudaError_t launch_pw_fwd(const PwFwdParams &q, int B, cudaStream_t st) {
    constexpr int KC = 32;
    const size_t smem = (static_cast<size_t>(KC) * R_T + static_cast<size_t>(KC) * (P_T + 4)) * sizeof(float);
    const int ntiles = (q.P + P_T - 1) / P_T;
    int per_sample = (kNumSMs * 2) / B;   // floor: the whole grid must be resident at 2 CTAs/SM (a 297th CTA would run as a second wave)
    per_sample = per_sample > ntiles ? ntiles : (per_sample < 1 ? 1 : per_sample);
    dim3 grid(per_sample, B);
    cudaError_t e = cudaFuncSetAttribute(pw_fwd_kernel<R_T, P_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    pw_fwd_kernel<R_T, P_T><<<grid, kMlpThreads, smem, st>>>(q);
    return cudaGetLastError();
}

}  // namespace ogc

extern "C" int ogc_fp_interp_concat(int b, int c2, int m, int c1, int n, const float *known_feats, const int *idx,
                                    const float *dist2, const float *skip, float *x, float *weight, void *stream) {
    using namespace ogc;
    if (b < 0 || c2 <= 0 || m <= 0 || c1 < 0 || n < 0) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n == 0) return OGC_OK;
    if (!known_feats || !idx || !dist2 || !x || !weight || (c1 > 0 && !skip)) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((n + kFpThreads - 1) / kFpThreads, (c2 + c1 + kFpChanChunk - 1) / kFpChanChunk, b);
    fp_interp_concat_kernel<<<grid, kFpThreads, 0, static_cast<cudaStream_t>(stream)>>>(c2, m, c1, n, known_feats, idx,
                                                                                       dist2, skip, x, weight);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_pw_mlp_layer_fwd(int b, int p, int cin, int cout, const float *x, const float *ss_prev,
                                    const float *wt, float *y, double *sums, void *stream) {
    using namespace ogc;
    if (b < 0 || p <= 0 || cin <= 0 || cout <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!x || !wt || !y || !sums) return OGC_ERR_INVALID_ARG;
    if (cout % 16 != 0 || cout > 256 || b > 65535) return OGC_ERR_UNSUPPORTED;
    PwFwdParams q;
    q.Cin = cin; q.Cout = cout; q.P = p; q.x = x; q.ss = ss_prev; q.Wt = wt; q.y = y; q.sums = sums;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (cout <= 32) e = launch_pw_fwd<32, 512>(q, b, st);
    else if (cout <= 64) e = launch_pw_fwd<64, 256>(q, b, st);
    else if (cout <= 128) e = launch_pw_fwd<128, 128>(q, b, st);
    else e = launch_pw_fwd<256, 64>(q, b, st);
    return e == cudaSuccess ? OGC_OK : static_cast<int>(e);
}

extern "C" int ogc_gn_relu_apply(int b, int c, int p, const float *y, const float *scale_shift, float *out,
                                 void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || p <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!y || !scale_shift || !out) return OGC_ERR_INVALID_ARG;
    const long long rows = static_cast<long long>(b) * c;
    const long long work = rows * ((p + 3) / 4);
    gn_relu_apply_kernel<<<static_cast<unsigned>((work + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rows, p, y, scale_shift, out);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_gn_relu_bwd_stats(int b, int c, int p, const float *dout, const float *y, const float *scale_shift,
                                     const float *mean_rstd, const float *gamma, float *dz, double *ab, float *dgamma,
                                     float *dbeta, void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || c % kGnGroups != 0 || p <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!dout || !y || !scale_shift || !mean_rstd || !gamma || !dz || !ab || !dgamma || !dbeta) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid(c, b);
    gn_relu_bwd_stats_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(c, p, dout, y, scale_shift, mean_rstd,
                                                                                 gamma, dz, ab, dgamma, dbeta);
    OGC_RETURN_LAUNCH_STATUS();
}
