// Fused set-abstraction MLP for sm_100a: group -> (conv1x1 -> GroupNorm(4) -> ReLU) x L -> max over nsample.
//
// Replaces the torch-level stack of the reference's SA module (utils/pointnet2_util.py:33-44 with
// QueryAndGroup pointnet2/pointnet2.py:283-294 and SharedMLP utils/nn_util.py:151-168):
// grouping_operation x2 + concat materialise a (B,3+C,M,S) tensor, every conv / GroupNorm / ReLU /
// max_pool2d is a separate pass over (B,C,M,S) tensors through HBM (~40 GB per training step at
// KITTI-SF sizes, measured 60 of 82 ms in profiles/r01_step_torchprofiler_v1_unfused_mlp.txt).
//
// Here every layer is ONE kernel: the operand tile is produced on the fly -- gathered from the point
// cloud / point-major feature rows through the neighbour indices (layer 1: the grouped tensor never
// exists), or read from the previous layer's pre-norm output with GroupNorm+ReLU applied in the
// loader -- multiplied by the weight tile held in shared memory (fp32 FMA, 8x8 register blocking, no
// TF32: bit-level behaviour of an fp32 conv), and the epilogue accumulates the GroupNorm statistics
// of the layer (fp64) and, for the last layer, the max / min over the nsample axis with their
// positions.  GroupNorm needs sample-wide statistics before the next layer can start, hence one
// kernel per layer; only the pre-norm outputs y_l round-trip through HBM (once).
//
// Backward: per layer one kernel for the input gradient (dX) and one for the weight gradient (dW);
// the GroupNorm+ReLU backward is evaluated on the fly in their loaders from (dz_l, y_l, group sums).
//
//   y_l = W_l a_{l-1},  yhat = (y - mu_g) rstd_g,  z = gamma yhat + beta,  a_l = relu(z)
//   dyhat = gamma dz ;  dy = rstd_g (dyhat - mean_g(dyhat) - yhat mean_g(dyhat yhat))
#include "mlp_common.cuh"

namespace ogc {


// Operand tile loaders: Bs[c][p] for c < Cin, p < P_T (row stride ldb), zero beyond P.
__device__ __forceinline__ void load_tile_gather(const MlpFwdParams &q, int b, int p_base, int P_T, float *Bs, int ldb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = kMlpThreads >> 5;
    const int Cf = q.Cf;
    if (Cf <= 8) {                                  // narrow rows (SA1: raw coordinates): one thread per position
        for (int p = threadIdx.x; p < P_T; p += kMlpThreads) {
            const int gp = p_base + p;
            if (gp < q.P) {
                const int j = __ldg(q.idx + static_cast<size_t>(b) * q.P + gp);
                const int m = gp / q.S;
                const float *pj = q.xyz + (static_cast<size_t>(b) * q.N + j) * 3;
                const float *pc = q.new_xyz + (static_cast<size_t>(b) * q.M + m) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) Bs[c * ldb + p] = __ldg(pj + c) - __ldg(pc + c);
                for (int c = 0; c < Cf; ++c) Bs[(3 + c) * ldb + p] = __ldg(q.feat_pm + (static_cast<size_t>(b) * q.N + j) * Cf + c);
            } else {
                for (int c = 0; c < q.Cin; ++c) Bs[c * ldb + p] = 0.f;
            }
        }
        return;
    }
    for (int p = warp; p < P_T; p += nwarp) {      // one warp per position: coalesced feature-row reads
        const int gp = p_base + p;
        if (gp < q.P) {
            const int j = __ldg(q.idx + static_cast<size_t>(b) * q.P + gp);
            const int m = gp / q.S;
            if (lane < 3)
                Bs[lane * ldb + p] = __ldg(q.xyz + (static_cast<size_t>(b) * q.N + j) * 3 + lane) -
                                     __ldg(q.new_xyz + (static_cast<size_t>(b) * q.M + m) * 3 + lane);
            const float *row = q.feat_pm + (static_cast<size_t>(b) * q.N + j) * Cf;
            for (int c = lane; c < Cf; c += 32) Bs[(3 + c) * ldb + p] = __ldg(row + c);
        } else {
            for (int c = lane; c < q.Cin; c += 32) Bs[c * ldb + p] = 0.f;
        }
    }
}

__device__ __forceinline__ void load_tile_dense(const float *__restrict__ y_prev, const float *__restrict__ ss, int Cin,
                                                int P, int b, int p_base, int P_T, float *Bs, int ldb) {
    const int q4 = P_T / 4;
    const int total = Cin * q4;
    // 4 independent 16-byte loads in flight per thread before any is consumed (latency-bound loader)
    for (int e0 = threadIdx.x; e0 < total; e0 += kMlpThreads * 4) {
        float4 v[4];
        float sc[4], sh[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * kMlpThreads;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            sc[u] = sh[u] = 0.f;
            if (e < total) {
                const int c = e / q4, p = (e - c * q4) * 4;
                sc[u] = __ldg(ss + (static_cast<size_t>(b) * Cin + c) * 2);
                sh[u] = __ldg(ss + (static_cast<size_t>(b) * Cin + c) * 2 + 1);
                const float *src = y_prev + (static_cast<size_t>(b) * Cin + c) * P + p_base + p;
                if (p_base + p + 3 < P && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
                    v[u] = __ldg(reinterpret_cast<const float4 *>(src));
                } else {
                    v[u].x = p_base + p + 0 < P ? __ldg(src + 0) : 0.f;
                    v[u].y = p_base + p + 1 < P ? __ldg(src + 1) : 0.f;
                    v[u].z = p_base + p + 2 < P ? __ldg(src + 2) : 0.f;
                    v[u].w = p_base + p + 3 < P ? __ldg(src + 3) : 0.f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * kMlpThreads;
            if (e >= total) continue;
            const int c = e / q4, p = (e - c * q4) * 4;
            float4 o;
            o.x = p_base + p + 0 < P ? fmaxf(fmaf(sc[u], v[u].x, sh[u]), 0.f) : 0.f;
            o.y = p_base + p + 1 < P ? fmaxf(fmaf(sc[u], v[u].y, sh[u]), 0.f) : 0.f;
            o.z = p_base + p + 2 < P ? fmaxf(fmaf(sc[u], v[u].z, sh[u]), 0.f) : 0.f;
            o.w = p_base + p + 3 < P ? fmaxf(fmaf(sc[u], v[u].w, sh[u]), 0.f) : 0.f;
            *reinterpret_cast<float4 *>(Bs + c * ldb + p) = o;
        }
    }
}

// ------------------------------------------------------------------------------------------ forward
template <int R_T, int P_T, bool GATHER, bool LAST>
__global__ void __launch_bounds__(kMlpThreads, 2)
mlp_fwd_kernel(MlpFwdParams q) {
    constexpr int TX = P_T / 8;
    constexpr int LDB = P_T + 4;
    extern __shared__ __align__(16) float smem[];
    __shared__ double gs[kGnGroups][2];
    float *As = smem;                      // [Cin][R_T]
    float *Bs = smem + q.Cin * R_T;        // [Cin][LDB]
    const int tid = threadIdx.x, ty = tid / TX, tx = tid % TX;
    const int b = blockIdx.y;
    const int Cout = q.Cout, cpg = Cout / kGnGroups;

    for (int e = tid; e < q.Cin * R_T; e += kMlpThreads) {
        const int k = e / R_T, r = e - k * R_T;
        As[e] = r < Cout ? __ldg(q.Wt + static_cast<size_t>(k) * Cout + r) : 0.f;
    }
    if (tid < kGnGroups * 2) (&gs[0][0])[tid] = 0.0;

    const int rows[2] = {ty * 4, R_T / 2 + ty * 4};
    const int cols[2] = {tx * 4, P_T / 2 + tx * 4};
    const int ntiles = (q.P + P_T - 1) / P_T;
    float ps[2] = {0.f, 0.f}, pq[2] = {0.f, 0.f};   // per-thread partial group sums (row chunk 0 / 1)

    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int p_base = t * P_T;
        __syncthreads();
        if (GATHER) load_tile_gather(q, b, p_base, P_T, Bs, LDB);
        else load_tile_dense(q.y_prev, q.ss_prev, q.Cin, q.P, b, p_base, P_T, Bs, LDB);
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        tile_gemm<R_T, P_T>(As, Bs, LDB, q.Cin, ty, tx, acc);

        // ---- epilogue: store y, GroupNorm statistics, (LAST) max/min over the nsample axis ----
#pragma unroll
        for (int rc = 0; rc < 2; ++rc) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = rows[rc] + i;
                if (r >= Cout) continue;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int p = p_base + cols[cc];
                    const float v0 = acc[rc * 4 + i][cc * 4 + 0], v1 = acc[rc * 4 + i][cc * 4 + 1],
                                v2 = acc[rc * 4 + i][cc * 4 + 2], v3 = acc[rc * 4 + i][cc * 4 + 3];
                    if (p + 3 < q.P) {
                        ps[rc] += (v0 + v1) + (v2 + v3);
                        pq[rc] += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
                        if (q.y) {
                            float *dst = q.y + (static_cast<size_t>(b) * Cout + r) * q.P + p;
                            if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) *reinterpret_cast<float4 *>(dst) = make_float4(v0, v1, v2, v3);
                            else { dst[0] = v0; dst[1] = v1; dst[2] = v2; dst[3] = v3; }
                        }
                    } else {
                        const float v[4] = {v0, v1, v2, v3};
                        for (int j = 0; j < 4; ++j)
                            if (p + j < q.P) {
                                ps[rc] += v[j];
                                pq[rc] += v[j] * v[j];
                                if (q.y) q.y[(static_cast<size_t>(b) * Cout + r) * q.P + p + j] = v[j];
                            }
                    }
                }
            }
        }
        if (LAST) {
            // S == 64: the 64 samples of one centre are 16 consecutive threads x 4 positions of a column chunk
            // (P_T >= 128), or both column chunks of 8 threads (P_T == 64).
#pragma unroll
            for (int rc = 0; rc < 2; ++rc)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = rows[rc] + i;
                    float vmx[2], vmn[2];
                    int imx[2], imn[2];
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int s0 = cols[cc] & 63;
                        vmx[cc] = -INFINITY; vmn[cc] = INFINITY; imx[cc] = 0; imn[cc] = 0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float v = acc[rc * 4 + i][cc * 4 + j];
                            const bool ok = p_base + cols[cc] + j < q.P;
                            if (ok && v > vmx[cc]) { vmx[cc] = v; imx[cc] = s0 + j; }
                            if (ok && v < vmn[cc]) { vmn[cc] = v; imn[cc] = s0 + j; }
                        }
                    }
                    constexpr int LANES = TX >= 16 ? 16 : TX;   // threads sharing a centre within a column chunk
                    if (TX < 16) {                               // P_T == 64: the two chunks are the same centre
                        if (vmx[1] > vmx[0]) { vmx[0] = vmx[1]; imx[0] = imx[1]; }
                        if (vmn[1] < vmn[0]) { vmn[0] = vmn[1]; imn[0] = imn[1]; }
                    }
#pragma unroll
                    for (int cc = 0; cc < (TX < 16 ? 1 : 2); ++cc) {
#pragma unroll
                        for (int o = LANES / 2; o > 0; o >>= 1) {
                            const float ov = __shfl_xor_sync(OGC_FULL_MASK, vmx[cc], o);
                            const int oi = __shfl_xor_sync(OGC_FULL_MASK, imx[cc], o);
                            if (ov > vmx[cc] || (ov == vmx[cc] && oi < imx[cc])) { vmx[cc] = ov; imx[cc] = oi; }
                            const float uv = __shfl_xor_sync(OGC_FULL_MASK, vmn[cc], o);
                            const int ui = __shfl_xor_sync(OGC_FULL_MASK, imn[cc], o);
                            if (uv < vmn[cc] || (uv == vmn[cc] && ui < imn[cc])) { vmn[cc] = uv; imn[cc] = ui; }
                        }
                        const int p = p_base + cols[cc];
                        if ((tx % LANES) == 0 && r < Cout && p < q.P) {
                            const size_t o = (static_cast<size_t>(b) * Cout + r) * q.M + p / q.S;
                            q.ymax[o] = vmx[cc]; q.ymin[o] = vmn[cc];
                            q.amax[o] = static_cast<unsigned char>(imx[cc]);
                            q.amin[o] = static_cast<unsigned char>(imn[cc]);
                        }
                    }
                }
        }
    }
    // group sums: threads sharing ty share their two groups; reduce over the tx lanes, then fp64 atomics
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

// sums (B,4,2) -> per-channel scale/shift (B,C,2) and per-group mean/rstd (B,4,2).  n = (C/4) * P.
__global__ void gn_finalize_kernel(int B, int C, double n, const double *__restrict__ sums, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float *__restrict__ ss, float *__restrict__ mean_rstd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i - b * C, g = c / (C / kGnGroups);
    const double mean = sums[(b * kGnGroups + g) * 2] / n;
    double var = sums[(b * kGnGroups + g) * 2 + 1] / n - mean * mean;   // biased, as nn.GroupNorm
    var = var > 0 ? var : 0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kGnEps)));
    const float sc = gamma[c] * rstd;
    ss[i * 2] = sc;
    ss[i * 2 + 1] = beta[c] - static_cast<float>(mean) * sc;
    if (c % (C / kGnGroups) == 0) {
        mean_rstd[(b * kGnGroups + g) * 2] = static_cast<float>(mean);
        mean_rstd[(b * kGnGroups + g) * 2 + 1] = rstd;
    }
}

// out = relu(max_s(scale*y+shift)) = relu(scale * (scale >= 0 ? ymax : ymin) + shift); sel = position of
// the winner within the nsample axis (255: clamped by the ReLU -> no gradient); ysel = its pre-norm value.
__global__ void sa_finish_kernel(int B, int C, int M, const float *__restrict__ ymax, const float *__restrict__ ymin,
                                 const unsigned char *__restrict__ amax, const unsigned char *__restrict__ amin,
                                 const float *__restrict__ ss, float *__restrict__ out, float *__restrict__ out_pm,
                                 int c_total, int c_offset, unsigned char *__restrict__ sel, float *__restrict__ ysel) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<size_t>(B) * C * M) return;
    const int m = static_cast<int>(i % M);
    const size_t bc = i / M;
    const int c = static_cast<int>(bc % C), b = static_cast<int>(bc / C);
    const float sc = ss[bc * 2], sh = ss[bc * 2 + 1];
    const bool up = sc >= 0.f;
    const float yv = up ? ymax[i] : ymin[i];
    const float z = fmaf(sc, yv, sh);
    const float o = fmaxf(z, 0.f);
    out[(static_cast<size_t>(b) * c_total + c_offset + c) * M + m] = o;
    if (out_pm) out_pm[(static_cast<size_t>(b) * M + m) * c_total + c_offset + c] = o;
    sel[i] = z > 0.f ? (up ? amax[i] : amin[i]) : static_cast<unsigned char>(255);
    ysel[i] = yv;
}

template <int R_T, int P_T>
static cudaError_t launch_fwd(const MlpFwdParams &q, int B, bool gather, bool last, cudaStream_t st) {
    const size_t smem = (static_cast<size_t>(q.Cin) * R_T + static_cast<size_t>(q.Cin) * (P_T + 4)) * sizeof(float);
    if (smem > static_cast<size_t>(kMaxSmemPerCta) - 1024) return cudaErrorInvalidValue;
    const int ntiles = (q.P + P_T - 1) / P_T;
    int per_sample = (kNumSMs * 2) / B;   // floor: the whole grid must be resident at 2 CTAs/SM (a 297th CTA would run as a second wave)
    if (per_sample > ntiles) per_sample = ntiles;
    if (per_sample < 1) per_sample = 1;
    dim3 grid(per_sample, B);
#define OGC_LAUNCH(G, L)                                                                                        \
    do {                                                                                                        \
        cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel<R_T, P_T, G, L>,                                    \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
        if (e != cudaSuccess) return e;                                                                         \
        mlp_fwd_kernel<R_T, P_T, G, L><<<grid, kMlpThreads, smem, st>>>(q);                                     \
    } while (0)
    if (gather && last) OGC_LAUNCH(true, true);
    else if (gather) OGC_LAUNCH(true, false);
    else if (last) OGC_LAUNCH(false, true);
    else OGC_LAUNCH(false, false);
#undef OGC_LAUNCH
    return cudaGetLastError();
}

}  // namespace ogc

// One SharedMLP layer over grouped neighbourhoods (see the header of this file and include/ogc_b200.h).
extern "C" int ogc_sa_mlp_layer_fwd(int b, int n, int m, int nsample, int cin, int cout, int gather, int last,
                                    const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                                    const float *y_prev, const float *ss_prev, const float *wt, float *y, double *sums,
                                    float *ymax, float *ymin, unsigned char *amax, unsigned char *amin, void *stream) {
    using namespace ogc;
    if (b < 0 || m <= 0 || nsample <= 0 || cin <= 0 || cout <= 0 || !wt || !sums) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (cout % 16 != 0 || cout > 256) return OGC_ERR_UNSUPPORTED;
    if (last && (nsample != 64 || !ymax || !ymin || !amax || !amin)) return last && nsample != 64 ? OGC_ERR_UNSUPPORTED : OGC_ERR_INVALID_ARG;
    if (gather && (!xyz || !new_xyz || !idx || cin < 3 || (cin > 3 && !feat_pm))) return OGC_ERR_INVALID_ARG;
    if (!gather && (!y_prev || !ss_prev)) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    MlpFwdParams q;
    q.Cin = cin; q.Cout = cout; q.P = m * nsample; q.S = nsample; q.M = m; q.N = n; q.Cf = cin - 3;
    q.xyz = xyz; q.new_xyz = new_xyz; q.feat_pm = feat_pm; q.idx = idx; q.y_prev = y_prev; q.ss_prev = ss_prev;
    q.Wt = wt; q.y = y; q.sums = sums; q.ymax = ymax; q.ymin = ymin; q.amax = amax; q.amin = amin;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (cout <= 32) e = launch_fwd<32, 512>(q, b, gather, last, st);
    else if (cout <= 64) e = launch_fwd<64, 256>(q, b, gather, last, st);
    else if (cout <= 128) e = launch_fwd<128, 128>(q, b, gather, last, st);
    else e = launch_fwd<256, 64>(q, b, gather, last, st);
    if (e == cudaErrorInvalidValue) return OGC_ERR_UNSUPPORTED;
    return e == cudaSuccess ? OGC_OK : static_cast<int>(e);
}

extern "C" int ogc_gn_finalize(int b, int c, long long count_per_group, const double *sums, const float *gamma,
                               const float *beta, float *scale_shift, float *mean_rstd, void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || c % kGnGroups != 0 || count_per_group <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!sums || !gamma || !beta || !scale_shift || !mean_rstd) return OGC_ERR_INVALID_ARG;
    gn_finalize_kernel<<<(b * c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        b, c, static_cast<double>(count_per_group), sums, gamma, beta, scale_shift, mean_rstd);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_sa_finish(int b, int c, int m, const float *ymax, const float *ymin, const unsigned char *amax,
                             const unsigned char *amin, const float *scale_shift, float *out, float *out_pm,
                             int c_total, int c_offset, unsigned char *sel, float *ysel, void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || m <= 0 || c_offset < 0 || c_offset + c > c_total) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!ymax || !ymin || !amax || !amin || !scale_shift || !out || !sel || !ysel) return OGC_ERR_INVALID_ARG;
    const size_t total = static_cast<size_t>(b) * c * m;
    sa_finish_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        b, c, m, ymax, ymin, amax, amin, scale_shift, out, out_pm, c_total, c_offset, sel, ysel);
    OGC_RETURN_LAUNCH_STATUS();
}
