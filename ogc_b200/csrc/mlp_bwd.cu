// Backward of the fused set-abstraction MLP (see mlp.cu for the forward and the notation).
//
// Replaces what autograd runs for the reference's SA module (utils/pointnet2_util.py:33-44): max_pool2d
// backward, 3x (ReLU backward, native_group_norm_backward, two GEMMs), cat/sub backward and
// group_points_grad (pointnet2/src/group_points_gpu.cu:8-25) -- each a pass over (B,C,M,S) tensors.
//
// Per layer l (last to first):
//   dY_l is never stored: both kernels below rebuild it in their loaders from dz_l (for the last layer
//   synthesised from the pooled gradient and the recorded arg-max position), the stored pre-norm y_l and
//   per-channel coefficients  dY = k1 dz - k2 - (y - mean) k3r   (GroupNorm backward folded into 3 FMAs).
//   mlp_dw_kernel : dW_l[co][ci] += sum_p dY_l[co][p] a_{l-1}[ci][p]      (a_0 gathered on the fly)
//   mlp_dx_kernel : dz_{l-1} = relu'(z_{l-1}) * (W_l^T dY_l)  + the GroupNorm-backward sums of layer l-1
//                   (layer 1: scatter-add W_1^T dY_1 into the point-major feature gradient instead).
#include "mlp_common.cuh"
#include "mlp_dy.cuh"

namespace ogc {

constexpr int kDwPK = 32;   // positions per reduction chunk of the dW kernel (static smem <= 48 KB)

// ---- sparse GroupNorm-backward sums of the LAST layer (dz is non-zero only at the arg-max positions) ----
// ab (B,4,2) += [sum gamma dz, sum gamma dz yhat];  dgamma[c] += sum dz yhat;  dbeta[c] += sum dz
__global__ void __launch_bounds__(128)
sa_last_stats_kernel(int C, int M, const float *__restrict__ go, int go_ctotal, int go_coff,
                     const unsigned char *__restrict__ sel, const float *__restrict__ ysel,
                     const float *__restrict__ mean_rstd, const float *__restrict__ gamma, double *__restrict__ ab,
                     float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int c = blockIdx.x, b = blockIdx.y, g = c / (C / kGnGroups);
    const float mean = mean_rstd[(b * kGnGroups + g) * 2], rstd = mean_rstd[(b * kGnGroups + g) * 2 + 1];
    float s = 0.f, sy = 0.f;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const size_t i = (static_cast<size_t>(b) * C + c) * M + m;
        if (sel[i] != 255) {
            const float dz = go[(static_cast<size_t>(b) * go_ctotal + go_coff + c) * M + m];
            s += dz;
            sy += dz * ((ysel[i] - mean) * rstd);
        }
    }
    __shared__ float red[2][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(OGC_FULL_MASK, s, o);
        sy += __shfl_xor_sync(OGC_FULL_MASK, sy, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = sy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        s = red[0][0] + red[0][1] + red[0][2] + red[0][3];
        sy = red[1][0] + red[1][1] + red[1][2] + red[1][3];
        atomicAdd(dbeta + c, s);
        atomicAdd(dgamma + c, sy);
        atomicAdd(ab + (b * kGnGroups + g) * 2, static_cast<double>(gamma[c]) * s);
        atomicAdd(ab + (b * kGnGroups + g) * 2 + 1, static_cast<double>(gamma[c]) * sy);
    }
}

// coef (B,C,4) = [k1 = rstd gamma, k2 = rstd A/n, k3r = rstd^2 B/n, mean]
__global__ void gn_bwd_coef_kernel(int B, int C, double n, const double *__restrict__ ab, const float *__restrict__ mean_rstd,
                                   const float *__restrict__ gamma, float *__restrict__ coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i - b * C, g = c / (C / kGnGroups);
    const double rstd = mean_rstd[(b * kGnGroups + g) * 2 + 1];
    coef[i * 4 + 0] = static_cast<float>(rstd * gamma[c]);
    coef[i * 4 + 1] = static_cast<float>(rstd * ab[(b * kGnGroups + g) * 2] / n);
    coef[i * 4 + 2] = static_cast<float>(rstd * rstd * ab[(b * kGnGroups + g) * 2 + 1] / n);
    coef[i * 4 + 3] = mean_rstd[(b * kGnGroups + g) * 2];
}

// ------------------------------------------------------------------------------------------ dX
struct MlpDxParams {
    DySrc dy;                    // layer l (K = dy.C output channels of layer l)
    int cin_full, row_off, rows; // W is (K, cin_full); output rows = W columns [row_off, row_off+rows)
    const float *W;
    // dense output (layer l-1 has a GroupNorm+ReLU in front of it)
    const float *y_prev, *ss_prev, *mean_rstd_prev, *gamma_prev;   // (B,rows,P), (B,rows,2), (B,4,2), (rows)
    float *dz_prev;              // (B,rows,P)
    double *ab_prev;             // (B,4,2)
    float *dgamma_prev, *dbeta_prev;
    // scatter output (layer 1): gradient of the point-major features
    const int *idx;              // (B,M,S)
    float *dfeat_pm;             // (B,N,dfeat_stride), channels [dfeat_off, dfeat_off+rows)
    int N, dfeat_stride, dfeat_off;
    // dense mode over a layer wider than one CTA's 128 rows: this launch covers channels [prev_off, prev_off+rows)
    // of prev_total (GroupNorm groups are defined on prev_total)
    int prev_total, prev_off;
    // plain output (feature-propagation layer 0: no norm / activation in front): dx (B,dx_ctotal,P), channels
    // [dx_coff, dx_coff+rows)
    float *dx;
    int dx_ctotal, dx_coff;
};

enum { kDxDense = 0, kDxScatter = 1, kDxPlain = 2 };

template <int R_T, int P_T, int MODE>
__global__ void __launch_bounds__(kMlpThreads, 2)
mlp_dx_kernel(MlpDxParams q) {
    constexpr bool SCATTER = MODE == kDxScatter;
    constexpr int TX = P_T / 8, LDB = P_T + 4, KC = 64;
    extern __shared__ __align__(16) float smem[];
    __shared__ float rowacc[R_T][2];
    float *As = smem;               // [KC][R_T]
    float *Bs = smem + KC * R_T;    // [KC][LDB]
    const int tid = threadIdx.x, ty = tid / TX, tx = tid % TX;
    const int b = blockIdx.y;
    const int K = q.dy.C, P = q.dy.P;
    const int rows[2] = {ty * 4, R_T / 2 + ty * 4};
    const int cols[2] = {tx * 4, P_T / 2 + tx * 4};
    const int ntiles = (P + P_T - 1) / P_T;
    for (int i = tid; i < R_T * 2; i += kMlpThreads) (&rowacc[0][0])[i] = 0.f;

    float sdz[8], sdzy[8], sc[8], sh[8], mu[8], rs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sdz[i] = sdzy[i] = 0.f;
        sc[i] = sh[i] = mu[i] = rs[i] = 0.f;
        const int r = rows[i >> 2] + (i & 3);
        if (MODE == kDxDense && r < q.rows) {
            const int g = (q.prev_off + r) / (q.prev_total / kGnGroups);
            sc[i] = __ldg(q.ss_prev + (static_cast<size_t>(b) * q.prev_total + q.prev_off + r) * 2);
            sh[i] = __ldg(q.ss_prev + (static_cast<size_t>(b) * q.prev_total + q.prev_off + r) * 2 + 1);
            mu[i] = __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2);
            rs[i] = __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2 + 1);
        }
    }

    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int p_base = t * P_T;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int kc = 0; kc < K; kc += KC) {
            const int kn = min(KC, K - kc);
            __syncthreads();
            for (int e = tid; e < kn * R_T; e += kMlpThreads) {
                const int k = e / R_T, r = e - k * R_T;
                As[e] = r < q.rows ? __ldg(q.W + static_cast<size_t>(kc + k) * q.cin_full + q.row_off + r) : 0.f;
            }
            constexpr int Q4 = P_T / 4;
            // 4 independent (y, dz, coef) quads in flight per thread before any is consumed: the loader is
            // latency-bound at 1-2 CTAs per SM
            for (int e0 = tid; e0 < kn * Q4; e0 += kMlpThreads * 4) {
                DyRaw raw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * kMlpThreads;
                    if (e < kn * Q4) {
                        const int k = e / Q4, p = (e - k * Q4) * 4;
                        dy_quad_load(q.dy, b, kc + k, p_base + p, raw[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * kMlpThreads;
                    if (e < kn * Q4) {
                        const int k = e / Q4, p = (e - k * Q4) * 4;
                        *reinterpret_cast<float4 *>(Bs + k * LDB + p) = dy_quad_finish(q.dy, p_base + p, raw[u]);
                    }
                }
            }
            __syncthreads();
            tile_gemm<R_T, P_T>(As, Bs, LDB, kn, ty, tx, acc);
        }
        // ---- epilogue ----
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rows[i >> 2] + (i & 3);
            if (r >= q.rows) continue;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int p = p_base + cols[cc];
                if (p >= P) continue;
                if (SCATTER) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (p + j < P) {
                            const int pt = __ldg(q.idx + static_cast<size_t>(b) * P + p + j);
                            atomicAdd(q.dfeat_pm + (static_cast<size_t>(b) * q.N + pt) * q.dfeat_stride + q.dfeat_off + r, acc[i][cc * 4 + j]);
                        }
                } else if (MODE == kDxPlain) {
                    float *dp = q.dx + (static_cast<size_t>(b) * q.dx_ctotal + q.dx_coff + r) * P + p;
                    if (p + 3 < P && (reinterpret_cast<uintptr_t>(dp) & 15u) == 0)
                        *reinterpret_cast<float4 *>(dp) = make_float4(acc[i][cc * 4], acc[i][cc * 4 + 1], acc[i][cc * 4 + 2], acc[i][cc * 4 + 3]);
                    else
                        for (int j = 0; j < 4; ++j)
                            if (p + j < P) dp[j] = acc[i][cc * 4 + j];
                } else {
                    const float *yp = q.y_prev + (static_cast<size_t>(b) * q.prev_total + q.prev_off + r) * P + p;
                    float *dp = q.dz_prev + (static_cast<size_t>(b) * q.prev_total + q.prev_off + r) * P + p;
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool ok = p + j < P;
                        const float yv = ok ? __ldg(yp + j) : 0.f;
                        const float g = (ok && fmaf(sc[i], yv, sh[i]) > 0.f) ? acc[i][cc * 4 + j] : 0.f;
                        o[j] = g;
                        sdz[i] += g;
                        sdzy[i] += g * ((yv - mu[i]) * rs[i]);
                    }
                    if (p + 3 < P && (reinterpret_cast<uintptr_t>(dp) & 15u) == 0) *reinterpret_cast<float4 *>(dp) = make_float4(o[0], o[1], o[2], o[3]);
                    else
                        for (int j = 0; j < 4; ++j)
                            if (p + j < P) dp[j] = o[j];
                }
            }
        }
    }
    if (MODE != kDxDense) return;
    // per-channel sums over this CTA's positions -> dgamma / dbeta and the group sums of layer l-1
    constexpr int W = TX >= 32 ? 32 : TX;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float a = sdz[i], c = sdzy[i];
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) {
            a += __shfl_xor_sync(OGC_FULL_MASK, a, o);
            c += __shfl_xor_sync(OGC_FULL_MASK, c, o);
        }
        const int r = rows[i >> 2] + (i & 3);
        if ((tid % W) == 0 && r < q.rows) {
            atomicAdd(&rowacc[r][0], a);
            atomicAdd(&rowacc[r][1], c);
        }
    }
    __syncthreads();
    for (int r = tid; r < q.rows; r += kMlpThreads) {
        const float a = rowacc[r][0], c = rowacc[r][1];
        atomicAdd(q.dbeta_prev + q.prev_off + r, a);
        atomicAdd(q.dgamma_prev + q.prev_off + r, c);
        const int g = (q.prev_off + r) / (q.prev_total / kGnGroups);
        const double gm = static_cast<double>(__ldg(q.gamma_prev + q.prev_off + r));
        atomicAdd(q.ab_prev + (b * kGnGroups + g) * 2, gm * a);
        atomicAdd(q.ab_prev + (b * kGnGroups + g) * 2 + 1, gm * c);
    }
}

// ------------------------------------------------------------------------------------------ dW
struct MlpDwParams {
    DySrc dy;                       // layer l: rows of dW = dy.C
    int Cin, B;
    // a_{l-1}: dense (y_prev with GroupNorm+ReLU) or gathered
    const float *y_prev, *ss_prev;  // (B,Cin,P), (B,Cin,2)
    const float *xyz, *new_xyz, *feat_pm;
    const int *idx;
    int N, Cf;
    float *dW;                      // (Cout, Cin), accumulated atomically
};

template <int RPT, int NC, bool GATHER>
__global__ void __launch_bounds__(kMlpThreads, 2)
mlp_dw_kernel(MlpDwParams q) {
    constexpr int R_T = 16 * RPT, C_T = 16 * NC, PK = kDwPK;
    __shared__ __align__(16) float As[PK][R_T];   // dY^T  [p][co]
    __shared__ __align__(16) float Bs[PK][C_T];   // a^T   [p][ci]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int row0 = blockIdx.z * R_T, col0 = blockIdx.y * C_T;
    const int Cout = q.dy.C, P = q.dy.P;
    const int chunks_per_sample = (P + PK - 1) / PK;
    const int total = q.B * chunks_per_sample;
    float acc[RPT][NC];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[i][j] = 0.f;

    for (int ch = blockIdx.x; ch < total; ch += gridDim.x) {
        const int b = ch / chunks_per_sample, p_base = (ch - b * chunks_per_sample) * PK;
        __syncthreads();
        // dY^T: thread -> (channel fastest, position quad)
        for (int e0 = tid; e0 < R_T * (PK / 4); e0 += kMlpThreads * 4) {
            DyRaw raw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * kMlpThreads;
                const int c = e % R_T, pq = e / R_T;
                if (e < R_T * (PK / 4) && row0 + c < Cout) dy_quad_load(q.dy, b, row0 + c, p_base + pq * 4, raw[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * kMlpThreads;
                if (e >= R_T * (PK / 4)) continue;
                const int c = e % R_T, pq = e / R_T;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + c < Cout) v = dy_quad_finish(q.dy, p_base + pq * 4, raw[u]);
                As[pq * 4 + 0][c] = v.x; As[pq * 4 + 1][c] = v.y; As[pq * 4 + 2][c] = v.z; As[pq * 4 + 3][c] = v.w;
            }
        }
        if (GATHER) {
            const int lane = tid & 31, warp = tid >> 5;
            for (int p = warp; p < PK; p += kMlpThreads / 32) {
                const int gp = p_base + p;
                if (gp < P) {
                    const int j = __ldg(q.idx + static_cast<size_t>(b) * P + gp);
                    const int m = gp / q.dy.S;
                    for (int c = lane; c < C_T; c += 32) {
                        const int ci = col0 + c;
                        float v = 0.f;
                        if (ci < 3) v = __ldg(q.xyz + (static_cast<size_t>(b) * q.N + j) * 3 + ci) - __ldg(q.new_xyz + (static_cast<size_t>(b) * q.dy.M + m) * 3 + ci);
                        else if (ci < q.Cin) v = __ldg(q.feat_pm + (static_cast<size_t>(b) * q.N + j) * q.Cf + ci - 3);
                        Bs[p][c] = v;
                    }
                } else {
                    for (int c = lane; c < C_T; c += 32) Bs[p][c] = 0.f;
                }
            }
        } else {
            for (int e = tid; e < C_T * (PK / 4); e += kMlpThreads) {
                const int c = e % C_T, pq = e / C_T, ci = col0 + c, gp = p_base + pq * 4;
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                if (ci < q.Cin && !q.ss_prev) {           // raw layer input (feature propagation layer 0)
                    const float *yp = q.y_prev + (static_cast<size_t>(b) * q.Cin + ci) * P + gp;
                    if (gp + 3 < P && (reinterpret_cast<uintptr_t>(yp) & 15u) == 0) {
                        const float4 t = __ldg(reinterpret_cast<const float4 *>(yp));
                        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
                    } else {
                        for (int j = 0; j < 4; ++j)
                            if (gp + j < P) o[j] = __ldg(yp + j);
                    }
                } else if (ci < q.Cin) {
                    const float s = __ldg(q.ss_prev + (static_cast<size_t>(b) * q.Cin + ci) * 2), h = __ldg(q.ss_prev + (static_cast<size_t>(b) * q.Cin + ci) * 2 + 1);
                    const float *yp = q.y_prev + (static_cast<size_t>(b) * q.Cin + ci) * P + gp;
                    if (gp + 3 < P && (reinterpret_cast<uintptr_t>(yp) & 15u) == 0) {
                        const float4 t = __ldg(reinterpret_cast<const float4 *>(yp));
                        o[0] = fmaxf(fmaf(s, t.x, h), 0.f); o[1] = fmaxf(fmaf(s, t.y, h), 0.f);
                        o[2] = fmaxf(fmaf(s, t.z, h), 0.f); o[3] = fmaxf(fmaf(s, t.w, h), 0.f);
                    } else {
                        for (int j = 0; j < 4; ++j)
                            if (gp + j < P) o[j] = fmaxf(fmaf(s, __ldg(yp + j), h), 0.f);
                    }
                }
                Bs[pq * 4 + 0][c] = o[0]; Bs[pq * 4 + 1][c] = o[1]; Bs[pq * 4 + 2][c] = o[2]; Bs[pq * 4 + 3][c] = o[3];
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < PK; ++k) {
            float a[RPT], bb[NC];
#pragma unroll
            for (int i = 0; i < RPT; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < NC; ++j) bb[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
                for (int j = 0; j < NC; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = row0 + ty + 16 * i;
        if (r >= Cout) continue;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const int c = col0 + tx + 16 * j;
            if (c < q.Cin) atomicAdd(q.dW + static_cast<size_t>(r) * q.Cin + c, acc[i][j]);
        }
    }
}

// dW of a gathered layer with a narrow input (SA level 1: 3 centred + 3 raw coordinates, C_in = 6).  The generic
// split-K kernel wastes its 16x16 thread grid on a 32x6 output; here a warp owns CPW output channels, its lanes
// stream the positions with coalesced 16-byte loads of dz / y, the gathered input tile sits in shared memory and
// the CPW x CIN partial sums stay in registers across all tiles of the CTA (one reduction at the very end).
template <int CIN, int CPW>
__global__ void __launch_bounds__(256, CPW <= 4 ? 2 : 1)
mlp_dw_small_kernel(MlpDwParams q) {
    constexpr int TP = 256;                       // positions per tile: 64 quads, two per lane
    __shared__ __align__(16) float a0[CIN][TP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = q.dy.P, Cout = q.dy.C;
    const int tiles_per_sample = (P + TP - 1) / TP;
    const int total = q.B * tiles_per_sample;
    float acc[CPW][CIN];
#pragma unroll
    for (int i = 0; i < CPW; ++i)
#pragma unroll
        for (int j = 0; j < CIN; ++j) acc[i][j] = 0.f;

    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int b = w / tiles_per_sample, p_base = (w - b * tiles_per_sample) * TP;
        __syncthreads();
        {
            const int gp = p_base + tid;
            float v[CIN];
#pragma unroll
            for (int c = 0; c < CIN; ++c) v[c] = 0.f;
            if (gp < P) {
                const int j = __ldg(q.idx + static_cast<size_t>(b) * P + gp);
                const int m = gp / q.dy.S;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[c] = __ldg(q.xyz + (static_cast<size_t>(b) * q.N + j) * 3 + c) - __ldg(q.new_xyz + (static_cast<size_t>(b) * q.dy.M + m) * 3 + c);
#pragma unroll
                for (int c = 3; c < CIN; ++c) v[c] = __ldg(q.feat_pm + (static_cast<size_t>(b) * q.N + j) * q.Cf + c - 3);
            }
#pragma unroll
            for (int c = 0; c < CIN; ++c) a0[c][tid] = v[c];
        }
        __syncthreads();
        DyRaw raw[CPW][2];
#pragma unroll
        for (int i = 0; i < CPW; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int co = warp * CPW + i;
                if (co < Cout) dy_quad_load(q.dy, b, co, p_base + (lane + 32 * h) * 4, raw[i][h]);
            }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 av[CIN];
#pragma unroll
            for (int c = 0; c < CIN; ++c) av[c] = *reinterpret_cast<const float4 *>(&a0[c][(lane + 32 * h) * 4]);
#pragma unroll
            for (int i = 0; i < CPW; ++i) {
                const int co = warp * CPW + i;
                if (co >= Cout) continue;
                const float4 d = dy_quad_finish(q.dy, p_base + (lane + 32 * h) * 4, raw[i][h]);
#pragma unroll
                for (int c = 0; c < CIN; ++c)
                    acc[i][c] = fmaf(d.x, av[c].x, fmaf(d.y, av[c].y, fmaf(d.z, av[c].z, fmaf(d.w, av[c].w, acc[i][c]))));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CPW; ++i)
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            float v = acc[i][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(OGC_FULL_MASK, v, o);
            const int co = warp * CPW + i;
            if (lane == 0 && co < Cout) atomicAdd(q.dW + static_cast<size_t>(co) * CIN + c, v);
        }
}

template <int R_T, int P_T>
static cudaError_t launch_dx(const MlpDxParams &q, int B, int mode, cudaStream_t st) {
    constexpr int KC = 64;
    const size_t smem = (static_cast<size_t>(KC) * R_T + static_cast<size_t>(KC) * (P_T + 4)) * sizeof(float);
    const int ntiles = (q.dy.P + P_T - 1) / P_T;
    int per_sample = (kNumSMs * 2) / B;   // floor: the whole grid must be resident at 2 CTAs/SM (a 297th CTA would run as a second wave)
    per_sample = per_sample > ntiles ? ntiles : (per_sample < 1 ? 1 : per_sample);
    dim3 grid(per_sample, B);
    cudaError_t e;
    if (mode == kDxScatter) {
        e = cudaFuncSetAttribute(mlp_dx_kernel<R_T, P_T, kDxScatter>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        mlp_dx_kernel<R_T, P_T, kDxScatter><<<grid, kMlpThreads, smem, st>>>(q);
    } else if (mode == kDxPlain) {
        e = cudaFuncSetAttribute(mlp_dx_kernel<R_T, P_T, kDxPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        mlp_dx_kernel<R_T, P_T, kDxPlain><<<grid, kMlpThreads, smem, st>>>(q);
    } else {
        e = cudaFuncSetAttribute(mlp_dx_kernel<R_T, P_T, kDxDense>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        mlp_dx_kernel<R_T, P_T, kDxDense><<<grid, kMlpThreads, smem, st>>>(q);
    }
    return cudaGetLastError();
}

template <int RPT, int NC>
static cudaError_t launch_dw(const MlpDwParams &q, bool gather, cudaStream_t st) {
    constexpr int R_T = 16 * RPT, C_T = 16 * NC;
    const int gy = (q.Cin + C_T - 1) / C_T, gz = (q.dy.C + R_T - 1) / R_T;
    const int total = q.B * ((q.dy.P + kDwPK - 1) / kDwPK);
    int gx = (kNumSMs * 2) / (gy * gz);
    gx = gx < 1 ? 1 : (gx > total ? total : gx);
    dim3 grid(gx, gy, gz);
    if (gather) mlp_dw_kernel<RPT, NC, true><<<grid, kMlpThreads, 0, st>>>(q);
    else mlp_dw_kernel<RPT, NC, false><<<grid, kMlpThreads, 0, st>>>(q);
    return cudaGetLastError();
}

template <int RPT>
static cudaError_t dispatch_dw_nc(const MlpDwParams &q, bool gather, cudaStream_t st) {
    const int cin = q.Cin;
    if (cin <= 16) return launch_dw<RPT, 1>(q, gather, st);
    if (cin <= 32) return launch_dw<RPT, 2>(q, gather, st);
    if (cin <= 64) return launch_dw<RPT, 4>(q, gather, st);
    if (cin > 128 && cin <= 144) return launch_dw<RPT, 9>(q, gather, st);
    return launch_dw<RPT, 8>(q, gather, st);     // <= 128 per column block; wider inputs use several blocks
}

}  // namespace ogc

extern "C" int ogc_sa_last_stats(int b, int c, int m, const float *go, int go_ctotal, int go_coff,
                                 const unsigned char *sel, const float *ysel, const float *mean_rstd,
                                 const float *gamma, double *ab, float *dgamma, float *dbeta, void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || c % kGnGroups != 0 || m <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!go || !sel || !ysel || !mean_rstd || !gamma || !ab || !dgamma || !dbeta) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid(c, b);
    sa_last_stats_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(c, m, go, go_ctotal, go_coff, sel, ysel,
                                                                             mean_rstd, gamma, ab, dgamma, dbeta);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_gn_bwd_coef(int b, int c, long long count_per_group, const double *ab, const float *mean_rstd,
                               const float *gamma, float *coef, void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || c % kGnGroups != 0 || count_per_group <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!ab || !mean_rstd || !gamma || !coef) return OGC_ERR_INVALID_ARG;
    gn_bwd_coef_kernel<<<(b * c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        b, c, static_cast<double>(count_per_group), ab, mean_rstd, gamma, coef);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_sa_mlp_layer_dx(int b, int n, int m, int nsample, int cout, int cin_full, int row_off, int rows,
                                   const float *dz, const float *go, int go_ctotal, int go_coff,
                                   const unsigned char *sel, const float *y, const float *coef, const float *w,
                                   const float *y_prev, const float *ss_prev, const float *mean_rstd_prev,
                                   const float *gamma_prev, float *dz_prev, double *ab_prev, float *dgamma_prev,
                                   float *dbeta_prev, const int *idx, float *dfeat_pm, int dfeat_stride, int dfeat_off,
                                   void *stream) {
    using namespace ogc;
    if (b < 0 || m <= 0 || nsample <= 0 || cout <= 0 || rows <= 0 || row_off < 0 || row_off + rows > cin_full || !w)
        return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    const bool scatter = dfeat_pm != nullptr;
    if (scatter && rows > 128) return OGC_ERR_UNSUPPORTED;
    MlpDxParams q;
    int rc = fill_dy(q.dy, cout, m, nsample, dz, go, go_ctotal, go_coff, sel, y, coef);
    if (rc != OGC_OK) return rc;
    if (scatter) {
        if (!idx) return OGC_ERR_INVALID_ARG;
    } else {
        if (!y_prev || !ss_prev || !mean_rstd_prev || !gamma_prev || !dz_prev || !ab_prev || !dgamma_prev || !dbeta_prev)
            return OGC_ERR_INVALID_ARG;
        if (rows % 16 != 0) return OGC_ERR_UNSUPPORTED;
    }
    q.cin_full = cin_full; q.W = w;
    q.y_prev = y_prev; q.ss_prev = ss_prev; q.mean_rstd_prev = mean_rstd_prev; q.gamma_prev = gamma_prev;
    q.dz_prev = dz_prev; q.ab_prev = ab_prev; q.dgamma_prev = dgamma_prev; q.dbeta_prev = dbeta_prev;
    q.idx = idx; q.dfeat_pm = dfeat_pm; q.N = n; q.dfeat_stride = dfeat_stride; q.dfeat_off = dfeat_off;
    q.dx = nullptr; q.dx_ctotal = q.dx_coff = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mode = scatter ? kDxScatter : kDxDense;
    // dense layers wider than 128 channels: one launch per 128-row block (GroupNorm groups span prev_total = rows)
    for (int off = 0; off < rows; off += 128) {
        const int chunk = rows - off < 128 ? rows - off : 128;
        q.row_off = row_off + off; q.rows = chunk; q.prev_total = rows; q.prev_off = off;
        cudaError_t e;
        if (chunk <= 32) e = launch_dx<32, 512>(q, b, mode, st);
        else if (chunk <= 64) e = launch_dx<64, 256>(q, b, mode, st);
        else e = launch_dx<128, 128>(q, b, mode, st);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    return OGC_OK;
}

extern "C" int ogc_pw_mlp_input_grad(int b, int p, int cout, int cin_full, int row_off, int rows, const float *dz,
                                     const float *y, const float *coef, const float *w, float *dx, int dx_ctotal,
                                     int dx_coff, void *stream) {
    using namespace ogc;
    if (b < 0 || p <= 0 || cout <= 0 || rows <= 0 || row_off < 0 || row_off + rows > cin_full || dx_coff < 0 ||
        dx_coff + rows > dx_ctotal)
        return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!dz || !y || !coef || !w || !dx) return OGC_ERR_INVALID_ARG;
    if (b > 65535 || rows > 128) return OGC_ERR_UNSUPPORTED;
    MlpDxParams q;
    int rc = fill_dy(q.dy, cout, p, 1, dz, nullptr, 0, 0, nullptr, y, coef);
    if (rc != OGC_OK) return rc;
    q.cin_full = cin_full; q.row_off = row_off; q.rows = rows; q.W = w;
    q.y_prev = q.ss_prev = q.mean_rstd_prev = q.gamma_prev = nullptr;
    q.dz_prev = nullptr; q.ab_prev = nullptr; q.dgamma_prev = q.dbeta_prev = nullptr;
    q.idx = nullptr; q.dfeat_pm = nullptr; q.N = 0; q.dfeat_stride = q.dfeat_off = 0;
    q.dx = dx; q.dx_ctotal = dx_ctotal; q.dx_coff = dx_coff; q.prev_total = rows; q.prev_off = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (rows <= 32) e = launch_dx<32, 512>(q, b, kDxPlain, st);
    else if (rows <= 64) e = launch_dx<64, 256>(q, b, kDxPlain, st);
    else e = launch_dx<128, 128>(q, b, kDxPlain, st);
    return e == cudaSuccess ? OGC_OK : static_cast<int>(e);
}

extern "C" int ogc_sa_mlp_layer_dw(int b, int n, int m, int nsample, int cout, int cin, int gather, const float *dz,
                                   const float *go, int go_ctotal, int go_coff, const unsigned char *sel,
                                   const float *y, const float *coef, const float *y_prev, const float *ss_prev,
                                   const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                                   float *dw, void *stream) {
    using namespace ogc;
    if (b < 0 || m <= 0 || nsample <= 0 || cout <= 0 || cin <= 0 || !dw) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    MlpDwParams q;
    int rc = fill_dy(q.dy, cout, m, nsample, dz, go, go_ctotal, go_coff, sel, y, coef);
    if (rc != OGC_OK) return rc;
    if (gather && (!xyz || !new_xyz || !idx || cin < 3 || (cin > 3 && !feat_pm))) return OGC_ERR_INVALID_ARG;
    if (!gather && !y_prev) return OGC_ERR_INVALID_ARG;    /* ss_prev == NULL: a_{l-1} = y_prev as is */
    q.Cin = cin; q.B = b; q.y_prev = y_prev; q.ss_prev = ss_prev; q.xyz = xyz; q.new_xyz = new_xyz;
    q.feat_pm = feat_pm; q.idx = idx; q.N = n; q.Cf = cin - 3; q.dW = dw;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (gather && cin == 6 && (cout == 32 || cout == 64) && (nsample % 4) == 0) {
        const int total = b * ((m * nsample + 255) / 256);
        // cout = 32: 128 registers, 2 CTAs/SM, one resident wave; cout = 64 keeps 8 x 6 sums per warp: 1 CTA/SM
        const int cap = cout == 32 ? kNumSMs * 2 : kNumSMs * 4;
        const int gx = total < cap ? total : cap;
        if (cout == 32) mlp_dw_small_kernel<6, 4><<<gx, 256, 0, st>>>(q);
        else mlp_dw_small_kernel<6, 8><<<gx, 256, 0, st>>>(q);
        OGC_RETURN_LAUNCH_STATUS();
    }
    if (cout <= 32) e = dispatch_dw_nc<2>(q, gather, st);
    else if (cout <= 64) e = dispatch_dw_nc<4>(q, gather, st);
    else e = dispatch_dw_nc<8>(q, gather, st);
    return e == cudaSuccess ? OGC_OK : static_cast<int>(e);
}
