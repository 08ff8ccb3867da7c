// Shared pieces of the fused SharedMLP kernels (mlp.cu: forward, mlp_bwd.cu: backward).
#pragma once
#include "common.cuh"

namespace ogc {

constexpr int kMlpThreads = 256;
constexpr int kGnGroups = 4;
constexpr float kGnEps = 1e-5f;  // nn.GroupNorm default (utils/nn_util.py:9)

// acc[i][j] += sum_k A[k][row_i] * B[k][col_j];  rows {ty*4..+3, R_T/2+ty*4..+3}, cols likewise with tx.
template <int R_T, int P_T>
__device__ __forceinline__ void tile_gemm(const float *__restrict__ As, const float *__restrict__ Bs, int ldb, int K,
                                          int ty, int tx, float (&acc)[8][8]) {
    const float *a0p = As + ty * 4, *a1p = As + R_T / 2 + ty * 4;
    const float *b0p = Bs + tx * 4, *b1p = Bs + P_T / 2 + tx * 4;
    // packed FMAs (FFMA2): accumulator pairs (col 2j, col 2j+1) x broadcast a[i]
    float2 c2[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c2[i][j] = make_float2(acc[i][2 * j], acc[i][2 * j + 1]);
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a0 = *reinterpret_cast<const float4 *>(a0p + k * R_T);
        const float4 a1 = *reinterpret_cast<const float4 *>(a1p + k * R_T);
        const float4 b0 = *reinterpret_cast<const float4 *>(b0p + k * ldb);
        const float4 b1 = *reinterpret_cast<const float4 *>(b1p + k * ldb);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float2 b[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) c2[i][j] = ffma2(aa, b[j], c2[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][2 * j] = c2[i][j].x; acc[i][2 * j + 1] = c2[i][j].y; }
}

struct MlpFwdParams {
    int Cin, Cout, P, S, M, N, Cf;
    const float *xyz, *new_xyz, *feat_pm;  // gather mode: (B,N,3), (B,M,3), (B,N,Cf)
    const int *idx;                        //              (B,M,S)
    const float *y_prev, *ss_prev;         // dense mode:  (B,Cin,P), (B,Cin,2) = GroupNorm scale, shift
    const float *Wt;                       // (Cin,Cout)
    float *y;                              // (B,Cout,P) or NULL
    double *sums;                          // (B,4,2): sum y, sum y^2 per group
    float *ymax, *ymin;                    // (B,Cout,M) when LAST
    unsigned char *amax, *amin;
};

}  // namespace ogc
