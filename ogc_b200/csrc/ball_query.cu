// K10  ball query for sm_100a.
//
// Replaces ball_query_kernel_fast (pointnet2/src/ball_query_gpu.cu:9-45): one thread per centre
// there, scanning xyz from global memory and writing idx[] element by element.
//
// Here: one WARP per centre.  The candidate cloud is staged in shared memory by the TMA bulk-copy
// engine; the 32 lanes test 32 candidates per step in ascending index order, a ballot + prefix
// popcount gives every hit its output slot, the warp stops as soon as nsample hits are found
// (the reference's early `break`, :42), and the finished row -- hits, then the first hit repeated
// (:35-39), or zeros when there is no hit (pointnet2/pointnet2.py:251) -- is written with one
// coalesced store per 32 slots instead of scattered 4-byte stores.
#include "common.cuh"

namespace ogc {

constexpr int kBqThreads = 512;
constexpr int kBqWarps = kBqThreads / 32;
constexpr int kBqTilePoints = 8192;

__global__ void __launch_bounds__(kBqThreads)
ball_query_warp_kernel(int n, int m, float radius2, int nsample, int rounds, const float *__restrict__ new_xyz,
                       const float *__restrict__ xyz, int *__restrict__ idx) {
    extern __shared__ __align__(16) float bq_smem[];
    __shared__ __align__(8) uint64_t bar;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bi = blockIdx.y;
    new_xyz += static_cast<size_t>(bi) * m * 3;
    xyz += static_cast<size_t>(bi) * n * 3;
    idx += static_cast<size_t>(bi) * m * nsample;

    const int tile_cap = min(n, kBqTilePoints);
    int *rows = reinterpret_cast<int *>(bq_smem + tile_cap * 3 + 4);   // [kBqWarps][nsample]
    int *row = rows + warp * nsample;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t parity = 0;

    const int ntiles = (n + kBqTilePoints - 1) / kBqTilePoints;
    const int q_base = blockIdx.x * rounds * kBqWarps;
    const float *tile = nullptr;
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int rd = 0; rd < rounds; ++rd) {
        const int q = q_base + rd * kBqWarps + warp;
        const bool has_q = q < m;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        if (has_q) {
            cx = __ldg(new_xyz + q * 3 + 0);
            cy = __ldg(new_xyz + q * 3 + 1);
            cz = __ldg(new_xyz + q * 3 + 2);
        }
        int cnt = 0, first = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int t0 = t * kBqTilePoints;
            const int tn = min(kBqTilePoints, n - t0);
            if (ntiles > 1 || rd == 0) {
                if (t > 0 || rd > 0) __syncthreads();
                tile = stage_floats(bq_smem, xyz + static_cast<size_t>(t0) * 3, tn * 3, &bar, parity);
                __syncthreads();
            }
            if (!has_q || cnt >= nsample) continue;
            for (int j0 = 0; j0 < tn && cnt < nsample; j0 += 64) {
                // two independent 32-wide probes per step for ILP; consumed in ascending order
                bool hit[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = j0 + 32 * u + lane;
                    const int jc = min(j, tn - 1);
                    const float d2 = sqdist(cx, cy, cz, tile[jc * 3 + 0], tile[jc * 3 + 1], tile[jc * 3 + 2]);
                    hit[u] = (j < tn) && (d2 < radius2);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const unsigned mask = __ballot_sync(OGC_FULL_MASK, hit[u]);
                    if (mask && cnt < nsample) {
                        const int slot = cnt + __popc(mask & lt_mask);
                        if (hit[u] && slot < nsample) row[slot] = t0 + j0 + 32 * u + lane;
                        if (cnt == 0) first = t0 + j0 + 32 * u + __ffs(mask) - 1;
                        cnt += __popc(mask);
                    }
                }
            }
        }
        if (has_q) {
            __syncwarp();
            const int filled = min(cnt, nsample);
            for (int s = lane; s < nsample; s += 32)
                idx[static_cast<size_t>(q) * nsample + s] = s < filled ? row[s] : first;  // first == 0 if no hit
            __syncwarp();
        }
    }
}

}  // namespace ogc

extern "C" int ogc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                              const float *xyz, int *idx, void *stream) {
    using namespace ogc;
    if (b < 0 || n < 0 || m < 0 || nsample < 0) return OGC_ERR_INVALID_ARG;
    if (b == 0 || m == 0 || nsample == 0) return OGC_OK;
    if (!new_xyz || !idx || (n > 0 && !xyz)) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    const int tile_pts = n < kBqTilePoints ? n : kBqTilePoints;
    const size_t smem = (static_cast<size_t>(tile_pts) * 3 + 4) * sizeof(float) +
                        static_cast<size_t>(kBqWarps) * nsample * sizeof(int);
    if (smem > static_cast<size_t>(kMaxSmemPerCta)) return OGC_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaFuncSetAttribute(ball_query_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    const long long total_q = static_cast<long long>(b) * m;
    int rounds = static_cast<int>(total_q / (static_cast<long long>(kBqWarps) * kNumSMs * 4));
    rounds = rounds < 1 ? 1 : (rounds > 8 ? 8 : rounds);
    const int qpc = rounds * kBqWarps;
    dim3 grid((m + qpc - 1) / qpc, b);
    const float radius2 = radius * radius;  // fp32 product, as src/ball_query_gpu.cu:23
    ball_query_warp_kernel<<<grid, kBqThreads, smem, st>>>(n, m, radius2, nsample, rounds, new_xyz, xyz, idx);
    OGC_RETURN_LAUNCH_STATUS();
}
