// K4 / K5  exact k-nearest-neighbour search (and three_nn = k 3) for sm_100a.
//
// Replaces knn_kernel_fast / three_nn_kernel_fast of the reference
// (pointnet2/src/interpolate_gpu.cu:9-57, :81-124): one THREAD per query there, scanning all m
// candidates from global memory and insertion-sorting into a per-thread `double best[200]`
// that lives in local memory.
//
// Here: one WARP per query.
//   * the candidate cloud is staged into shared memory once per CTA by the TMA bulk-copy engine
//     (cp.async.bulk, AoS xyz: lane stride 3 words is bank-conflict free);
//   * the 32 lanes evaluate 32 candidates per step (x4 unrolled) with the reference's rounding
//     order (common.cuh sqdist) and filter them against the current k-th best distance;
//   * the running top-k is a sorted list distributed over the lanes' registers
//     (element i -> register i/32 of lane i%32); a surviving candidate is inserted with one
//     ballot/popc to find its rank and a shuffle-up to make room.
// Candidates are visited in ascending index order and inserted after every element with
// distance <= theirs, which reproduces the reference's strict '<' insertion (:41-51): the result
// is the k smallest candidates under (distance, index) order, ascending; NaN / inf distances are
// never inserted; if m < k the tail stays (+inf, 0) -- (float)1e40 in the reference (:32-35,:54).
#include "common.cuh"

namespace ogc {

constexpr int kKnnThreads = 512;                 // 16 warps = 16 queries in flight per CTA
constexpr int kKnnWarps = kKnnThreads / 32;
constexpr int kKnnTilePoints = 8192;             // 96 KB of candidates per stage -> 2 CTAs / SM

template <int KR>
struct TopK {
    float d[KR];
    int i[KR];
    float tau;   // admission threshold: min(distance of element k-1, tau0) (warp-uniform)
    float tau0;  // caller's bound on useful squared distances (+inf = exact k-NN)

    __device__ __forceinline__ void init(float bound) {
#pragma unroll
        for (int r = 0; r < KR; ++r) { d[r] = __int_as_float(0x7f800000); i[r] = 0; }
        tau0 = bound;
        tau = bound;
    }

    // Warp-uniform call: insert (cd, ci); precondition cd < tau.
    __device__ __forceinline__ void insert(float cd, int ci, int k, int lane) {
        int p = 0;
#pragma unroll
        for (int r = 0; r < KR; ++r) p += __popc(__ballot_sync(OGC_FULL_MASK, d[r] <= cd));
#pragma unroll
        for (int r = KR - 1; r >= 0; --r) {
            const int lo = 32 * r;
            if (p >= lo + 32) break;  // this and all lower registers are untouched (uniform)
            const float ud = __shfl_up_sync(OGC_FULL_MASK, d[r], 1);
            const int ui = __shfl_up_sync(OGC_FULL_MASK, i[r], 1);
            if (p >= lo) {
                const int lp = p - lo;
                if (lane > lp) { d[r] = ud; i[r] = ui; }
                else if (lane == lp) { d[r] = cd; i[r] = ci; }
            } else {
                // whole register shifts by one; lane 0 receives the last element of register r-1
                const float cdn = __shfl_sync(OGC_FULL_MASK, d[r > 0 ? r - 1 : 0], 31);
                const int cin = __shfl_sync(OGC_FULL_MASK, i[r > 0 ? r - 1 : 0], 31);
                if (lane > 0) { d[r] = ud; i[r] = ui; }
                else { d[r] = cdn; i[r] = cin; }
            }
        }
        const int kr = (k - 1) >> 5, kl = (k - 1) & 31;
        float t = d[0];
#pragma unroll
        for (int r = 1; r < KR; ++r) t = (kr == r) ? d[r] : t;
        tau = fminf(__shfl_sync(OGC_FULL_MASK, t, kl), tau0);
    }
};

template <int KR, bool SQRT_OUT>
__global__ void __launch_bounds__(kKnnThreads)
knn_warp_kernel(int n, int m, int k, int rounds, float tau0, const float *__restrict__ unknown,
                const float *__restrict__ known, float *__restrict__ dist_out, int *__restrict__ idx_out) {
    extern __shared__ __align__(16) float knn_smem[];
    __shared__ __align__(8) uint64_t bar;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bi = blockIdx.y;
    unknown += static_cast<size_t>(bi) * n * 3;
    known += static_cast<size_t>(bi) * m * 3;
    dist_out += static_cast<size_t>(bi) * n * k;
    idx_out += static_cast<size_t>(bi) * n * k;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t parity = 0;

    const int ntiles = (m + kKnnTilePoints - 1) / kKnnTilePoints;
    const int q_base = blockIdx.x * rounds * kKnnWarps;
    const float *tile = nullptr;

    for (int rd = 0; rd < rounds; ++rd) {
        const int q = q_base + rd * kKnnWarps + warp;
        const bool has_q = q < n;                       // warp-uniform
        float ux = 0.f, uy = 0.f, uz = 0.f;
        if (has_q) {
            ux = __ldg(unknown + q * 3 + 0);
            uy = __ldg(unknown + q * 3 + 1);
            uz = __ldg(unknown + q * 3 + 2);
        }
        TopK<KR> top;
        top.init(tau0);

        for (int t = 0; t < ntiles; ++t) {
            const int t0 = t * kKnnTilePoints;
            const int tn = min(kKnnTilePoints, m - t0);
            if (ntiles > 1 || rd == 0) {
                if (t > 0 || rd > 0) __syncthreads();   // everyone is done with the previous tile
                tile = stage_floats(knn_smem, known + static_cast<size_t>(t0) * 3, tn * 3, &bar, parity);
                __syncthreads();
            }
            if (!has_q) continue;
            // 128 candidates per step: slot u of lane l is candidate j0 + 32u + l (ascending order
            // = u-major, lane-minor, which is the order survivors are inserted in).
            for (int j0 = 0; j0 < tn; j0 += 128) {
                float dd[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u + lane;
                    const int jc = min(j, tn - 1);
                    const float d = sqdist(ux, uy, uz, tile[jc * 3 + 0], tile[jc * 3 + 1], tile[jc * 3 + 2]);
                    dd[u] = (j < tn) ? d : __int_as_float(0x7f800000);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    unsigned hits = __ballot_sync(OGC_FULL_MASK, dd[u] < top.tau);
                    while (hits) {
                        const int src = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const float cd = __shfl_sync(OGC_FULL_MASK, dd[u], src);
                        if (cd < top.tau)  // tau may have dropped since the ballot (uniform branch)
                            top.insert(cd, t0 + j0 + 32 * u + src, k, lane);
                    }
                }
            }
        }
        if (has_q) {
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                const int e = 32 * r + lane;
                if (e < k) {
                    const float d = top.d[r];
                    dist_out[static_cast<size_t>(q) * k + e] = SQRT_OUT ? __fsqrt_rn(d) : d;
                    idx_out[static_cast<size_t>(q) * k + e] = top.i[r];
                }
            }
        }
    }
}

template <int KR, bool SQ>
static cudaError_t launch_knn(int b, int n, int m, int k, float tau0, const float *unknown, const float *known,
                              float *dist, int *idx, cudaStream_t st) {
    const int tile_pts = m < kKnnTilePoints ? m : kKnnTilePoints;
    const size_t smem = (static_cast<size_t>(tile_pts) * 3 + 4) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(knn_warp_kernel<KR, SQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    // queries per CTA = rounds * 16; aim for >= ~4 CTAs per SM worth of blocks, at most 8 rounds
    const long long total_q = static_cast<long long>(b) * n;
    int rounds = static_cast<int>(total_q / (static_cast<long long>(kKnnWarps) * kNumSMs * 4));
    rounds = rounds < 1 ? 1 : (rounds > 8 ? 8 : rounds);
    const int qpc = rounds * kKnnWarps;
    dim3 grid((n + qpc - 1) / qpc, b);
    knn_warp_kernel<KR, SQ><<<grid, kKnnThreads, smem, st>>>(n, m, k, rounds, tau0, unknown, known, dist, idx);
    return cudaGetLastError();
}

template <bool SQ>
static int knn_dispatch(int b, int n, int m, int k, float tau0, const float *unknown, const float *known, float *dist,
                        int *idx, void *stream) {
    if (b < 0 || n < 0 || m < 0 || k < 1 || k > 224) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n == 0) return OGC_OK;
    if (!unknown || !dist || !idx || (m > 0 && !known)) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (k <= 32) e = launch_knn<1, SQ>(b, n, m, k, tau0, unknown, known, dist, idx, st);
    else if (k <= 64) e = launch_knn<2, SQ>(b, n, m, k, tau0, unknown, known, dist, idx, st);
    else if (k <= 128) e = launch_knn<4, SQ>(b, n, m, k, tau0, unknown, known, dist, idx, st);
    else e = launch_knn<7, SQ>(b, n, m, k, tau0, unknown, known, dist, idx, st);
    return e == cudaSuccess ? OGC_OK : static_cast<int>(e);
}

}  // namespace ogc

extern "C" int ogc_knn(int b, int n, int m, int k, const float *unknown, const float *known, float *dist2,
                       int *idx, void *stream) {
    return ogc::knn_dispatch<false>(b, n, m, k, __builtin_inff(), unknown, known, dist2, idx, stream);
}

extern "C" int ogc_knn_sqrt(int b, int n, int m, int k, const float *unknown, const float *known, float *dist,
                            int *idx, void *stream) {
    return ogc::knn_dispatch<true>(b, n, m, k, __builtin_inff(), unknown, known, dist, idx, stream);
}

extern "C" int ogc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                            int *idx, void *stream) {
    return ogc::knn_dispatch<false>(b, n, m, 3, __builtin_inff(), unknown, known, dist2, idx, stream);
}

// k-NN restricted to candidates with distance <= max_dist (sqrt'ed output like ogc_knn_sqrt).  For the call sites
// that discard farther neighbours anyway -- QueryAndGroup / KnnLoss replace every neighbour with dist > radius by
// the nearest one (pointnet2/pointnet2.py:284-286, losses/seg_loss_unsup.py:121-122) -- the result AFTER that
// clipping is identical to the exact k-NN, while far candidates never enter the running top-k (most insertions of
// the exact search in sparse outdoor clouds).  Slots with no candidate inside the bound hold (+inf, 0).
// Only valid when every query's nearest neighbour lies inside the bound (queries that are members of `known`).
extern "C" int ogc_knn_bounded(int b, int n, int m, int k, float max_dist, const float *unknown, const float *known,
                               float *dist, int *idx, void *stream) {
    if (!(max_dist >= 0.f)) return OGC_ERR_INVALID_ARG;
    // admit d2 < tau0 with tau0 just above max_dist^2, so every candidate with sqrtf(d2) <= max_dist is kept
    const float r2 = max_dist * max_dist;
    const float tau0 = r2 * 1.00001f + 1e-30f;
    return ogc::knn_dispatch<true>(b, n, m, k, tau0, unknown, known, dist, idx, stream);
}
