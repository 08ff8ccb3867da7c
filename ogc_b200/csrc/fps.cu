// K1  furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel<block_size> of the reference
// (pointnet2/src/sampling_gpu.cu:93-209).  Same algorithm -- M sequential "update running
// minimum distance, pick the arg-max" steps per cloud -- but:
//   * the cloud and its running minimum live in REGISTERS (reference: re-read from global and
//     the running minimum read+written in global memory every iteration, :130-135);
//   * the block arg-max is two redux.sync per warp + one __syncthreads + two redux.sync
//     (reference: a 10-level shared-memory tree with 11 __syncthreads, :141-203);
//   * the winner's coordinates come from a shared-memory copy staged once by the TMA
//     bulk-copy engine.
//
// Bit-exact index parity.  The reference resolves ties between equal distances through the
// shape of its reduction: thread tid scans k = tid, tid+bs, ... keeping the FIRST maximum
// (strict '>', :136-137), and the tree step __update(tid, tid+half) keeps the LOWER slot on
// equality (:86-91).  The last tree level compares even against odd tids, the one before
// tid%4==0 against tid%4==2, ... so among equal maxima the winner is the thread with the
// smallest BIT-REVERSED tid (over log2(bs) bits), and within it the smallest k.  We keep the
// reference's thread->point mapping (bs = 2^floor(log2 n) <= 1024 threads, k = tid + s*bs) and
// reduce the key (distance bits, bitrev(tid), s) so the same point wins.
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace ogc {

constexpr int kFpsMaxThreads = 1024;
constexpr int kFpsRegPoints = 16384;  // largest n handled by the register-resident kernel

// cuda_utils.h:10-14 opt_n_threads(), same double-precision expression.
static int fps_block_size(int n) {
    const int pow_2 = static_cast<int>(std::log(static_cast<double>(n)) / std::log(2.0));
    int t = 1 << pow_2;
    if (t > kFpsMaxThreads) t = kFpsMaxThreads;
    if (t < 1) t = 1;
    return t;
}

// Block-wide arg-max of (value bits, key).  Larger value wins; among equal values the smaller key.
// Exactly one barrier; `rec` is double-buffered by iteration parity by the caller.
__device__ __forceinline__ uint32_t block_argmax_key(uint32_t vbits, uint32_t key, uint2 *rec, int warp, int lane,
                                                     int nwarps) {
    const uint32_t wmax = __reduce_max_sync(OGC_FULL_MASK, vbits);
    const uint32_t cand = (vbits == wmax) ? key : 0xffffffffu;
    const uint32_t wkey = __reduce_min_sync(OGC_FULL_MASK, cand);
    if (lane == 0) rec[warp] = make_uint2(wmax, wkey);
    __syncthreads();
    uint2 r = make_uint2(0u, 0xffffffffu);
    if (lane < nwarps) r = rec[lane];
    const uint32_t bmax = __reduce_max_sync(OGC_FULL_MASK, r.x);
    const uint32_t c2 = (r.x == bmax) ? r.y : 0xffffffffu;
    return __reduce_min_sync(OGC_FULL_MASK, c2);
}

// One CTA per cloud, blockDim.x = max(bs, 32).  Thread tid < bs owns points k = tid + s*bs, s < PPT.
// XYZ_IN_REGS: coordinates in registers (PPT <= 8), else re-read from the shared-memory copy.
template <int PPT, bool XYZ_IN_REGS>
__global__ void __launch_bounds__(kFpsMaxThreads, 1)
fps_reg_kernel(int n, int m, int bs, int log2bs, const float *__restrict__ dataset, int *__restrict__ idxs) {
    extern __shared__ __align__(16) float fps_smem[];
    __shared__ uint2 rec[2][32];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    dataset += static_cast<size_t>(blockIdx.x) * n * 3;
    idxs += static_cast<size_t>(blockIdx.x) * m;

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    const float *pts = stage_floats(fps_smem, dataset, n * 3, &bar, parity);
    __syncthreads();

    const bool active = tid < bs;
    float px[XYZ_IN_REGS ? PPT : 1], py[XYZ_IN_REGS ? PPT : 1], pz[XYZ_IN_REGS ? PPT : 1];
    float mind[PPT];
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        const int k = tid + s * bs;
        const bool valid = active && k < n;
        if (XYZ_IN_REGS) {
            px[s] = valid ? pts[k * 3 + 0] : 0.f;
            py[s] = valid ? pts[k * 3 + 1] : 0.f;
            pz[s] = valid ? pts[k * 3 + 2] : 0.f;
        }
        mind[s] = valid ? 1e10f : -1.0f;  // -1 never beats the reference's initial best = -1 (:126)
    }
    // tie key: bit-reversed tid in the high bits, slot s in the low 5 bits (PPT <= 16)
    // (bs == 1 has log2bs == 0: no reversal, and a shift by 32 would be undefined)
    const uint32_t prio =
        (active && log2bs > 0) ? ((__brev(static_cast<uint32_t>(tid)) >> (32 - log2bs)) << 5) : 0u;

    int old = 0;
    if (tid == 0) idxs[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
        float best = -1.0f;
        int bests = 0;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            float x2, y2, z2;
            if (XYZ_IN_REGS) {
                x2 = px[s]; y2 = py[s]; z2 = pz[s];
            } else {
                const int k = min(tid + s * bs, n - 1);
                x2 = pts[k * 3 + 0]; y2 = pts[k * 3 + 1]; z2 = pts[k * 3 + 2];
            }
            const float d = sqdist(x2, y2, z2, x1, y1, z1);
            const float d2 = fminf(d, mind[s]);
            mind[s] = d2;
            if (d2 > best) { best = d2; bests = s; }
        }
        // best >= 0 for every active thread (slot 0 is always valid because bs <= n)
        const uint32_t vbits = active ? __float_as_uint(best) : 0u;
        const uint32_t key = active ? (prio | static_cast<uint32_t>(bests)) : 0xffffffffu;
        const uint32_t win = block_argmax_key(vbits, key, rec[j & 1], warp, lane, nwarps);
        const uint32_t wtid = log2bs == 0 ? 0u : (__brev(win >> 5) >> (32 - log2bs));
        old = static_cast<int>(wtid) + static_cast<int>(win & 31u) * bs;
        if (tid == 0) idxs[j] = old;
    }
}

// ---- raw scenes (16384 < n <= 131072, B small): one thread-block CLUSTER per cloud ------------------------------------
// The single-CTA fallback below streams the whole cloud and its running minimum through L2 every iteration
// (n / 1024 points per thread).  Here CL CTAs (8, or 16 with the non-portable size) split the cloud: thread tid of
// cluster rank r owns points k = tid + 1024 * (r + CL * s), s < PPT <= 8, with coordinates AND running minimum in
// registers.  Per iteration: local update + block arg-max as in fps_reg_kernel, then the winning thread of every CTA
// writes its candidate (distance, tie key, xyz) into a slot of EVERY CTA's shared memory (distributed shared memory,
// st.shared::cluster), one cluster barrier, and every CTA picks the winner of the CL candidates locally -- so the next
// centre's coordinates arrive with the candidate and nothing is re-read from global memory.
// Tie rule (pure function of the point index, hence independent of the partition): largest distance, then smallest
// bit-reversed (k mod 1024), then smallest k -- the reference's 1024-thread tree (see the header of this file).
struct __align__(16) FpsCand {
    uint32_t vbits, key;
    float x, y, z;
    uint32_t pad[3];
};

__device__ __forceinline__ uint2 block_argmax_pair(uint32_t vbits, uint32_t key, uint2 *rec, int warp, int lane) {
    const uint32_t wmax = __reduce_max_sync(OGC_FULL_MASK, vbits);
    const uint32_t cand = (vbits == wmax) ? key : 0xffffffffu;
    const uint32_t wkey = __reduce_min_sync(OGC_FULL_MASK, cand);
    if (lane == 0) rec[warp] = make_uint2(wmax, wkey);
    __syncthreads();
    const uint2 r = rec[lane];                       // 32 warps
    const uint32_t bmax = __reduce_max_sync(OGC_FULL_MASK, r.x);
    const uint32_t c2 = (r.x == bmax) ? r.y : 0xffffffffu;
    return make_uint2(bmax, __reduce_min_sync(OGC_FULL_MASK, c2));
}

template <int CL, int PPT>
__global__ void __launch_bounds__(kFpsMaxThreads, 1)
fps_cluster_kernel(int n, int m, const float *__restrict__ dataset, int *__restrict__ idxs) {
    __shared__ uint2 rec[2][32];
    __shared__ FpsCand cand[2][CL];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int cloud = blockIdx.x / CL;
    dataset += static_cast<size_t>(cloud) * n * 3;
    idxs += static_cast<size_t>(cloud) * m;

    float px[PPT], py[PPT], pz[PPT], mind[PPT];
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        const int k = tid + kFpsMaxThreads * (static_cast<int>(rank) + CL * s);
        const bool valid = k < n;
        px[s] = valid ? __ldg(dataset + static_cast<size_t>(k) * 3 + 0) : 0.f;
        py[s] = valid ? __ldg(dataset + static_cast<size_t>(k) * 3 + 1) : 0.f;
        pz[s] = valid ? __ldg(dataset + static_cast<size_t>(k) * 3 + 2) : 0.f;
        mind[s] = valid ? 1e10f : -1.0f;
    }
    const uint32_t prio = (__brev(static_cast<uint32_t>(tid)) >> 22) << 8;     // 10-bit reversal above the chunk bits
    float ox = __ldg(dataset), oy = __ldg(dataset + 1), oz = __ldg(dataset + 2);
    if (rank == 0 && tid == 0) idxs[0] = 0;

    for (int j = 1; j < m; ++j) {
        float best = -1.0f;
        int bests = 0;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const float d = sqdist(px[s], py[s], pz[s], ox, oy, oz);
            const float d2 = fminf(d, mind[s]);
            mind[s] = d2;
            if (d2 > best) { best = d2; bests = s; }
        }
        const bool any = best >= 0.f;                       // false for threads that own no valid point
        const uint32_t vbits = any ? __float_as_uint(best) : 0u;
        const uint32_t key = any ? (prio | static_cast<uint32_t>(rank + CL * bests)) : 0xffffffffu;
        const uint2 win = block_argmax_pair(vbits, key, rec[j & 1], warp, lane);
        if (key == win.y && any) {
            // this CTA's candidate -> slot `rank` of every CTA of the cluster
            const uint32_t local = smem_u32(&cand[j & 1][rank]);
#pragma unroll
            for (int r = 0; r < CL; ++r) {
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
                asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "r"(win.x), "r"(win.y),
                             "r"(__float_as_uint(px[bests])), "r"(__float_as_uint(py[bests]))
                             : "memory");
                asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(remote + 16), "r"(__float_as_uint(pz[bests])) : "memory");
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        // every CTA picks the cluster-wide winner from its own copy of the CL candidates
        uint32_t bv = 0u, bk = 0xffffffffu;
        float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
        for (int r = 0; r < CL; ++r) {
            const FpsCand c = cand[j & 1][r];
            if (c.vbits > bv || (c.vbits == bv && c.key < bk)) { bv = c.vbits; bk = c.key; bx = c.x; by = c.y; bz = c.z; }
        }
        ox = bx; oy = by; oz = bz;
        if (rank == 0 && tid == 0)
            idxs[j] = static_cast<int>(__brev(bk >> 8) >> 22) + kFpsMaxThreads * static_cast<int>(bk & 255u);
    }
    // no CTA may exit while a peer can still address its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CL, int PPT>
static cudaError_t launch_cluster(int b, int n, int m, const float *dataset, int *idxs, cudaStream_t st) {
    auto kern = fps_cluster_kernel<CL, PPT>;
    if (CL > 8) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(b) * CL);
    cfg.blockDim = dim3(kFpsMaxThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, n, m, dataset, idxs);
}

// Fallback for n > kFpsRegPoints: running minimum in the caller's `temp` scratch (global / L2),
// coordinates re-read through L1/L2.  Same tie rule; bs = 1024 here, slot index can exceed 31 so the
// key carries the full point index instead.
__global__ void __launch_bounds__(kFpsMaxThreads, 1)
fps_large_kernel(int n, int m, const float *__restrict__ dataset, float *__restrict__ temp,
                 int *__restrict__ idxs) {
    __shared__ uint2 rec[2][32];
    __shared__ uint32_t win_k[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int bs = kFpsMaxThreads;
    dataset += static_cast<size_t>(blockIdx.x) * n * 3;
    temp += static_cast<size_t>(blockIdx.x) * n;
    idxs += static_cast<size_t>(blockIdx.x) * m;
    const uint32_t prio = __brev(static_cast<uint32_t>(tid)) >> 22;  // 10-bit reversal

    int old = 0;
    if (tid == 0) idxs[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = __ldg(dataset + old * 3 + 0), y1 = __ldg(dataset + old * 3 + 1),
                    z1 = __ldg(dataset + old * 3 + 2);
        float best = -1.0f;
        int besti = 0;
        for (int k = tid; k < n; k += bs) {
            const float d = sqdist(__ldg(dataset + k * 3 + 0), __ldg(dataset + k * 3 + 1),
                                   __ldg(dataset + k * 3 + 2), x1, y1, z1);
            const float d2 = fminf(d, j == 1 ? 1e10f : temp[k]);
            temp[k] = d2;
            if (d2 > best) { best = d2; besti = k; }
        }
        const uint32_t win = block_argmax_key(__float_as_uint(best), prio, rec[j & 1], warp, lane, 32);
        // the unique thread whose prio won publishes its point index
        if (prio == win) win_k[j & 1] = static_cast<uint32_t>(besti);
        __syncthreads();
        old = static_cast<int>(win_k[j & 1]);
        if (tid == 0) idxs[j] = old;
    }
}

template <int PPT, bool R>
static cudaError_t launch_reg(int b, int n, int m, int bs, int log2bs, const float *dataset, int *idxs,
                              cudaStream_t st) {
    const size_t smem = (static_cast<size_t>(n) * 3 + 4) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(fps_reg_kernel<PPT, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const int threads = bs < 32 ? 32 : bs;
    fps_reg_kernel<PPT, R><<<b, threads, smem, st>>>(n, m, bs, log2bs, dataset, idxs);
    return cudaGetLastError();
}

}  // namespace ogc

extern "C" int ogc_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp, int *idxs,
                                           void *stream) {
    using namespace ogc;
    if (b < 0 || n <= 0 || m < 0 || !dataset || !idxs) return OGC_ERR_INVALID_ARG;
    if (b == 0 || m == 0) return OGC_OK;  // reference: kernel returns immediately for m <= 0 (:101)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int bs = fps_block_size(n);
    int log2bs = 0;
    while ((1 << log2bs) < bs) ++log2bs;
    cudaError_t e;
    if (n <= kFpsRegPoints) {
        const int ppt = (n + bs - 1) / bs;
        if (ppt <= 1) e = launch_reg<1, true>(b, n, m, bs, log2bs, dataset, idxs, st);
        else if (ppt <= 2) e = launch_reg<2, true>(b, n, m, bs, log2bs, dataset, idxs, st);
        else if (ppt <= 4) e = launch_reg<4, true>(b, n, m, bs, log2bs, dataset, idxs, st);
        else if (ppt <= 8) e = launch_reg<8, true>(b, n, m, bs, log2bs, dataset, idxs, st);
        else e = launch_reg<16, false>(b, n, m, bs, log2bs, dataset, idxs, st);
    } else {
        // raw scenes: a cluster of 8 (<= 65536 points) or 16 (<= 131072) CTAs per cloud; the single-CTA kernel remains
        // the fallback for larger clouds and for devices / partitions that cannot place the cluster
        e = cudaErrorInvalidValue;
        const int chunks = getenv("OGC_FPS_NO_CLUSTER") ? (1 << 30) : (n + kFpsMaxThreads - 1) / kFpsMaxThreads;   // env: measurement only
        if (chunks <= 8 * 4) e = launch_cluster<8, 4>(b, n, m, dataset, idxs, st);
        else if (chunks <= 8 * 8) e = launch_cluster<8, 8>(b, n, m, dataset, idxs, st);
        else if (chunks <= 16 * 8) e = launch_cluster<16, 8>(b, n, m, dataset, idxs, st);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            if (!temp) return OGC_ERR_WORKSPACE;
            fps_large_kernel<<<b, kFpsMaxThreads, 0, st>>>(n, m, dataset, temp, idxs);
            e = cudaGetLastError();
        }
    }
    return e == cudaSuccess ? OGC_OK : static_cast<int>(e);
}
