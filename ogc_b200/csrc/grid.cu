// Uniform-grid acceleration of the RADIUS-BOUNDED neighbourhood searches of the OGC losses and set-abstraction levels:
//   * bounded k-NN  (ogc_knn_bounded: KnnLoss k = 32, r = 1 on 8192 x 8192, QueryAndGroup k = 64;
//                    losses/seg_loss_unsup.py:112-129, pointnet2/pointnet2.py:283-286)
//   * ball query    (BallQLoss r = 2, 64 samples; losses/seg_loss_unsup.py:143-158, src/ball_query_gpu.cu:9-45)
// Both only ever use candidates within a radius, yet the brute-force kernels (knn.cu, ball_query.cu) evaluate all
// n x m pairs: 67 M pair evaluations per cloud for a neighbourhood that holds a few dozen points.
//
// Pre-pass (one CTA per cloud): bounding box -> a grid whose cells are at least `cell` (>= the search radius) wide and at
// most 32 x 8 x 32 = 8192 cells, counting sort of the cloud by cell -> (x, y, z, original index) records in cell order
// + the cell start table.  Query (one warp per query): the 3 x 3 x 3 cell neighbourhood is 9 CONTIGUOUS record ranges
// (cells adjacent in x are adjacent in memory); 32 candidates per step straight from L1 / L2 (queries are processed in
// cell order when the query cloud is the sorted cloud itself, so consecutive warps share their ranges).
//
// Results are BIT-IDENTICAL to the brute-force kernels (tests/test_gpu_grid.py):
//   * distances use the same rounding order (common.cuh sqdist);
//   * k-NN keeps the k smallest candidates under (distance, index) order with an explicit index tie-break (candidates
//     no longer arrive in index order);
//   * the ball query needs the FIRST nsample in-radius indices in ascending index order: in-radius candidates set bits of
//     a per-warp bitmap over the cloud, which is then read out in order.
#include "common.cuh"

namespace ogc {

constexpr int kGridMaxCells = 8192;
constexpr int kGridNx = 32, kGridNy = 8, kGridNz = 32;     // per-axis caps (product = kGridMaxCells)
constexpr int kGridBuildThreads = 1024;
constexpr int kGridParams = 16;                            // floats per cloud: origin[3], inv_cell[3], dims[3] (as int bits)

struct GridView {
    float ox, oy, oz, ix, iy, iz;
    int nx, ny, nz;
};

__device__ __forceinline__ GridView load_grid(const float *__restrict__ p) {
    GridView g;
    g.ox = p[0]; g.oy = p[1]; g.oz = p[2]; g.ix = p[3]; g.iy = p[4]; g.iz = p[5];
    g.nx = __float_as_int(p[6]); g.ny = __float_as_int(p[7]); g.nz = __float_as_int(p[8]);
    return g;
}

__device__ __forceinline__ int cell_coord(float x, float o, float inv, int n) {
    // clamp in float first: far-away / non-finite coordinates must not overflow the int conversion
    const float c = floorf((x - o) * inv);
    return static_cast<int>(fminf(fmaxf(c, -1.f), static_cast<float>(n)));   // -1 / n = outside the box on that side
}

// ---------------------------------------------------------------------------------------------- build
__global__ void __launch_bounds__(kGridBuildThreads)
grid_build_kernel(int m, float cell, const float *__restrict__ xyz, float4 *__restrict__ sorted,
                  int *__restrict__ cell_start, float *__restrict__ params) {
    __shared__ int hist[kGridMaxCells + 1];
    __shared__ float red[6][32];
    __shared__ float gp[kGridParams];
    __shared__ int warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    xyz += static_cast<size_t>(b) * m * 3;
    sorted += static_cast<size_t>(b) * m;
    cell_start += static_cast<size_t>(b) * (kGridMaxCells + 1);
    params += static_cast<size_t>(b) * kGridParams;

    // ---- bounding box (non-finite coordinates are ignored here and land in a border cell below) ----
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = tid; i < m; i += kGridBuildThreads) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(xyz + i * 3 + c);
            if (fabsf(v) < 1.0e30f) { lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(OGC_FULL_MASK, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(OGC_FULL_MASK, hi[c], o));
        }
        if (lane == 0) { red[c][warp] = lo[c]; red[3 + c][warp] = hi[c]; }
    }
    for (int i = tid; i <= kGridMaxCells; i += kGridBuildThreads) hist[i] = 0;
    __syncthreads();
    if (tid == 0) {
        const int caps[3] = {kGridNx, kGridNy, kGridNz};
        for (int c = 0; c < 3; ++c) {
            float l = red[c][0], h = red[3 + c][0];
            for (int w = 1; w < kGridBuildThreads / 32; ++w) { l = fminf(l, red[c][w]); h = fmaxf(h, red[3 + c][w]); }
            if (!(h >= l)) { l = 0.f; h = 0.f; }                   // empty / all non-finite
            const float ext = h - l;
            // cells at least `cell` wide (27-cell neighbourhoods cover the search radius), at most caps[c] per axis
            float w = fmaxf(cell, ext / static_cast<float>(caps[c]) * 1.0001f);
            if (!(w > 0.f)) w = 1.f;
            int n = static_cast<int>(floorf(ext / w)) + 1;
            n = n < 1 ? 1 : (n > caps[c] ? caps[c] : n);
            gp[c] = l;
            gp[3 + c] = 1.f / w;
            gp[6 + c] = __int_as_float(n);
        }
    }
    __syncthreads();
    if (tid < kGridParams) params[tid] = tid < 9 ? gp[tid] : 0.f;
    const GridView g = load_grid(gp);
    auto cell_of = [&](int i, float &x, float &y, float &z) {
        x = __ldg(xyz + i * 3); y = __ldg(xyz + i * 3 + 1); z = __ldg(xyz + i * 3 + 2);
        const int cx = min(max(cell_coord(x, g.ox, g.ix, g.nx), 0), g.nx - 1);
        const int cy = min(max(cell_coord(y, g.oy, g.iy, g.ny), 0), g.ny - 1);
        const int cz = min(max(cell_coord(z, g.oz, g.iz, g.nz), 0), g.nz - 1);
        return (cz * g.ny + cy) * g.nx + cx;
    };
    // ---- histogram, exclusive scan, scatter ----
    for (int i = tid; i < m; i += kGridBuildThreads) {
        float x, y, z;
        atomicAdd(&hist[cell_of(i, x, y, z)], 1);
    }
    __syncthreads();
    // block-wide exclusive scan of kGridMaxCells counters: 8 per thread
    {
        constexpr int PER = kGridMaxCells / kGridBuildThreads;
        int v[PER], s = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) { v[j] = hist[tid * PER + j]; s += v[j]; }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(OGC_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(OGC_FULL_MASK, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;                                // inclusive over warps
        }
        __syncthreads();
        int run = incl - s + (warp > 0 ? warp_tot[warp - 1] : 0);
#pragma unroll
        for (int j = 0; j < PER; ++j) { hist[tid * PER + j] = run; run += v[j]; }
        if (tid == kGridBuildThreads - 1) hist[kGridMaxCells] = run;
    }
    __syncthreads();
    for (int i = tid; i <= kGridMaxCells; i += kGridBuildThreads) cell_start[i] = hist[i];
    __syncthreads();
    for (int i = tid; i < m; i += kGridBuildThreads) {
        float x, y, z;
        const int c = cell_of(i, x, y, z);
        const int slot = atomicAdd(&hist[c], 1);               // order inside a cell is irrelevant (see the header)
        sorted[slot] = make_float4(x, y, z, __int_as_float(i));
    }
}

// ---------------------------------------------------------------------------------------------- shared by the queries
// The 9 contiguous record ranges of a query's 3 x 3 x 3 cell neighbourhood: calls f(begin, end) for each.
template <typename F>
__device__ __forceinline__ void for_each_range(const GridView &g, const int *__restrict__ cs, float qx, float qy, float qz, F f) {
    const int cx = cell_coord(qx, g.ox, g.ix, g.nx), cy = cell_coord(qy, g.oy, g.iy, g.ny), cz = cell_coord(qz, g.oz, g.iz, g.nz);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
    if (x0 > x1) return;
    for (int z = max(cz - 1, 0); z <= min(cz + 1, g.nz - 1); ++z)
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.ny - 1); ++y) {
            const int row = (z * g.ny + y) * g.nx;
            const int beg = __ldg(cs + row + x0), end = __ldg(cs + row + x1 + 1);
            if (end > beg) f(beg, end);
        }
}

constexpr int kGridQThreads = 256;
constexpr int kGridQWarps = kGridQThreads / 32;

// ---------------------------------------------------------------------------------------------- bounded k-NN
template <int KR>
struct TopKTie {           // sorted list of the k best (distance, index) pairs across the warp's registers, explicit index tie-break
    float d[KR];
    int i[KR];
    float tau;             // admission: cd < tau, or cd == tau and the list is not yet "full below tau0"
    float tau0;

    __device__ __forceinline__ void init(float bound) {
#pragma unroll
        for (int r = 0; r < KR; ++r) { d[r] = __int_as_float(0x7f800000); i[r] = 0x7fffffff; }
        tau0 = bound;
        tau = bound;
    }
    // warp-uniform: insert (cd, ci) at its rank under (distance, index) order; element k-1 falls off
    __device__ __forceinline__ void insert(float cd, int ci, int k, int lane) {
        int p = 0;
#pragma unroll
        for (int r = 0; r < KR; ++r) p += __popc(__ballot_sync(OGC_FULL_MASK, d[r] < cd || (d[r] == cd && i[r] < ci)));
        if (p >= k) return;
#pragma unroll
        for (int r = KR - 1; r >= 0; --r) {
            const int lo = 32 * r;
            if (p >= lo + 32) break;
            const float ud = __shfl_up_sync(OGC_FULL_MASK, d[r], 1);
            const int ui = __shfl_up_sync(OGC_FULL_MASK, i[r], 1);
            if (p >= lo) {
                const int lp = p - lo;
                if (lane > lp) { d[r] = ud; i[r] = ui; }
                else if (lane == lp) { d[r] = cd; i[r] = ci; }
            } else {
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

template <int KR>
__global__ void __launch_bounds__(kGridQThreads)
knn_grid_kernel(int n, int m, int k, float tau0, int self_query, const float *__restrict__ unknown,
                const float4 *__restrict__ sorted, const int *__restrict__ cell_start, const float *__restrict__ params,
                float *__restrict__ dist_out, int *__restrict__ idx_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bi = blockIdx.y;
    sorted += static_cast<size_t>(bi) * m;
    const int *cs = cell_start + static_cast<size_t>(bi) * (kGridMaxCells + 1);
    const GridView g = load_grid(params + static_cast<size_t>(bi) * kGridParams);
    const int w = blockIdx.x * kGridQWarps + warp;
    if (w >= n) return;
    int q = w;
    float qx, qy, qz;
    if (self_query) {                       // queries in cell order: neighbouring warps share their candidate ranges
        const float4 r = __ldg(sorted + w);
        qx = r.x; qy = r.y; qz = r.z; q = __float_as_int(r.w);
    } else {
        const float *u = unknown + (static_cast<size_t>(bi) * n + w) * 3;
        qx = __ldg(u); qy = __ldg(u + 1); qz = __ldg(u + 2);
    }
    TopKTie<KR> top;
    top.init(tau0);
    for_each_range(g, cs, qx, qy, qz, [&](int beg, int end) {
        for (int j0 = beg; j0 < end; j0 += 32) {
            const int j = j0 + lane;
            float cd = __int_as_float(0x7f800000);
            int ci = 0;
            if (j < end) {
                const float4 c = __ldg(sorted + j);
                cd = sqdist(qx, qy, qz, c.x, c.y, c.z);
                ci = __float_as_int(c.w);
            }
            // a candidate equal to the current k-th distance may still displace it through the index tie-break
            unsigned hits = __ballot_sync(OGC_FULL_MASK, cd < tau0 && cd <= top.tau);
            while (hits) {
                const int src = __ffs(hits) - 1;
                hits &= hits - 1;
                const float sd = __shfl_sync(OGC_FULL_MASK, cd, src);
                const int si = __shfl_sync(OGC_FULL_MASK, ci, src);
                if (sd <= top.tau) top.insert(sd, si, k, lane);
            }
        }
    });
    float *drow = dist_out + (static_cast<size_t>(bi) * n + q) * k;
    int *irow = idx_out + (static_cast<size_t>(bi) * n + q) * k;
#pragma unroll
    for (int r = 0; r < KR; ++r) {
        const int e = 32 * r + lane;
        if (e < k) {
            const bool filled = top.i[r] != 0x7fffffff;
            drow[e] = filled ? __fsqrt_rn(top.d[r]) : __int_as_float(0x7f800000);
            irow[e] = filled ? top.i[r] : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------- ball query
__global__ void __launch_bounds__(kGridQThreads)
ball_query_grid_kernel(int n, int m, float radius2, int nsample, int self_query, const float *__restrict__ new_xyz,
                       const float4 *__restrict__ sorted, const int *__restrict__ cell_start,
                       const float *__restrict__ params, int *__restrict__ idx) {
    extern __shared__ uint32_t bitmap_s[];           // [kGridQWarps][words]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bi = blockIdx.y;
    const int words = (m + 31) >> 5, wpad = (words + 31) & ~31;
    uint32_t *bm = bitmap_s + warp * wpad;
    sorted += static_cast<size_t>(bi) * m;
    const int *cs = cell_start + static_cast<size_t>(bi) * (kGridMaxCells + 1);
    const GridView g = load_grid(params + static_cast<size_t>(bi) * kGridParams);
    const int w = blockIdx.x * kGridQWarps + warp;
    if (w >= n) return;
    int q = w;
    float qx, qy, qz;
    if (self_query) {
        const float4 r = __ldg(sorted + w);
        qx = r.x; qy = r.y; qz = r.z; q = __float_as_int(r.w);
    } else {
        const float *u = new_xyz + (static_cast<size_t>(bi) * n + w) * 3;
        qx = __ldg(u); qy = __ldg(u + 1); qz = __ldg(u + 2);
    }
    for (int i = lane; i < wpad; i += 32) bm[i] = 0u;
    __syncwarp();
    for_each_range(g, cs, qx, qy, qz, [&](int beg, int end) {
        for (int j0 = beg; j0 < end; j0 += 32) {
            const int j = j0 + lane;
            if (j < end) {
                const float4 c = __ldg(sorted + j);
                if (sqdist(qx, qy, qz, c.x, c.y, c.z) < radius2) {
                    const int ci = __float_as_int(c.w);
                    atomicOr(&bm[ci >> 5], 1u << (ci & 31));
                }
            }
        }
    });
    __syncwarp();
    // read the bitmap out in ascending index order until nsample indices are found
    int *row = idx + (static_cast<size_t>(bi) * n + q) * nsample;
    int cnt = 0, first = 0;
    for (int w0 = 0; w0 < wpad && cnt < nsample; w0 += 32) {
        uint32_t bits = bm[w0 + lane];
        const int c = __popc(bits);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(OGC_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(OGC_FULL_MASK, incl, 31);
        if (total == 0) continue;
        if (cnt == 0) {
            const unsigned has = __ballot_sync(OGC_FULL_MASK, c > 0);
            const int fl = __ffs(has) - 1;
            first = __shfl_sync(OGC_FULL_MASK, (w0 + lane) * 32 + __ffs(bits) - 1, fl);
        }
        int slot = cnt + incl - c;
        while (bits && slot < nsample) {
            const int bpos = __ffs(bits) - 1;
            bits &= bits - 1;
            row[slot++] = (w0 + lane) * 32 + bpos;
        }
        cnt += total;
    }
    __syncwarp();
    const int filled = min(cnt, nsample);
    for (int s = filled + lane; s < nsample; s += 32) row[s] = first;      // first == 0 when there is no hit
}

static inline float grid_cell_for(float radius) { return radius * 1.001f + 1e-12f; }

}  // namespace ogc

// Workspace sizes (bytes) of ogc_grid_build for a batch of b clouds of m points.
extern "C" long long ogc_grid_sorted_bytes(int b, int m) { return static_cast<long long>(b) * m * 16; }
extern "C" long long ogc_grid_table_bytes(int b) { return static_cast<long long>(b) * (ogc::kGridMaxCells + 1) * 4; }
extern "C" long long ogc_grid_params_bytes(int b) { return static_cast<long long>(b) * ogc::kGridParams * 4; }

// Sort every cloud of `xyz` (b,m,3) into a uniform grid with cells >= radius (see the header of this file).
extern "C" int ogc_grid_build(int b, int m, float radius, const float *xyz, void *sorted, int *cell_start, float *params,
                              void *stream) {
    using namespace ogc;
    if (b < 0 || m < 0 || !(radius > 0.f)) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!sorted || !cell_start || !params || (m > 0 && !xyz)) return OGC_ERR_INVALID_ARG;
    grid_build_kernel<<<b, kGridBuildThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        m, grid_cell_for(radius), xyz, static_cast<float4 *>(sorted), cell_start, params);
    OGC_RETURN_LAUNCH_STATUS();
}

// ogc_knn_bounded on a grid built with radius >= max_dist.  unknown == NULL: the queries are the sorted cloud itself
// (n == m).  Same outputs as ogc_knn_bounded, bit for bit.
extern "C" int ogc_knn_grid(int b, int n, int m, int k, float max_dist, const float *unknown, const void *sorted,
                            const int *cell_start, const float *params, float *dist, int *idx, void *stream) {
    using namespace ogc;
    if (b < 0 || n < 0 || m < 0 || k < 1 || k > 224 || !(max_dist >= 0.f)) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n == 0) return OGC_OK;
    if (!sorted || !cell_start || !params || !dist || !idx) return OGC_ERR_INVALID_ARG;
    if (!unknown && n != m) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    const float r2 = max_dist * max_dist;
    const float tau0 = r2 * 1.00001f + 1e-30f;               // the admission bound of ogc_knn_bounded
    dim3 grid((n + kGridQWarps - 1) / kGridQWarps, b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float4 *s4 = static_cast<const float4 *>(sorted);
    const int self = unknown == nullptr;
    if (k <= 32) knn_grid_kernel<1><<<grid, kGridQThreads, 0, st>>>(n, m, k, tau0, self, unknown, s4, cell_start, params, dist, idx);
    else if (k <= 64) knn_grid_kernel<2><<<grid, kGridQThreads, 0, st>>>(n, m, k, tau0, self, unknown, s4, cell_start, params, dist, idx);
    else if (k <= 128) knn_grid_kernel<4><<<grid, kGridQThreads, 0, st>>>(n, m, k, tau0, self, unknown, s4, cell_start, params, dist, idx);
    else knn_grid_kernel<7><<<grid, kGridQThreads, 0, st>>>(n, m, k, tau0, self, unknown, s4, cell_start, params, dist, idx);
    OGC_RETURN_LAUNCH_STATUS();
}

// ogc_ball_query on a grid built with radius >= `radius`.  new_xyz == NULL: the centres are the sorted cloud itself.
// m <= 32768 (per-warp bitmap in shared memory).  Same output as ogc_ball_query, bit for bit.
extern "C" int ogc_ball_query_grid(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                                   const void *sorted, const int *cell_start, const float *params, int *idx,
                                   void *stream) {
    using namespace ogc;
    if (b < 0 || n < 0 || m < 0 || nsample < 0) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n == 0 || nsample == 0) return OGC_OK;
    if (!sorted || !cell_start || !params || !idx) return OGC_ERR_INVALID_ARG;
    if (!new_xyz && n != m) return OGC_ERR_INVALID_ARG;
    if (b > 65535 || m > 32768) return OGC_ERR_UNSUPPORTED;
    const int words = (m + 31) >> 5, wpad = (words + 31) & ~31;
    const size_t smem = static_cast<size_t>(kGridQWarps) * wpad * 4;
    dim3 grid((n + kGridQWarps - 1) / kGridQWarps, b);
    const float radius2 = radius * radius;                   // fp32 product, as src/ball_query_gpu.cu:23
    ball_query_grid_kernel<<<grid, kGridQThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        n, m, radius2, nsample, new_xyz == nullptr, new_xyz, static_cast<const float4 *>(sorted), cell_start, params, idx);
    OGC_RETURN_LAUNCH_STATUS();
}
