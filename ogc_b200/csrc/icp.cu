// Object-aware ICP correspondence step for sm_100a.
//
// Replaces, per ICP iteration, the N x N pipeline of object_aware_icp (oa_icp.py:64-75):
//     dist12 = -cdist(pc1 + flow, pc2) / T ;  corr = softmax(dist12, -1) ;  corr *= consistency12 ;
//     corr /= clamp(rowsum(corr), 1e-10) ;  flow = corr @ pc2 - pc1
// with consistency12 = mask1 mask2^T (oa_icp.py:57), which the reference materialises as several (B,N,N) fp32
// tensors (268 MB each per cloud at N = 8192).  Here nothing N x N exists: one warp per source point streams the
// target cloud (xyz + mask rows staged in shared memory tiles), keeps an online-softmax state
// (running max, Z = sum e, A = sum e c, V = sum e c p2) per lane and merges the 32 lanes at the end:
//     flow = (V / Z) / max(A / Z, 1e-10) - p1
// Distances are true differences in fp32 (the reference's cdist uses the |a|^2+|b|^2-2ab form whose cancellation
// error the temperature amplifies: SURVEY.md 7, hard part 8) -- parity is judged against the reference evaluated
// in float64.
// Two forms of each kernel: thread-per-query (round 2: warp-uniform target data, packed FFMA2 over target pairs, two
// passes; 2x the round-1 rate when there are enough queries to fill the GPU) and warp-per-query (round 1: online softmax,
// 32 lanes split the targets; used for small problems).  Measured, 8192 points, 20 iterations: 1 / 4 / 8 / 16 / 64 clouds =
// 5.8 / 12.1 / 19.7 / 36.1 / 98.8 ms.
#include "common.cuh"

namespace ogc {

constexpr int kIcpTile = 512;      // target points per shared-memory tile
constexpr int kIcpMaxK = 16;

__device__ __forceinline__ float sqrt_approx(float x) {      // MUFU.SQRT: <= 2 ulp, one instruction instead of the ~10 of the IEEE sequence
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {       // MUFU.EX2
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}


__device__ __forceinline__ float2 ld_smem_f2(const float *p) { return *reinterpret_cast<const float2 *>(p); }

// Round-2 form: THREAD = source point.  Every datum of a target (coordinates, K mask entries) is then warp-uniform -- one
// broadcast shared-memory read serves 32 pairs -- and the per-pair work is pure arithmetic on registers.  Targets are taken
// two at a time on packed fp32 FMAs (a 3-register FFMA issues every other cycle, FFMA2 retires two per slot): the tile
// is stored structure-of-arrays so that (x_t, x_t+1) and (m2_t[k], m2_t+1[k]) are adjacent.  TWO passes over the target
// cloud instead of an online softmax: pass 1 finds the nearest target (the softmax maximum is -d_min / T), pass 2
// accumulates with that fixed maximum -- no rescaling branch, no lane merge.  exp(x) = 2^(x log2 e), log2 e folded into
// 1/T; MUFU sqrt / ex2.  Two source points per thread share the broadcast reads.  29 -> ~18 instructions per pair.

// THREADS x Q source points per CTA: (256, 2) when the grid still fills the GPU twice over, else (256, 1), else (64, 1)
// (object_aware_icp on 4 clouds, mask_voting on 8 frames: few, small clouds).
template <int K, int kIcpCThreads, int kIcpQ>
__global__ void __launch_bounds__(kIcpCThreads)
icp_correspond_kernel(int n1, int n2, float inv_temp, const float *__restrict__ pc1, const float *__restrict__ flow,
                      const float *__restrict__ pc2, const float *__restrict__ mask1, const float *__restrict__ mask2,
                      float *__restrict__ flow_out) {
    __shared__ __align__(16) float sx[kIcpTile], sy[kIcpTile], sz[kIcpTile];
    __shared__ __align__(16) float sm[K * kIcpTile];            // [k][target]
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * (kIcpCThreads * kIcpQ) + threadIdx.x;      // queries q0, q0 + 256: coalesced loads / stores
    pc2 += static_cast<size_t>(b) * n2 * 3;
    mask2 += static_cast<size_t>(b) * n2 * K;
    const float scale2 = -inv_temp * 1.4426950408889634f;       // s2 = -d / T * log2(e)

    float px[kIcpQ], py[kIcpQ], pz[kIcpQ];
    float2 qx[kIcpQ], qy[kIcpQ], qz[kIcpQ], m1[kIcpQ][K];
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) {
        const int q = q0 + u * kIcpCThreads;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        px[u] = py[u] = pz[u] = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) m1[u][k] = make_float2(0.f, 0.f);
        if (q < n1) {
            const size_t o = (static_cast<size_t>(b) * n1 + q) * 3;
            px[u] = __ldg(pc1 + o); py[u] = __ldg(pc1 + o + 1); pz[u] = __ldg(pc1 + o + 2);
            fx = px[u] + __ldg(flow + o); fy = py[u] + __ldg(flow + o + 1); fz = pz[u] + __ldg(flow + o + 2);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float v = __ldg(mask1 + (static_cast<size_t>(b) * n1 + q) * K + k);
                m1[u][k] = make_float2(v, v);
            }
        }
        qx[u] = make_float2(fx, fx); qy[u] = make_float2(fy, fy); qz[u] = make_float2(fz, fz);
    }
    const float2 mone2 = make_float2(-1.f, -1.f), one2 = make_float2(1.f, 1.f), zero2 = make_float2(0.f, 0.f);
    auto load_xyz = [&](int t0, int tn) {
        for (int i = threadIdx.x; i < kIcpTile; i += kIcpCThreads) {
            const bool in = i < tn;                  // padding targets sit at +inf distance: weight 0, never the nearest
            sx[i] = in ? __ldg(pc2 + static_cast<size_t>(t0 + i) * 3) : 1e18f;
            sy[i] = in ? __ldg(pc2 + static_cast<size_t>(t0 + i) * 3 + 1) : 1e18f;
            sz[i] = in ? __ldg(pc2 + static_cast<size_t>(t0 + i) * 3 + 2) : 1e18f;
        }
    };
    // ---- pass 1: squared distance to the nearest target ----
    float2 dmin[kIcpQ];
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) dmin[u] = make_float2(INFINITY, INFINITY);
    for (int t0 = 0; t0 < n2; t0 += kIcpTile) {
        const int tn = min(kIcpTile, n2 - t0);
        __syncthreads();
        load_xyz(t0, tn);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < kIcpTile; j += 2) {
            const float2 x2 = ld_smem_f2(sx + j), y2 = ld_smem_f2(sy + j), z2 = ld_smem_f2(sz + j);
#pragma unroll
            for (int u = 0; u < kIcpQ; ++u) {
                const float2 dx = ffma2(mone2, x2, qx[u]), dy = ffma2(mone2, y2, qy[u]), dz = ffma2(mone2, z2, qz[u]);
                float2 d2 = ffma2(dx, dx, zero2);
                d2 = ffma2(dy, dy, d2);
                d2 = ffma2(dz, dz, d2);
                dmin[u].x = fminf(dmin[u].x, d2.x);
                dmin[u].y = fminf(dmin[u].y, d2.y);
            }
        }
    }
    // - (softmax maximum in base-2 units), from the SAME arithmetic pass 2 uses for every pair
    float2 nmx[kIcpQ];
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) {
        const float nm = -(sqrt_approx(fminf(dmin[u].x, dmin[u].y)) * scale2);
        nmx[u] = make_float2(nm, nm);
    }
    const float2 sc2 = make_float2(scale2, scale2);
    // ---- pass 2: softmax-weighted sums with the maximum fixed ----
    float2 Z[kIcpQ], A[kIcpQ], vx[kIcpQ], vy[kIcpQ], vz[kIcpQ];
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) Z[u] = A[u] = vx[u] = vy[u] = vz[u] = zero2;
    for (int t0 = 0; t0 < n2; t0 += kIcpTile) {
        const int tn = min(kIcpTile, n2 - t0);
        __syncthreads();
        load_xyz(t0, tn);
        for (int i = threadIdx.x; i < kIcpTile * K; i += kIcpCThreads) {
            const int t = i / K, k = i - t * K;
            sm[k * kIcpTile + t] = t < tn ? __ldg(mask2 + static_cast<size_t>(t0) * K + i) : 0.f;
        }
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < kIcpTile; j += 2) {
            const float2 x2 = ld_smem_f2(sx + j), y2 = ld_smem_f2(sy + j), z2 = ld_smem_f2(sz + j);
            float2 m2[K];
#pragma unroll
            for (int k = 0; k < K; ++k) m2[k] = ld_smem_f2(sm + k * kIcpTile + j);
#pragma unroll
            for (int u = 0; u < kIcpQ; ++u) {
                const float2 dx = ffma2(mone2, x2, qx[u]), dy = ffma2(mone2, y2, qy[u]), dz = ffma2(mone2, z2, qz[u]);
                float2 d2 = ffma2(dx, dx, zero2);
                d2 = ffma2(dy, dy, d2);
                d2 = ffma2(dz, dz, d2);
                const float2 arg = ffma2(make_float2(sqrt_approx(d2.x), sqrt_approx(d2.y)), sc2, nmx[u]);     // s2 - max <= 0
                float2 c = zero2;
#pragma unroll
                for (int k = 0; k < K; ++k) c = ffma2(m1[u][k], m2[k], c);
                const float2 e = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
                const float2 ec = ffma2(e, c, zero2);
                Z[u] = ffma2(e, one2, Z[u]);
                A[u] = ffma2(ec, one2, A[u]);
                vx[u] = ffma2(ec, x2, vx[u]); vy[u] = ffma2(ec, y2, vy[u]); vz[u] = ffma2(ec, z2, vz[u]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) {
        const int q = q0 + u * kIcpCThreads;
        if (q >= n1) continue;
        const float invZ = 1.0f / (Z[u].x + Z[u].y);
        const float rs = fmaxf((A[u].x + A[u].y) * invZ, 1e-10f);
        const size_t o = (static_cast<size_t>(b) * n1 + q) * 3;
        flow_out[o] = ((vx[u].x + vx[u].y) * invZ) / rs - px[u];
        flow_out[o + 1] = ((vy[u].x + vy[u].y) * invZ) / rs - py[u];
        flow_out[o + 2] = ((vz[u].x + vz[u].y) * invZ) / rs - pz[u];
    }
}


// ---- soft correspondence transfer (vote.py) ---------------------------------------------------------------
// out[m,:] = sum_n softmax_n(-|q_m - key_n| / T) * val[n,:]      q (b,n1,3), key (b,n2,3), val (b,n2,K) -> out (b,n1,K)
//
// Replaces corr = softmax(-cdist(pc1 + flow, pc2) / T) (vote.py:17-28) FOLLOWED BY its only use, corr @ mask
// (vote.py:121), and -- because rows of a product of row-stochastic matrices still sum to one -- the chained
// propagation corr(t,t+2) = normalise(corr(t,t+1) @ corr(t+1,t+2)) (vote.py:50-57) as repeated application:
// corr(t,t+2) @ M = corr(t,t+1) @ (corr(t+1,t+2) @ M).  No N x N tensor and no N^3 bmm ever exists.
// Same form as icp_correspond_kernel (thread = query, two keys per packed FFMA2, two passes, 2 queries per thread).
template <int K, int kIcpCThreads, int kIcpQ>
__global__ void __launch_bounds__(kIcpCThreads)
softmax_transfer_kernel(int n1, int n2, float inv_temp, const float *__restrict__ query, const float *__restrict__ key,
                        const float *__restrict__ val, float *__restrict__ out) {
    __shared__ __align__(16) float sx[kIcpTile], sy[kIcpTile], sz[kIcpTile];
    __shared__ __align__(16) float sv[K * kIcpTile];            // [k][key]
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * (kIcpCThreads * kIcpQ) + threadIdx.x;
    key += static_cast<size_t>(b) * n2 * 3;
    val += static_cast<size_t>(b) * n2 * K;
    const float scale2 = -inv_temp * 1.4426950408889634f;

    float2 qx[kIcpQ], qy[kIcpQ], qz[kIcpQ];
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) {
        const int q = q0 + u * kIcpCThreads;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        if (q < n1) {
            const size_t o = (static_cast<size_t>(b) * n1 + q) * 3;
            fx = __ldg(query + o); fy = __ldg(query + o + 1); fz = __ldg(query + o + 2);
        }
        qx[u] = make_float2(fx, fx); qy[u] = make_float2(fy, fy); qz[u] = make_float2(fz, fz);
    }
    const float2 mone2 = make_float2(-1.f, -1.f), one2 = make_float2(1.f, 1.f), zero2 = make_float2(0.f, 0.f);
    auto load_xyz = [&](int t0, int tn) {
        for (int i = threadIdx.x; i < kIcpTile; i += kIcpCThreads) {
            const bool in = i < tn;
            sx[i] = in ? __ldg(key + static_cast<size_t>(t0 + i) * 3) : 1e18f;
            sy[i] = in ? __ldg(key + static_cast<size_t>(t0 + i) * 3 + 1) : 1e18f;
            sz[i] = in ? __ldg(key + static_cast<size_t>(t0 + i) * 3 + 2) : 1e18f;
        }
    };
    float2 dmin[kIcpQ];
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) dmin[u] = make_float2(INFINITY, INFINITY);
    for (int t0 = 0; t0 < n2; t0 += kIcpTile) {
        const int tn = min(kIcpTile, n2 - t0);
        __syncthreads();
        load_xyz(t0, tn);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < kIcpTile; j += 2) {
            const float2 x2 = ld_smem_f2(sx + j), y2 = ld_smem_f2(sy + j), z2 = ld_smem_f2(sz + j);
#pragma unroll
            for (int u = 0; u < kIcpQ; ++u) {
                const float2 dx = ffma2(mone2, x2, qx[u]), dy = ffma2(mone2, y2, qy[u]), dz = ffma2(mone2, z2, qz[u]);
                float2 d2 = ffma2(dx, dx, zero2);
                d2 = ffma2(dy, dy, d2);
                d2 = ffma2(dz, dz, d2);
                dmin[u].x = fminf(dmin[u].x, d2.x);
                dmin[u].y = fminf(dmin[u].y, d2.y);
            }
        }
    }
    float2 nmx[kIcpQ];
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) {
        const float nm = -(sqrt_approx(fminf(dmin[u].x, dmin[u].y)) * scale2);
        nmx[u] = make_float2(nm, nm);
    }
    const float2 sc2 = make_float2(scale2, scale2);
    float2 Z[kIcpQ], acc[kIcpQ][K];
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) {
        Z[u] = zero2;
#pragma unroll
        for (int k = 0; k < K; ++k) acc[u][k] = zero2;
    }
    for (int t0 = 0; t0 < n2; t0 += kIcpTile) {
        const int tn = min(kIcpTile, n2 - t0);
        __syncthreads();
        load_xyz(t0, tn);
        for (int i = threadIdx.x; i < kIcpTile * K; i += kIcpCThreads) {
            const int t = i / K, k = i - t * K;
            sv[k * kIcpTile + t] = t < tn ? __ldg(val + static_cast<size_t>(t0) * K + i) : 0.f;
        }
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < kIcpTile; j += 2) {
            const float2 x2 = ld_smem_f2(sx + j), y2 = ld_smem_f2(sy + j), z2 = ld_smem_f2(sz + j);
            float2 v2[K];
#pragma unroll
            for (int k = 0; k < K; ++k) v2[k] = ld_smem_f2(sv + k * kIcpTile + j);
#pragma unroll
            for (int u = 0; u < kIcpQ; ++u) {
                const float2 dx = ffma2(mone2, x2, qx[u]), dy = ffma2(mone2, y2, qy[u]), dz = ffma2(mone2, z2, qz[u]);
                float2 d2 = ffma2(dx, dx, zero2);
                d2 = ffma2(dy, dy, d2);
                d2 = ffma2(dz, dz, d2);
                const float2 arg = ffma2(make_float2(sqrt_approx(d2.x), sqrt_approx(d2.y)), sc2, nmx[u]);
                const float2 e = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
                Z[u] = ffma2(e, one2, Z[u]);
#pragma unroll
                for (int k = 0; k < K; ++k) acc[u][k] = ffma2(e, v2[k], acc[u][k]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) {
        const int q = q0 + u * kIcpCThreads;
        if (q >= n1) continue;
        const float invZ = 1.0f / (Z[u].x + Z[u].y);
#pragma unroll
        for (int k = 0; k < K; ++k) out[(static_cast<size_t>(b) * n1 + q) * K + k] = (acc[u][k].x + acc[u][k].y) * invZ;
    }
}

// ---- warp-per-query forms (round 1): 32 lanes share one source point and split the targets; 32x the threads of the
// thread-per-query forms above, which is what a SMALL problem needs (mask_voting on single frames, a handful of clouds) ----
constexpr int kIcpThreads = 512;
constexpr int kIcpWarps = kIcpThreads / 32;

template <int K>
__global__ void __launch_bounds__(kIcpThreads)
icp_correspond_warp_kernel(int n1, int n2, float inv_temp, const float *__restrict__ pc1, const float *__restrict__ flow,
                      const float *__restrict__ pc2, const float *__restrict__ mask1, const float *__restrict__ mask2,
                      float *__restrict__ flow_out) {
    __shared__ float sp[kIcpTile * 3];
    __shared__ float sm[kIcpTile * K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int q = blockIdx.x * kIcpWarps + warp;
    const bool has_q = q < n1;
    pc2 += static_cast<size_t>(b) * n2 * 3;
    mask2 += static_cast<size_t>(b) * n2 * K;

    float qx = 0.f, qy = 0.f, qz = 0.f, px = 0.f, py = 0.f, pz = 0.f, m1[K];
#pragma unroll
    for (int k = 0; k < K; ++k) m1[k] = 0.f;
    if (has_q) {
        const size_t o = (static_cast<size_t>(b) * n1 + q) * 3;
        px = __ldg(pc1 + o); py = __ldg(pc1 + o + 1); pz = __ldg(pc1 + o + 2);
        qx = px + __ldg(flow + o); qy = py + __ldg(flow + o + 1); qz = pz + __ldg(flow + o + 2);
#pragma unroll
        for (int k = 0; k < K; ++k) m1[k] = __ldg(mask1 + (static_cast<size_t>(b) * n1 + q) * K + k);
    }
    float mx = -INFINITY, Z = 0.f, A = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
    for (int t0 = 0; t0 < n2; t0 += kIcpTile) {
        const int tn = min(kIcpTile, n2 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 3; i += kIcpThreads) sp[i] = __ldg(pc2 + static_cast<size_t>(t0) * 3 + i);
        for (int i = threadIdx.x; i < tn * K; i += kIcpThreads) sm[i] = __ldg(mask2 + static_cast<size_t>(t0) * K + i);
        __syncthreads();
        if (!has_q) continue;
        for (int j = lane; j < tn; j += 32) {
            const float x = sp[j * 3], y = sp[j * 3 + 1], z = sp[j * 3 + 2];
            const float dx = qx - x, dy = qy - y, dz = qz - z;
            const float s = -sqrtf(dx * dx + dy * dy + dz * dz) * inv_temp;
            float c = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) c = fmaf(m1[k], sm[j * K + k], c);
            if (s > mx) {                     // rescale the running sums to the new maximum
                const float r = __expf(mx - s);
                Z *= r; A *= r; vx *= r; vy *= r; vz *= r;
                mx = s;
            }
            const float e = __expf(s - mx);
            const float ec = e * c;
            Z += e; A += ec;
            vx = fmaf(ec, x, vx); vy = fmaf(ec, y, vy); vz = fmaf(ec, z, vz);
        }
    }
    if (!has_q) return;
    // merge the 32 lane states
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float omx = __shfl_xor_sync(OGC_FULL_MASK, mx, o);
        const float oZ = __shfl_xor_sync(OGC_FULL_MASK, Z, o), oA = __shfl_xor_sync(OGC_FULL_MASK, A, o);
        const float ox = __shfl_xor_sync(OGC_FULL_MASK, vx, o), oy = __shfl_xor_sync(OGC_FULL_MASK, vy, o),
                    oz = __shfl_xor_sync(OGC_FULL_MASK, vz, o);
        const float nm = fmaxf(mx, omx);
        const float ra = (mx == -INFINITY) ? 0.f : __expf(mx - nm), rb = (omx == -INFINITY) ? 0.f : __expf(omx - nm);
        Z = Z * ra + oZ * rb; A = A * ra + oA * rb;
        vx = vx * ra + ox * rb; vy = vy * ra + oy * rb; vz = vz * ra + oz * rb;
        mx = nm;
    }
    if (lane == 0) {
        const float invZ = 1.0f / Z;
        const float rs = fmaxf(A * invZ, 1e-10f);
        const size_t o = (static_cast<size_t>(b) * n1 + q) * 3;
        flow_out[o] = (vx * invZ) / rs - px;
        flow_out[o + 1] = (vy * invZ) / rs - py;
        flow_out[o + 2] = (vz * invZ) / rs - pz;
    }
}


// ---- soft correspondence transfer (vote.py) ---------------------------------------------------------------
// out[m,:] = sum_n softmax_n(-|q_m - key_n| / T) * val[n,:]      q (b,n1,3), key (b,n2,3), val (b,n2,K) -> out (b,n1,K)
//
// Replaces corr = softmax(-cdist(pc1 + flow, pc2) / T) (vote.py:17-28) FOLLOWED BY its only use, corr @ mask
// (vote.py:121), and -- because rows of a product of row-stochastic matrices still sum to one -- the chained
// propagation corr(t,t+2) = normalise(corr(t,t+1) @ corr(t+1,t+2)) (vote.py:50-57) as repeated application:
// corr(t,t+2) @ M = corr(t,t+1) @ (corr(t+1,t+2) @ M).  No N x N tensor and no N^3 bmm ever exists.
template <int K>
__global__ void __launch_bounds__(kIcpThreads)
softmax_transfer_warp_kernel(int n1, int n2, float inv_temp, const float *__restrict__ query, const float *__restrict__ key,
                        const float *__restrict__ val, float *__restrict__ out) {
    __shared__ float sp[kIcpTile * 3];
    __shared__ float sv[kIcpTile * K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int q = blockIdx.x * kIcpWarps + warp;
    const bool has_q = q < n1;
    key += static_cast<size_t>(b) * n2 * 3;
    val += static_cast<size_t>(b) * n2 * K;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (has_q) {
        const size_t o = (static_cast<size_t>(b) * n1 + q) * 3;
        qx = __ldg(query + o); qy = __ldg(query + o + 1); qz = __ldg(query + o + 2);
    }
    float mx = -INFINITY, Z = 0.f, acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    for (int t0 = 0; t0 < n2; t0 += kIcpTile) {
        const int tn = min(kIcpTile, n2 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 3; i += kIcpThreads) sp[i] = __ldg(key + static_cast<size_t>(t0) * 3 + i);
        for (int i = threadIdx.x; i < tn * K; i += kIcpThreads) sv[i] = __ldg(val + static_cast<size_t>(t0) * K + i);
        __syncthreads();
        if (!has_q) continue;
        for (int j = lane; j < tn; j += 32) {
            const float dx = qx - sp[j * 3], dy = qy - sp[j * 3 + 1], dz = qz - sp[j * 3 + 2];
            const float s = -sqrtf(dx * dx + dy * dy + dz * dz) * inv_temp;
            if (s > mx) {                     // rescale the running sums to the new maximum
                const float r = __expf(mx - s);
                Z *= r;
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] *= r;
                mx = s;
            }
            const float e = __expf(s - mx);
            Z += e;
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = fmaf(e, sv[j * K + k], acc[k]);
        }
    }
    if (!has_q) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float omx = __shfl_xor_sync(OGC_FULL_MASK, mx, o);
        const float oZ = __shfl_xor_sync(OGC_FULL_MASK, Z, o);
        const float nm = fmaxf(mx, omx);
        const float ra = (mx == -INFINITY) ? 0.f : __expf(mx - nm), rb = (omx == -INFINITY) ? 0.f : __expf(omx - nm);
        Z = Z * ra + oZ * rb;
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = acc[k] * ra + __shfl_xor_sync(OGC_FULL_MASK, acc[k], o) * rb;
        mx = nm;
    }
    if (lane == 0) {
        const float invZ = 1.0f / Z;
#pragma unroll
        for (int k = 0; k < K; ++k) out[(static_cast<size_t>(b) * n1 + q) * K + k] = acc[k] * invZ;
    }
}


}  // namespace ogc

namespace {

// (threads, queries per thread) by grid size
inline int icp_variant(int n1, int b) {
    const long long want = 2LL * ogc::kNumSMs;
    if (static_cast<long long>(n1) * b < 4LL * 32 * ogc::kNumSMs) return 3;      // < 4 warps of queries per SM: warp per query
    if (static_cast<long long>((n1 + 511) / 512) * b >= want) return 0;
    if (static_cast<long long>((n1 + 255) / 256) * b >= want) return 1;
    return 2;
}

template <int K>
void launch_correspond(int variant, int b, int n1, int n2, float it, const float *pc1, const float *flow, const float *pc2,
                       const float *mask1, const float *mask2, float *flow_out, cudaStream_t st) {
    using namespace ogc;
    if (variant == 0)
        icp_correspond_kernel<K, 256, 2><<<dim3((n1 + 511) / 512, b), 256, 0, st>>>(n1, n2, it, pc1, flow, pc2, mask1, mask2, flow_out);
    else if (variant == 1)
        icp_correspond_kernel<K, 256, 1><<<dim3((n1 + 255) / 256, b), 256, 0, st>>>(n1, n2, it, pc1, flow, pc2, mask1, mask2, flow_out);
    else if (variant == 2)
        icp_correspond_kernel<K, 64, 1><<<dim3((n1 + 63) / 64, b), 64, 0, st>>>(n1, n2, it, pc1, flow, pc2, mask1, mask2, flow_out);
    else
        icp_correspond_warp_kernel<K><<<dim3((n1 + kIcpWarps - 1) / kIcpWarps, b), kIcpThreads, 0, st>>>(n1, n2, it, pc1, flow, pc2, mask1, mask2, flow_out);
}

template <int K>
void launch_transfer(int variant, int b, int n1, int n2, float it, const float *query, const float *key, const float *val,
                     float *out, cudaStream_t st) {
    using namespace ogc;
    if (variant == 0)
        softmax_transfer_kernel<K, 256, 2><<<dim3((n1 + 511) / 512, b), 256, 0, st>>>(n1, n2, it, query, key, val, out);
    else if (variant == 1)
        softmax_transfer_kernel<K, 256, 1><<<dim3((n1 + 255) / 256, b), 256, 0, st>>>(n1, n2, it, query, key, val, out);
    else if (variant == 2)
        softmax_transfer_kernel<K, 64, 1><<<dim3((n1 + 63) / 64, b), 64, 0, st>>>(n1, n2, it, query, key, val, out);
    else
        softmax_transfer_warp_kernel<K><<<dim3((n1 + kIcpWarps - 1) / kIcpWarps, b), kIcpThreads, 0, st>>>(n1, n2, it, query, key, val, out);
}

}  // namespace

extern "C" int ogc_icp_correspond(int b, int n1, int n2, int k, float temperature, const float *pc1, const float *flow,
                                  const float *pc2, const float *mask1, const float *mask2, float *flow_out,
                                  void *stream) {
    using namespace ogc;
    if (b < 0 || n1 < 0 || n2 <= 0 || k < 1 || k > kIcpMaxK || !(temperature > 0.f)) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n1 == 0) return OGC_OK;
    if (!pc1 || !flow || !pc2 || !mask1 || !mask2 || !flow_out) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float it = 1.0f / temperature;
    const int var = icp_variant(n1, b);
    switch (k) {
#define OGC_ICP_CASE(KK) case KK: launch_correspond<KK>(var, b, n1, n2, it, pc1, flow, pc2, mask1, mask2, flow_out, st); break;
        OGC_ICP_CASE(1) OGC_ICP_CASE(2) OGC_ICP_CASE(3) OGC_ICP_CASE(4) OGC_ICP_CASE(5) OGC_ICP_CASE(6) OGC_ICP_CASE(7)
        OGC_ICP_CASE(8) OGC_ICP_CASE(9) OGC_ICP_CASE(10) OGC_ICP_CASE(11) OGC_ICP_CASE(12) OGC_ICP_CASE(13)
        OGC_ICP_CASE(14) OGC_ICP_CASE(15) OGC_ICP_CASE(16)
#undef OGC_ICP_CASE
    }
    OGC_RETURN_LAUNCH_STATUS();
}


extern "C" int ogc_softmax_transfer(int b, int n1, int n2, int k, float temperature, const float *query, const float *key,
                                    const float *val, float *out, void *stream) {
    using namespace ogc;
    if (b < 0 || n1 < 0 || n2 <= 0 || k < 1 || k > kIcpMaxK || !(temperature > 0.f)) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n1 == 0) return OGC_OK;
    if (!query || !key || !val || !out) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float it = 1.0f / temperature;
    const int var = icp_variant(n1, b);
    switch (k) {
#define OGC_ST_CASE(KK) case KK: launch_transfer<KK>(var, b, n1, n2, it, query, key, val, out, st); break;
        OGC_ST_CASE(1) OGC_ST_CASE(2) OGC_ST_CASE(3) OGC_ST_CASE(4) OGC_ST_CASE(5) OGC_ST_CASE(6) OGC_ST_CASE(7)
        OGC_ST_CASE(8) OGC_ST_CASE(9) OGC_ST_CASE(10) OGC_ST_CASE(11) OGC_ST_CASE(12) OGC_ST_CASE(13)
        OGC_ST_CASE(14) OGC_ST_CASE(15) OGC_ST_CASE(16)
#undef OGC_ST_CASE
    }
    OGC_RETURN_LAUNCH_STATUS();
}
