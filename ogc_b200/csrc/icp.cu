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
#include "common.cuh"

namespace ogc {

constexpr int kIcpThreads = 512;
constexpr int kIcpWarps = kIcpThreads / 32;
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

constexpr int kIcpQ = 4;           // source points per warp: one read of a target point from shared memory serves all of them

// The shared-memory reads (3 + K words per pair) bounded the one-query-per-warp version; with kIcpQ queries per warp they
// are amortised, and the queries are processed in PAIRS on packed fp32 FMAs (a 3-register FFMA issues every other cycle,
// FFMA2 retires two per slot): distance, the K-term consistency dot product and the running sums of two queries per
// instruction.  exp(x) = 2^(x log2 e) with log2 e folded into 1/T.  Each query keeps its own online-softmax state and visits
// the targets in the same order as the one-query version.  199 -> 139 ms for 64 clouds x 8192 points x 20 iterations.
template <int K>
__global__ void __launch_bounds__(kIcpThreads)
icp_correspond_kernel(int n1, int n2, float inv_temp, const float *__restrict__ pc1, const float *__restrict__ flow,
                      const float *__restrict__ pc2, const float *__restrict__ mask1, const float *__restrict__ mask2,
                      float *__restrict__ flow_out) {
    __shared__ float sp[kIcpTile * 3];
    __shared__ float sm[kIcpTile * K];
    constexpr int NP = kIcpQ / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int q0 = (blockIdx.x * kIcpWarps + warp) * kIcpQ;
    pc2 += static_cast<size_t>(b) * n2 * 3;
    mask2 += static_cast<size_t>(b) * n2 * K;
    const float scale2 = -inv_temp * 1.4426950408889634f;       // s2 = -d / T * log2(e)

    float px[kIcpQ], py[kIcpQ], pz[kIcpQ];
    float2 qx[NP], qy[NP], qz[NP], m1[NP][K], Z[NP], A[NP], vx[NP], vy[NP], vz[NP];
    float mx[kIcpQ];
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) {
        const int q = q0 + u;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        px[u] = py[u] = pz[u] = 0.f;
        if (q < n1) {
            const size_t o = (static_cast<size_t>(b) * n1 + q) * 3;
            px[u] = __ldg(pc1 + o); py[u] = __ldg(pc1 + o + 1); pz[u] = __ldg(pc1 + o + 2);
            fx = px[u] + __ldg(flow + o); fy = py[u] + __ldg(flow + o + 1); fz = pz[u] + __ldg(flow + o + 2);
        }
        (u & 1 ? qx[u >> 1].y : qx[u >> 1].x) = fx;
        (u & 1 ? qy[u >> 1].y : qy[u >> 1].x) = fy;
        (u & 1 ? qz[u >> 1].y : qz[u >> 1].x) = fz;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float v = q < n1 ? __ldg(mask1 + (static_cast<size_t>(b) * n1 + q) * K + k) : 0.f;
            (u & 1 ? m1[u >> 1][k].y : m1[u >> 1][k].x) = v;
        }
        mx[u] = -INFINITY;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) Z[p] = A[p] = vx[p] = vy[p] = vz[p] = make_float2(0.f, 0.f);
    const float2 one2 = make_float2(1.f, 1.f), zero2 = make_float2(0.f, 0.f);
    const bool has_q = q0 < n1;
    for (int t0 = 0; t0 < n2; t0 += kIcpTile) {
        const int tn = min(kIcpTile, n2 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 3; i += kIcpThreads) sp[i] = __ldg(pc2 + static_cast<size_t>(t0) * 3 + i);
        for (int i = threadIdx.x; i < tn * K; i += kIcpThreads) sm[i] = __ldg(mask2 + static_cast<size_t>(t0) * K + i);
        __syncthreads();
        if (!has_q) continue;
        for (int j = lane; j < tn; j += 32) {
            const float x = sp[j * 3], y = sp[j * 3 + 1], z = sp[j * 3 + 2];
            const float2 nx = make_float2(-x, -x), ny = make_float2(-y, -y), nz = make_float2(-z, -z);
            const float2 x2 = make_float2(x, x), y2 = make_float2(y, y), z2 = make_float2(z, z);
            float2 m2[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { const float v = sm[j * K + k]; m2[k] = make_float2(v, v); }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float2 dx = ffma2(one2, qx[p], nx), dy = ffma2(one2, qy[p], ny), dz = ffma2(one2, qz[p], nz);
                float2 d2 = ffma2(dx, dx, zero2);
                d2 = ffma2(dy, dy, d2);
                d2 = ffma2(dz, dz, d2);
                const float s0 = sqrt_approx(d2.x) * scale2, s1 = sqrt_approx(d2.y) * scale2;
                float2 c = zero2;
#pragma unroll
                for (int k = 0; k < K; ++k) c = ffma2(m1[p][k], m2[k], c);
                if (s0 > mx[2 * p]) {                 // rescale the running sums of the first query to its new maximum
                    const float r = ex2_approx(mx[2 * p] - s0);
                    Z[p].x *= r; A[p].x *= r; vx[p].x *= r; vy[p].x *= r; vz[p].x *= r;
                    mx[2 * p] = s0;
                }
                if (s1 > mx[2 * p + 1]) {
                    const float r = ex2_approx(mx[2 * p + 1] - s1);
                    Z[p].y *= r; A[p].y *= r; vx[p].y *= r; vy[p].y *= r; vz[p].y *= r;
                    mx[2 * p + 1] = s1;
                }
                const float2 e = make_float2(ex2_approx(s0 - mx[2 * p]), ex2_approx(s1 - mx[2 * p + 1]));
                const float2 ec = ffma2(e, c, zero2);
                Z[p] = ffma2(e, one2, Z[p]);
                A[p] = ffma2(ec, one2, A[p]);
                vx[p] = ffma2(ec, x2, vx[p]); vy[p] = ffma2(ec, y2, vy[p]); vz[p] = ffma2(ec, z2, vz[p]);
            }
        }
    }
    if (!has_q) return;
#pragma unroll
    for (int u = 0; u < kIcpQ; ++u) {
        // merge the 32 lane states of query u (base-2 exponents)
        const int p = u >> 1;
        float mxu = mx[u], Zu = u & 1 ? Z[p].y : Z[p].x, Au = u & 1 ? A[p].y : A[p].x;
        float vxu = u & 1 ? vx[p].y : vx[p].x, vyu = u & 1 ? vy[p].y : vy[p].x, vzu = u & 1 ? vz[p].y : vz[p].x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float omx = __shfl_xor_sync(OGC_FULL_MASK, mxu, o);
            const float oZ = __shfl_xor_sync(OGC_FULL_MASK, Zu, o), oA = __shfl_xor_sync(OGC_FULL_MASK, Au, o);
            const float ox = __shfl_xor_sync(OGC_FULL_MASK, vxu, o), oy = __shfl_xor_sync(OGC_FULL_MASK, vyu, o),
                        oz = __shfl_xor_sync(OGC_FULL_MASK, vzu, o);
            const float nm = fmaxf(mxu, omx);
            const float ra = (mxu == -INFINITY) ? 0.f : ex2_approx(mxu - nm), rb = (omx == -INFINITY) ? 0.f : ex2_approx(omx - nm);
            Zu = Zu * ra + oZ * rb; Au = Au * ra + oA * rb;
            vxu = vxu * ra + ox * rb; vyu = vyu * ra + oy * rb; vzu = vzu * ra + oz * rb;
            mxu = nm;
        }
        const int q = q0 + u;
        if (lane == 0 && q < n1) {
            const float invZ = 1.0f / Zu;
            const float rs = fmaxf(Au * invZ, 1e-10f);
            const size_t o = (static_cast<size_t>(b) * n1 + q) * 3;
            flow_out[o] = (vxu * invZ) / rs - px[u];
            flow_out[o + 1] = (vyu * invZ) / rs - py[u];
            flow_out[o + 2] = (vzu * invZ) / rs - pz[u];
        }
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
softmax_transfer_kernel(int n1, int n2, float inv_temp, const float *__restrict__ query, const float *__restrict__ key,
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

extern "C" int ogc_icp_correspond(int b, int n1, int n2, int k, float temperature, const float *pc1, const float *flow,
                                  const float *pc2, const float *mask1, const float *mask2, float *flow_out,
                                  void *stream) {
    using namespace ogc;
    if (b < 0 || n1 < 0 || n2 <= 0 || k < 1 || k > kIcpMaxK || !(temperature > 0.f)) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n1 == 0) return OGC_OK;
    if (!pc1 || !flow || !pc2 || !mask1 || !mask2 || !flow_out) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((n1 + kIcpWarps * kIcpQ - 1) / (kIcpWarps * kIcpQ), b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float it = 1.0f / temperature;
    switch (k) {
#define OGC_ICP_CASE(KK) case KK: icp_correspond_kernel<KK><<<grid, kIcpThreads, 0, st>>>(n1, n2, it, pc1, flow, pc2, mask1, mask2, flow_out); break;
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
    dim3 grid((n1 + kIcpWarps - 1) / kIcpWarps, b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float it = 1.0f / temperature;
    switch (k) {
#define OGC_ST_CASE(KK) case KK: softmax_transfer_kernel<KK><<<grid, kIcpThreads, 0, st>>>(n1, n2, it, query, key, val, out); break;
        OGC_ST_CASE(1) OGC_ST_CASE(2) OGC_ST_CASE(3) OGC_ST_CASE(4) OGC_ST_CASE(5) OGC_ST_CASE(6) OGC_ST_CASE(7)
        OGC_ST_CASE(8) OGC_ST_CASE(9) OGC_ST_CASE(10) OGC_ST_CASE(11) OGC_ST_CASE(12) OGC_ST_CASE(13)
        OGC_ST_CASE(14) OGC_ST_CASE(15) OGC_ST_CASE(16)
#undef OGC_ST_CASE
    }
    OGC_RETURN_LAUNCH_STATUS();
}
