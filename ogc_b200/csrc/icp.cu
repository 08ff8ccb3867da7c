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

template <int K>
__global__ void __launch_bounds__(kIcpThreads)
icp_correspond_kernel(int n1, int n2, float inv_temp, const float *__restrict__ pc1, const float *__restrict__ flow,
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
    dim3 grid((n1 + kIcpWarps - 1) / kIcpWarps, b);
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
