// K2/K3 gather_points, K8/K9 group_points, K6/K7 three_interpolate (forward + backward), sm_100a.
//
// Replaces the one-thread-per-output-element kernels of the reference
// (pointnet2/src/sampling_gpu.cu:8-63, group_points_gpu.cu:8-66, interpolate_gpu.cu:149-214),
// which launch one thread per (batch, channel, element) and re-read idx / weight once per channel.
// These are pure data-movement ops (HBM bound): each thread here owns FOUR consecutive output
// elements, loads their indices once (one 16-byte load), keeps them in registers and walks a
// chunk of channels, so the index traffic is amortised over the channels and every store is a
// coalesced 16-byte store.  gather_points is group_points with nsample = 1.
// The backward ops accumulate with red.global.add.f32 exactly like the reference's atomicAdd
// (summation order unspecified in both).
#include "common.cuh"

namespace ogc {

constexpr int kGgThreads = 256;
constexpr int kChanChunk = 8;

// out[b,c,e] = points[b,c,idx[b,e]],  e in [0,E)
template <bool VEC4>
__global__ void __launch_bounds__(kGgThreads)
group_fwd_kernel(int c, int n, long long E, const float *__restrict__ points, const int *__restrict__ idx,
                 float *__restrict__ out) {
    const int bi = blockIdx.z;
    const int c0 = blockIdx.y * kChanChunk;
    const int c1 = min(c, c0 + kChanChunk);
    const long long t = static_cast<long long>(blockIdx.x) * kGgThreads + threadIdx.x;
    const int *ib = idx + static_cast<size_t>(bi) * E;
    if (VEC4) {
        const long long e = t * 4;
        if (e >= E) return;
        const int4 id = __ldg(reinterpret_cast<const int4 *>(ib + e));
        for (int ci = c0; ci < c1; ++ci) {
            const float *src = points + (static_cast<size_t>(bi) * c + ci) * n;
            float4 v;
            v.x = __ldg(src + id.x); v.y = __ldg(src + id.y); v.z = __ldg(src + id.z); v.w = __ldg(src + id.w);
            __stcs(reinterpret_cast<float4 *>(out + (static_cast<size_t>(bi) * c + ci) * E + e), v);
        }
    } else {
        if (t >= E) return;
        const int id = __ldg(ib + t);
        for (int ci = c0; ci < c1; ++ci)
            out[(static_cast<size_t>(bi) * c + ci) * E + t] = __ldg(points + (static_cast<size_t>(bi) * c + ci) * n + id);
    }
}

// grad_points[b,c,idx[b,e]] += grad_out[b,c,e]
template <bool VEC4>
__global__ void __launch_bounds__(kGgThreads)
group_bwd_kernel(int c, int n, long long E, const float *__restrict__ grad_out, const int *__restrict__ idx,
                 float *__restrict__ grad_points) {
    const int bi = blockIdx.z;
    const int c0 = blockIdx.y * kChanChunk;
    const int c1 = min(c, c0 + kChanChunk);
    const long long t = static_cast<long long>(blockIdx.x) * kGgThreads + threadIdx.x;
    const int *ib = idx + static_cast<size_t>(bi) * E;
    if (VEC4) {
        const long long e = t * 4;
        if (e >= E) return;
        const int4 id = __ldg(reinterpret_cast<const int4 *>(ib + e));
        for (int ci = c0; ci < c1; ++ci) {
            const float4 g = __ldcs(reinterpret_cast<const float4 *>(grad_out + (static_cast<size_t>(bi) * c + ci) * E + e));
            float *dst = grad_points + (static_cast<size_t>(bi) * c + ci) * n;
            atomicAdd(dst + id.x, g.x); atomicAdd(dst + id.y, g.y);
            atomicAdd(dst + id.z, g.z); atomicAdd(dst + id.w, g.w);
        }
    } else {
        if (t >= E) return;
        const int id = __ldg(ib + t);
        for (int ci = c0; ci < c1; ++ci)
            atomicAdd(grad_points + (static_cast<size_t>(bi) * c + ci) * n + id,
                      grad_out[(static_cast<size_t>(bi) * c + ci) * E + t]);
    }
}

// out[b,c,p] = fma(w2,f[i2], fma(w0,f[i0], w1*f[i1]))   (rounding order of the reference SASS)
__global__ void __launch_bounds__(kGgThreads)
interp_fwd_kernel(int c, int m, int n, const float *__restrict__ points, const int *__restrict__ idx,
                  const float *__restrict__ weight, float *__restrict__ out) {
    const int bi = blockIdx.z;
    const int c0 = blockIdx.y * kChanChunk;
    const int c1 = min(c, c0 + kChanChunk);
    const int p = blockIdx.x * kGgThreads + threadIdx.x;
    if (p >= n) return;
    const size_t o3 = (static_cast<size_t>(bi) * n + p) * 3;
    const int i0 = __ldg(idx + o3), i1 = __ldg(idx + o3 + 1), i2 = __ldg(idx + o3 + 2);
    const float w0 = __ldg(weight + o3), w1 = __ldg(weight + o3 + 1), w2 = __ldg(weight + o3 + 2);
    for (int ci = c0; ci < c1; ++ci) {
        const float *src = points + (static_cast<size_t>(bi) * c + ci) * m;
        const float v = __fmaf_rn(w2, __ldg(src + i2), __fmaf_rn(w0, __ldg(src + i0), __fmul_rn(w1, __ldg(src + i1))));
        out[(static_cast<size_t>(bi) * c + ci) * n + p] = v;
    }
}

// grad_points[b,c,i_j] += grad_out[b,c,p] * w_j
__global__ void __launch_bounds__(kGgThreads)
interp_bwd_kernel(int c, int n, int m, const float *__restrict__ grad_out, const int *__restrict__ idx,
                  const float *__restrict__ weight, float *__restrict__ grad_points) {
    const int bi = blockIdx.z;
    const int c0 = blockIdx.y * kChanChunk;
    const int c1 = min(c, c0 + kChanChunk);
    const int p = blockIdx.x * kGgThreads + threadIdx.x;
    if (p >= n) return;
    const size_t o3 = (static_cast<size_t>(bi) * n + p) * 3;
    const int i0 = __ldg(idx + o3), i1 = __ldg(idx + o3 + 1), i2 = __ldg(idx + o3 + 2);
    const float w0 = __ldg(weight + o3), w1 = __ldg(weight + o3 + 1), w2 = __ldg(weight + o3 + 2);
    for (int ci = c0; ci < c1; ++ci) {
        const float g = grad_out[(static_cast<size_t>(bi) * c + ci) * n + p];
        float *dst = grad_points + (static_cast<size_t>(bi) * c + ci) * m;
        atomicAdd(dst + i0, __fmul_rn(g, w0));
        atomicAdd(dst + i1, __fmul_rn(g, w1));
        atomicAdd(dst + i2, __fmul_rn(g, w2));
    }
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int group_common(bool fwd, int b, int c, int n, long long E, const float *src, const int *idx, float *dst,
                        void *stream) {
    if (b < 0 || c < 0 || n < 0 || E < 0) return OGC_ERR_INVALID_ARG;
    if (b == 0 || c == 0 || E == 0) return OGC_OK;
    if (!src || !idx || !dst) return OGC_ERR_INVALID_ARG;
    if (b > 65535 || (c + kChanChunk - 1) / kChanChunk > 65535) return OGC_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float *big = fwd ? dst : src;  // the (b,c,E) tensor
    const bool vec = (E % 4 == 0) && aligned16(idx) && aligned16(big);
    const long long work = vec ? E / 4 : E;
    dim3 grid(static_cast<unsigned>((work + kGgThreads - 1) / kGgThreads), (c + kChanChunk - 1) / kChanChunk, b);
    if (fwd) {
        if (vec) group_fwd_kernel<true><<<grid, kGgThreads, 0, st>>>(c, n, E, src, idx, dst);
        else group_fwd_kernel<false><<<grid, kGgThreads, 0, st>>>(c, n, E, src, idx, dst);
    } else {
        if (vec) group_bwd_kernel<true><<<grid, kGgThreads, 0, st>>>(c, n, E, src, idx, dst);
        else group_bwd_kernel<false><<<grid, kGgThreads, 0, st>>>(c, n, E, src, idx, dst);
    }
    OGC_RETURN_LAUNCH_STATUS();
}

}  // namespace ogc

extern "C" int ogc_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                                float *out, void *stream) {
    if (npoints < 0 || nsample < 0) return OGC_ERR_INVALID_ARG;
    return ogc::group_common(true, b, c, n, static_cast<long long>(npoints) * nsample, points, idx, out, stream);
}

extern "C" int ogc_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                     const int *idx, float *grad_points, void *stream) {
    if (npoints < 0 || nsample < 0) return OGC_ERR_INVALID_ARG;
    return ogc::group_common(false, b, c, n, static_cast<long long>(npoints) * nsample, grad_out, idx, grad_points,
                             stream);
}

extern "C" int ogc_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                                 void *stream) {
    return ogc::group_common(true, b, c, n, npoints, points, idx, out, stream);
}

extern "C" int ogc_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                                      float *grad_points, void *stream) {
    return ogc::group_common(false, b, c, n, npoints, grad_out, idx, grad_points, stream);
}

extern "C" int ogc_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                     const float *weight, float *out, void *stream) {
    using namespace ogc;
    if (b < 0 || c < 0 || m < 0 || n < 0) return OGC_ERR_INVALID_ARG;
    if (b == 0 || c == 0 || n == 0) return OGC_OK;
    if (!points || !idx || !weight || !out) return OGC_ERR_INVALID_ARG;
    if (b > 65535 || (c + kChanChunk - 1) / kChanChunk > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((n + kGgThreads - 1) / kGgThreads, (c + kChanChunk - 1) / kChanChunk, b);
    interp_fwd_kernel<<<grid, kGgThreads, 0, static_cast<cudaStream_t>(stream)>>>(c, m, n, points, idx, weight, out);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                          const float *weight, float *grad_points, void *stream) {
    using namespace ogc;
    if (b < 0 || c < 0 || m < 0 || n < 0) return OGC_ERR_INVALID_ARG;
    if (b == 0 || c == 0 || n == 0) return OGC_OK;
    if (!grad_out || !idx || !weight || !grad_points) return OGC_ERR_INVALID_ARG;
    if (b > 65535 || (c + kChanChunk - 1) / kChanChunk > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((n + kGgThreads - 1) / kGgThreads, (c + kChanChunk - 1) / kChanChunk, b);
    interp_bwd_kernel<<<grid, kGgThreads, 0, static_cast<cudaStream_t>(stream)>>>(c, n, m, grad_out, idx, weight,
                                                                                grad_points);
    OGC_RETURN_LAUNCH_STATUS();
}
