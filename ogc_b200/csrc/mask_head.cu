// Mask head of MaskFormer3D for sm_100a: per-point cosine similarity against the K slot embeddings, temperature,
// softmax -- forward and backward in one kernel each.
//
// Replaces (models/segnet_kitti.py:85-88)
//     mask = einsum('bdn,bdk->bnk', F.normalize(point_feats, dim=1), F.normalize(slot, dim=1)) / 0.05; softmax(-1)
// which torch runs as norm + clamp + div over the (B,64,N) feature map, a batched GEMM with a 10-wide output,
// a scale and a softmax -- and ~15 more passes over (B,64,N) / (B,N,K) tensors in the backward, one of them a
// K = 8192 reduction GEMM on 32 CTAs (205 us, profiles/r01_ncu_launches_bench_v6_narrow_ffma2.csv).
//
// Thread = point: its D = 64 features are read once (channel-major, coalesced across the warp), the K normalised
// slot vectors sit in shared memory.  Backward: d feats in the same pass; d slots (K x D sums over all points) from a
// shared-memory staging of the CTA's 128 points, one atomic per output per CTA.
#include "common.cuh"

namespace ogc {

constexpr int kMhThreads = 128;
constexpr int kMhD = 64;
constexpr int kMhMaxK = 16;

__global__ void __launch_bounds__(kMhThreads)
mask_head_fwd_kernel(int N, int K, float inv_temp, const float *__restrict__ feats, const float *__restrict__ slots_hat,
                     float *__restrict__ mask) {
    __shared__ float sh[kMhMaxK][kMhD];
    const int b = blockIdx.y, n = blockIdx.x * kMhThreads + threadIdx.x;
    for (int e = threadIdx.x; e < kMhD * K; e += kMhThreads) {
        const int d = e / K, k = e - d * K;
        sh[k][d] = __ldg(slots_hat + (static_cast<size_t>(b) * kMhD + d) * K + k);
    }
    __syncthreads();
    if (n >= N) return;
    float dot[kMhMaxK];
#pragma unroll
    for (int k = 0; k < kMhMaxK; ++k) dot[k] = 0.f;
    float nrm2 = 0.f;
    const float *fp = feats + static_cast<size_t>(b) * kMhD * N + n;
#pragma unroll 8
    for (int d = 0; d < kMhD; ++d) {
        const float f = __ldg(fp + static_cast<size_t>(d) * N);
        nrm2 = fmaf(f, f, nrm2);
#pragma unroll
        for (int k = 0; k < kMhMaxK; ++k)
            if (k < K) dot[k] = fmaf(f, sh[k][d], dot[k]);
    }
    const float inv = inv_temp / fmaxf(sqrtf(nrm2), 1e-12f);      // F.normalize: x / max(|x|, eps)
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMhMaxK; ++k)
        if (k < K) { dot[k] *= inv; mx = fmaxf(mx, dot[k]); }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kMhMaxK; ++k)
        if (k < K) { dot[k] = expf(dot[k] - mx); sum += dot[k]; }
    const float rs = 1.f / sum;
    float *mp = mask + (static_cast<size_t>(b) * N + n) * K;
#pragma unroll
    for (int k = 0; k < kMhMaxK; ++k)
        if (k < K) mp[k] = dot[k] * rs;
}

__global__ void __launch_bounds__(kMhThreads)
mask_head_bwd_kernel(int N, int K, float inv_temp, const float *__restrict__ feats, const float *__restrict__ slots_hat,
                     const float *__restrict__ mask, const float *__restrict__ dmask, float *__restrict__ dfeats,
                     float *__restrict__ dslots_hat) {
    __shared__ float sh[kMhMaxK][kMhD];
    __shared__ float fh_s[kMhThreads][kMhD + 1];       // normalised features of the CTA's points
    __shared__ float dl_s[kMhThreads][kMhMaxK];        // d logits
    const int tid = threadIdx.x, b = blockIdx.y, n = blockIdx.x * kMhThreads + tid;
    for (int e = tid; e < kMhD * K; e += kMhThreads) {
        const int d = e / K, k = e - d * K;
        sh[k][d] = __ldg(slots_hat + (static_cast<size_t>(b) * kMhD + d) * K + k);
    }
    __syncthreads();
    const bool live = n < N;
    float f[kMhD];
    float nrm2 = 0.f;
    const float *fp = feats + static_cast<size_t>(b) * kMhD * N + n;
#pragma unroll
    for (int d = 0; d < kMhD; ++d) {
        f[d] = live ? __ldg(fp + static_cast<size_t>(d) * N) : 0.f;
        nrm2 = fmaf(f[d], f[d], nrm2);
    }
    const float nrm = sqrtf(nrm2);
    const float inv = 1.f / fmaxf(nrm, 1e-12f);
    float dl[kMhMaxK];
    {
        float t = 0.f;
        float m[kMhMaxK];
#pragma unroll
        for (int k = 0; k < kMhMaxK; ++k) {
            m[k] = 0.f; dl[k] = 0.f;
            if (k < K && live) {
                m[k] = __ldg(mask + (static_cast<size_t>(b) * N + n) * K + k);
                dl[k] = __ldg(dmask + (static_cast<size_t>(b) * N + n) * K + k);
                t = fmaf(m[k], dl[k], t);
            }
        }
#pragma unroll
        for (int k = 0; k < kMhMaxK; ++k) {
            dl[k] = m[k] * (dl[k] - t) * inv_temp;       // softmax backward, then the 1/temperature scale
            dl_s[tid][k] = dl[k];
        }
    }
    // d fhat = sum_k dl_k shat_k ;  d f = (d fhat - fhat (fhat . d fhat)) / |f|   (no projection below eps)
    float proj = 0.f;
    float g[kMhD];
#pragma unroll
    for (int d = 0; d < kMhD; ++d) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < kMhMaxK; ++k)
            if (k < K) a = fmaf(dl[k], sh[k][d], a);
        g[d] = a;
        f[d] *= inv;                                     // fhat
        proj = fmaf(f[d], a, proj);
        fh_s[tid][d] = f[d];
    }
    if (nrm < 1e-12f) proj = 0.f;
    if (live) {
        float *gp = dfeats + static_cast<size_t>(b) * kMhD * N + n;
#pragma unroll
        for (int d = 0; d < kMhD; ++d) gp[static_cast<size_t>(d) * N] = (g[d] - f[d] * proj) * inv;
    }
    __syncthreads();
    // d shat[k][d] = sum over the CTA's points of dl[p][k] * fhat[p][d]
    for (int e = tid; e < kMhD * K; e += kMhThreads) {
        const int k = e / kMhD, d = e - k * kMhD;
        float a = 0.f;
#pragma unroll 8
        for (int p = 0; p < kMhThreads; ++p) a = fmaf(dl_s[p][k], fh_s[p][d], a);
        atomicAdd(dslots_hat + (static_cast<size_t>(b) * kMhD + d) * K + k, a);
    }
}

}  // namespace ogc

extern "C" int ogc_mask_head_fwd(int b, int d, int n, int k, float inv_temperature, const float *feats,
                                 const float *slots_hat, float *mask, void *stream) {
    using namespace ogc;
    if (b < 0 || d <= 0 || n <= 0 || k <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!feats || !slots_hat || !mask) return OGC_ERR_INVALID_ARG;
    if (d != kMhD || k > kMhMaxK || b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((n + kMhThreads - 1) / kMhThreads, b);
    mask_head_fwd_kernel<<<grid, kMhThreads, 0, static_cast<cudaStream_t>(stream)>>>(n, k, inv_temperature, feats, slots_hat, mask);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_mask_head_bwd(int b, int d, int n, int k, float inv_temperature, const float *feats,
                                 const float *slots_hat, const float *mask, const float *dmask, float *dfeats,
                                 float *dslots_hat, void *stream) {
    using namespace ogc;
    if (b < 0 || d <= 0 || n <= 0 || k <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!feats || !slots_hat || !mask || !dmask || !dfeats || !dslots_hat) return OGC_ERR_INVALID_ARG;
    if (d != kMhD || k > kMhMaxK || b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((n + kMhThreads - 1) / kMhThreads, b);
    mask_head_bwd_kernel<<<grid, kMhThreads, 0, static_cast<cudaStream_t>(stream)>>>(n, k, inv_temperature, feats, slots_hat, mask,
                                                                                    dmask, dfeats, dslots_hat);
    OGC_RETURN_LAUNCH_STATUS();
}
