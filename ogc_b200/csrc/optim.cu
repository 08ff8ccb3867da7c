// Training-step plumbing kernels on ONE flat fp32 parameter / gradient buffer (sm_100a).
//
// They replace, for the data-parallel step of SURVEY.md 8(e):
//   ogc_count_nan   the per-parameter NaN scan `torch.any(torch.isnan(param.grad))` + early return of
//                   train_seg.py:81-83 (one D2H sync per parameter tensor in the reference) by one
//                   kernel that leaves a device-side counter -- no host sync;
//   ogc_adam_step   `optimizer.step()` (train_seg.py:85; torch.optim.Adam, lr / betas / eps /
//                   L2 weight-decay semantics) as one launch over the flat buffer, with the
//                   1/world_size gradient scale of the all-reduce folded in, and skipped on the device
//                   when the (all-reduced) NaN counter is non-zero.
#include "common.cuh"

namespace ogc {

__global__ void __launch_bounds__(256) count_nan_kernel(long long n, const float *__restrict__ g, float *__restrict__ counter) {
    long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    int bad = 0;
    for (; i < n; i += stride) bad |= isnan(g[i]) ? 1 : 0;
    bad = __any_sync(OGC_FULL_MASK, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(counter, 1.0f);
}

__global__ void __launch_bounds__(256)
adam_kernel(long long n, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
            float *__restrict__ v, float lr, float beta1, float beta2, float eps, float weight_decay,
            float bias_c1, float bias_c2_sqrt, float grad_scale, const float *__restrict__ skip_counter) {
    if (skip_counter && *skip_counter != 0.f) return;  // NaN somewhere (on any rank): keep the weights
    long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const float step = lr / bias_c1;
    for (; i < n; i += stride) {
        float gi = g[i] * grad_scale;
        const float pi = p[i];
        if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;          // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;     // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bias_c2_sqrt + eps;
        p[i] = pi - step * (mi / denom);
    }
}

// Device-resident optimizer state for CUDA-graph replay: the step counter and the learning rate live in device
// memory so that a captured step can be replayed with a new lr and an advancing bias correction.
// state[0] = step t (as float, exact up to 2^24), state[1] = lr for this step.
__global__ void adam_advance_kernel(float *__restrict__ state, const float *__restrict__ skip_counter) {
    if (skip_counter && *skip_counter != 0.f) return;
    state[0] += 1.0f;
}

__global__ void __launch_bounds__(256)
adam_dev_kernel(long long n, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                float *__restrict__ v, const float *__restrict__ state, float beta1, float beta2, float eps,
                float weight_decay, float grad_scale, const float *__restrict__ skip_counter) {
    if (skip_counter && *skip_counter != 0.f) return;
    const float t = state[0], lr = state[1];
    const float bias_c1 = 1.f - powf(beta1, t);
    const float bias_c2_sqrt = sqrtf(1.f - powf(beta2, t));
    long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const float step = lr / bias_c1;
    for (; i < n; i += stride) {
        float gi = g[i] * grad_scale;
        const float pi = p[i];
        if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step * (mi / (sqrtf(vi) / bias_c2_sqrt + eps));
    }
}

}  // namespace ogc

extern "C" int ogc_count_nan(long long n, const float *grad, float *counter, void *stream) {
    using namespace ogc;
    if (n < 0 || !counter) return OGC_ERR_INVALID_ARG;
    if (n == 0) return OGC_OK;
    if (!grad) return OGC_ERR_INVALID_ARG;
    const long long want = (n + 255) / 256;
    const int blocks = static_cast<int>(want < kNumSMs * 8 ? want : kNumSMs * 8);
    count_nan_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(n, grad, counter);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_adam_step(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                             const float *skip_counter, void *stream) {
    using namespace ogc;
    if (n < 0 || step < 1) return OGC_ERR_INVALID_ARG;
    if (n == 0) return OGC_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq) return OGC_ERR_INVALID_ARG;
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    const long long want = (n + 255) / 256;
    const int blocks = static_cast<int>(want < kNumSMs * 8 ? want : kNumSMs * 8);
    adam_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(n, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2,
                                                                      eps, weight_decay, static_cast<float>(bc1),
                                                                      static_cast<float>(sqrt(bc2)), grad_scale, skip_counter);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_adam_step_dev(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                                 float *state, float beta1, float beta2, float eps, float weight_decay,
                                 float grad_scale, const float *skip_counter, void *stream) {
    using namespace ogc;
    if (n < 0) return OGC_ERR_INVALID_ARG;
    if (n == 0) return OGC_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !state) return OGC_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    adam_advance_kernel<<<1, 1, 0, st>>>(state, skip_counter);
    const long long want = (n + 255) / 256;
    const int blocks = static_cast<int>(want < kNumSMs * 8 ? want : kNumSMs * 8);
    adam_dev_kernel<<<blocks, 256, 0, st>>>(n, param, grad, exp_avg, exp_avg_sq, state, beta1, beta2, eps, weight_decay,
                                           grad_scale, skip_counter);
    OGC_RETURN_LAUNCH_STATUS();
}
