// BatchNorm shared MLP of the FlowStep3D blocks for sm_100a: (conv1x1 -> BatchNorm2d(batch statistics) -> ReLU) x L
// -> max over nsample, forward and backward.
//
// Replaces the torch-level stack of the reference's PointNetSetAbstraction / FlowEmbedding MLPs
// (utils/flowstep3d_util.py:52-64 and :126-137: per layer a Conv2d 1x1 (cuDNN / cuBLAS), a BatchNorm2d in training
// mode (cuDNN, statistics over batch x npoint x nsample), a ReLU, and a torch.max over the nsample axis at the end),
// each a pass over a (B,C,M,S) tensor, forward and backward.
//
// The contractions are the pointwise kernels of the feature-propagation block (fp_mlp.cu: pw_fwd_kernel with the
// previous layer's normalisation + ReLU folded into the operand loader; mlp_bwd.cu: mlp_dw_kernel / mlp_dx_kernel with
// dY rebuilt from dz, y and a per-channel coefficient table).  Those kernels read per-(sample, channel) tables, so
// BatchNorm only changes how the tables are made: this file holds the per-CHANNEL statistics over (batch, positions)
// and the tables derived from them (the same value repeated for every sample), the pooling over nsample, and the
// backward entry.  No normalised / rectified tensor is materialised; only the pre-norm y_l are stored.
#include "mlp_common.cuh"

namespace ogc {

constexpr float kBnEps = 1e-5f;   // nn.BatchNorm2d default (utils/flowstep3d_util.py:30,95)

__device__ __forceinline__ void block_sum2_atomic(double a, double b, double *dst) {
    __shared__ double red[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(OGC_FULL_MASK, a, o);
        b += __shfl_xor_sync(OGC_FULL_MASK, b, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = 0.0;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
        atomicAdd(dst, a);
        atomicAdd(dst + 1, b);
    }
}

// sums (C,2) fp64 += [sum y, sum y^2] over the P positions of one (sample, channel) row.  One CTA per row.
__global__ void __launch_bounds__(256)
bn_stats_kernel(int C, int P, const float *__restrict__ y, double *__restrict__ sums) {
    const int c = blockIdx.x, b = blockIdx.y;
    const float *yp = y + (static_cast<size_t>(b) * C + c) * P;
    float s = 0.f, q = 0.f;
    if ((P % 4 == 0) && (reinterpret_cast<uintptr_t>(yp) & 15u) == 0) {
        for (int p = threadIdx.x * 4; p < P; p += blockDim.x * 4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(yp + p));
            s += (v.x + v.y) + (v.z + v.w);
            q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
    } else {
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            const float v = __ldg(yp + p);
            s += v;
            q += v * v;
        }
    }
    block_sum2_atomic(static_cast<double>(s), static_cast<double>(q), sums + c * 2);
}

// Batch statistics -> tables.  ss (B,C,2) = [gamma*rstd, beta - mean*gamma*rstd] repeated over the samples,
// mean_rstd (C,2); biased variance for the normalisation, unbiased for the running estimate (nn.BatchNorm2d).
__global__ void bn_finalize_kernel(int B, int C, double n, const double *__restrict__ sums, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float *__restrict__ ss, float *__restrict__ mean_rstd,
                                   float *__restrict__ running_mean, float *__restrict__ running_var, float momentum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // thread = (sample, channel): every sample's table entry is
    if (i >= B * C) return;                                     // computed by its own thread (same inputs, same value)
    const int b = i / C, c = i - b * C;
    const double mean = sums[c * 2] / n;
    double var = sums[c * 2 + 1] / n - mean * mean;
    var = var > 0 ? var : 0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kBnEps)));
    const float sc = gamma[c] * rstd;
    ss[static_cast<size_t>(i) * 2] = sc;
    ss[static_cast<size_t>(i) * 2 + 1] = beta[c] - static_cast<float>(mean) * sc;
    if (b != 0) return;
    mean_rstd[c * 2] = static_cast<float>(mean);
    mean_rstd[c * 2 + 1] = rstd;
    if (running_mean && running_var) {
        const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
}

// out (B,C,M) = max over the S slots of a centre of relu(scale*y + shift) (ss != NULL) or of y itself (bare
// convolution blocks: use_act = False); sel (B,C,M) = first slot holding the maximum, 255 when the ReLU clamps it
// (no gradient).  Thread = (row, centre).
__global__ void __launch_bounds__(256)
bn_pool_kernel(long long rows, int M, int S, const float *__restrict__ y, const float *__restrict__ ss,
               float *__restrict__ out, unsigned char *__restrict__ sel) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows * M) return;
    const long long r = i / M;
    const float *yp = y + i * S;
    const bool act = ss != nullptr;
    const float sc = act ? __ldg(ss + r * 2) : 1.f, sh = act ? __ldg(ss + r * 2 + 1) : 0.f;
    float best = 0.f;
    int arg = 255;
    if ((S % 4 == 0) && (reinterpret_cast<uintptr_t>(yp) & 15u) == 0) {
        for (int s = 0; s < S; s += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(yp + s));
            const float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float z = act ? fmaf(sc, v[j], sh) : v[j];
                if (arg == 255 ? (!act || z > 0.f) : z > best) { best = z; arg = s + j; }
            }
        }
    } else {
        for (int s = 0; s < S; ++s) {
            const float v = __ldg(yp + s);
            const float z = act ? fmaf(sc, v, sh) : v;
            if (arg == 255 ? (!act || z > 0.f) : z > best) { best = z; arg = s; }
        }
    }
    out[i] = best;
    sel[i] = static_cast<unsigned char>(arg);
}

// Backward entry: dz (B,C,M*S) = the pooled gradient go (B,C,M) placed at the winning slot, zero elsewhere, and --
// when the block normalises (mean_rstd != NULL) -- the BatchNorm-backward sums of the last layer
// ab (C,2) fp64 += [sum dz, sum dz * yhat].  One CTA per (channel, sample) row.
__global__ void __launch_bounds__(256)
bn_pool_bwd_kernel(int C, int M, int S, const float *__restrict__ go, const unsigned char *__restrict__ sel,
                   const float *__restrict__ y, const float *__restrict__ mean_rstd, float *__restrict__ dz,
                   double *__restrict__ ab) {
    const int c = blockIdx.x, b = blockIdx.y;
    const size_t row = static_cast<size_t>(b) * C + c;
    const float mean = mean_rstd ? mean_rstd[c * 2] : 0.f, rstd = mean_rstd ? mean_rstd[c * 2 + 1] : 0.f;
    const float *gp = go + row * M, *yp = y + row * M * S;
    const unsigned char *sp = sel + row * M;
    float *zp = dz + row * M * S;
    const bool vec = (S % 4 == 0) && (reinterpret_cast<uintptr_t>(zp) & 15u) == 0;
    float s = 0.f, sy = 0.f;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const int a = sp[m];
        const float g = a < S ? __ldg(gp + m) : 0.f;
        if (vec) {
            for (int q = 0; q < S; q += 4)
                *reinterpret_cast<float4 *>(zp + static_cast<size_t>(m) * S + q) =
                    make_float4(a == q ? g : 0.f, a == q + 1 ? g : 0.f, a == q + 2 ? g : 0.f, a == q + 3 ? g : 0.f);
        } else {
            for (int q = 0; q < S; ++q) zp[static_cast<size_t>(m) * S + q] = a == q ? g : 0.f;
        }
        if (a < S) {
            s += g;
            sy += g * ((__ldg(yp + static_cast<size_t>(m) * S + a) - mean) * rstd);
        }
    }
    if (ab) block_sum2_atomic(static_cast<double>(s), static_cast<double>(sy), ab + c * 2);
}

// BatchNorm-backward sums of an inner layer from its (already ReLU-masked) dz and pre-norm y:
// ab (C,2) fp64 += [sum dz, sum dz * yhat].  One CTA per (channel, sample) row.
__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(int C, int P, const float *__restrict__ dz, const float *__restrict__ y,
                    const float *__restrict__ mean_rstd, double *__restrict__ ab) {
    const int c = blockIdx.x, b = blockIdx.y;
    const size_t row = static_cast<size_t>(b) * C + c;
    const float mean = mean_rstd[c * 2], rstd = mean_rstd[c * 2 + 1];
    const float *zp = dz + row * P, *yp = y + row * P;
    float s = 0.f, sy = 0.f;
    if ((P % 4 == 0) && ((reinterpret_cast<uintptr_t>(zp) | reinterpret_cast<uintptr_t>(yp)) & 15u) == 0) {
        for (int p = threadIdx.x * 4; p < P; p += blockDim.x * 4) {
            const float4 d = __ldg(reinterpret_cast<const float4 *>(zp + p));
            const float4 v = __ldg(reinterpret_cast<const float4 *>(yp + p));
            s += (d.x + d.y) + (d.z + d.w);
            sy += d.x * ((v.x - mean) * rstd) + d.y * ((v.y - mean) * rstd) + d.z * ((v.z - mean) * rstd) + d.w * ((v.w - mean) * rstd);
        }
    } else {
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            const float d = __ldg(zp + p);
            s += d;
            sy += d * ((__ldg(yp + p) - mean) * rstd);
        }
    }
    block_sum2_atomic(static_cast<double>(s), static_cast<double>(sy), ab + c * 2);
}

// coef (B,C,4) = [gamma rstd, gamma rstd A/n, gamma rstd^2 Bx/n, mean] repeated over the samples, so that the
// shared backward kernels rebuild dY = coef0*dz - coef1 - (y - coef3)*coef2 = gamma rstd (dz - A/n - yhat Bx/n);
// dgamma[c] += Bx, dbeta[c] += A.
__global__ void bn_bwd_coef_kernel(int B, int C, double n, const double *__restrict__ ab, const float *__restrict__ mean_rstd,
                                   const float *__restrict__ gamma, float *__restrict__ coef, float *__restrict__ dgamma,
                                   float *__restrict__ dbeta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // thread = (sample, channel), as bn_finalize_kernel
    if (i >= B * C) return;
    const int b = i / C, c = i - b * C;
    const double A = ab[c * 2], Bx = ab[c * 2 + 1];
    const float mean = mean_rstd[c * 2], rstd = mean_rstd[c * 2 + 1];
    const float k1 = gamma[c] * rstd;
    *reinterpret_cast<float4 *>(coef + static_cast<size_t>(i) * 4) =
        make_float4(k1, static_cast<float>(static_cast<double>(k1) * A / n),
                    static_cast<float>(static_cast<double>(k1) * rstd * Bx / n), mean);
    if (b != 0) return;
    atomicAdd(dgamma + c, static_cast<float>(Bx));
    atomicAdd(dbeta + c, static_cast<float>(A));
}

}  // namespace ogc

extern "C" int ogc_bn_stats(int b, int c, int p, const float *y, double *sums, void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || p <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!y || !sums) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    bn_stats_kernel<<<dim3(c, b), 256, 0, static_cast<cudaStream_t>(stream)>>>(c, p, y, sums);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_bn_finalize(int b, int c, long long count, const double *sums, const float *gamma, const float *beta,
                               float *scale_shift, float *mean_rstd, float *running_mean, float *running_var,
                               float momentum, void *stream) {
    using namespace ogc;
    if (b <= 0 || c <= 0 || count <= 0) return OGC_ERR_INVALID_ARG;
    if (!sums || !gamma || !beta || !scale_shift || !mean_rstd) return OGC_ERR_INVALID_ARG;
    bn_finalize_kernel<<<(b * c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        b, c, static_cast<double>(count), sums, gamma, beta, scale_shift, mean_rstd, running_mean, running_var, momentum);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_bn_pool(int b, int c, int m, int nsample, const float *y, const float *scale_shift, float *out,
                           unsigned char *sel, void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || m <= 0 || nsample <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!y || !out || !sel) return OGC_ERR_INVALID_ARG;
    if (nsample >= 255) return OGC_ERR_UNSUPPORTED;
    const long long rows = static_cast<long long>(b) * c, work = rows * m;
    bn_pool_kernel<<<static_cast<unsigned>((work + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rows, m, nsample, y, scale_shift, out, sel);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_bn_pool_bwd(int b, int c, int m, int nsample, const float *go, const unsigned char *sel, const float *y,
                               const float *mean_rstd, float *dz, double *ab, void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || m <= 0 || nsample <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!go || !sel || !y || !dz || ((mean_rstd == nullptr) != (ab == nullptr))) return OGC_ERR_INVALID_ARG;
    if (b > 65535 || nsample >= 255) return OGC_ERR_UNSUPPORTED;
    bn_pool_bwd_kernel<<<dim3(c, b), 256, 0, static_cast<cudaStream_t>(stream)>>>(c, m, nsample, go, sel, y, mean_rstd, dz, ab);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_bn_bwd_stats(int b, int c, int p, const float *dz, const float *y, const float *mean_rstd, double *ab,
                                void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || p <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!dz || !y || !mean_rstd || !ab) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    bn_bwd_stats_kernel<<<dim3(c, b), 256, 0, static_cast<cudaStream_t>(stream)>>>(c, p, dz, y, mean_rstd, ab);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_bn_bwd_coef(int b, int c, long long count, const double *ab, const float *mean_rstd, const float *gamma,
                               float *coef, float *dgamma, float *dbeta, void *stream) {
    using namespace ogc;
    if (b <= 0 || c <= 0 || count <= 0) return OGC_ERR_INVALID_ARG;
    if (!ab || !mean_rstd || !gamma || !coef || !dgamma || !dbeta) return OGC_ERR_INVALID_ARG;
    bn_bwd_coef_kernel<<<(b * c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        b, c, static_cast<double>(count), ab, mean_rstd, gamma, coef, dgamma, dbeta);
    OGC_RETURN_LAUNCH_STATUS();
}
