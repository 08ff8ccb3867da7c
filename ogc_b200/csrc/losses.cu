// Fused OGC-loss kernels for sm_100a (forward AND the gradient w.r.t. the soft mask in one pass).
//
// They replace torch-level code of the reference, not CUDA kernels:
//   ogc_weighted_kabsch / ogc_dynamic_loss   losses/seg_loss_unsup.py:10-61 (fit_motion_svd_batch) and
//                                            :72-98 (DynamicLoss.forward) + its autograd backward.
//       The reference builds diag_embed(mask): a (B*K, N, N) tensor = 10.7 GB at B=4, K=10, N=8192
//       (:36) to obtain 3x3 covariances.  Here: one CTA per cloud, two passes over the N points for the
//       weighted means and the centred 3x3 covariances of all K segments, a per-segment 3x3 SVD in
//       fp64 on one thread each, then one pass for the loss and d loss / d mask.  Nothing N x N exists.
//   ogc_neighbor_l1                          :112-129 (KnnLoss) / :143-158 (BallQLoss), loss_norm = 1:
//       radius clip + grouping_operation(mask) + |.|_1 + mean over neighbours, and the backward
//       (group_points_grad's atomicAdd scatter + the centre term) fused in one kernel: the (B,K,N,S)
//       gathered-mask tensor (21 MB / cloud at S=64) is never materialised.
//   ogc_mask_contingency / ogc_invariance_loss   :212-280 (match_mask_by_iou's one-hot einsum IoU;
//       InvarianceLoss.distance with the permutation given as an index vector).
#include <math.h>

#include "common.cuh"
#include "lsap.cuh"

namespace ogc {

constexpr int kLossThreads = 1024;
constexpr int kMaxSlots = 32;  // K (n_slot) limit of the fused loss kernels

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(OGC_FULL_MASK, v, o);
    return v;
}

// ---- 3x3 rotation from the covariance S = sum w (p1-mu1)(p2-mu2)^T  (Kabsch) ---------------------
// torch: U,S,V = svd(S); R = V diag(1,1,det(V U^T)) U^T.  Algebraically R = v1 u1^T + v2 u2^T +
// (v1 x v2)(u1 x u2)^T for the two leading singular pairs -- the reflection fix is implicit and the
// smallest singular pair is never needed, so rank-2 covariances (planar segments) are handled
// exactly.  V from a cyclic Jacobi eigen-decomposition of S^T S in fp64.
__device__ void kabsch_rotation(const double S[9], double R[9]) {
    double B[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int r = 0; r < 3; ++r) s += S[r * 3 + i] * S[r * 3 + j];
            B[i][j] = s;
        }
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = fabs(B[0][1]) + fabs(B[0][2]) + fabs(B[1][2]);
        const double diag = fabs(B[0][0]) + fabs(B[1][1]) + fabs(B[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(B[p][q]) < 1e-300) continue;
                const double theta = (B[q][q] - B[p][p]) / (2.0 * B[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int r = 0; r < 3; ++r) {  // B <- B J
                    const double bp = B[r][p], bq = B[r][q];
                    B[r][p] = c * bp - s * bq;
                    B[r][q] = s * bp + c * bq;
                }
                for (int r = 0; r < 3; ++r) {  // B <- J^T B
                    const double bp = B[p][r], bq = B[q][r];
                    B[p][r] = c * bp - s * bq;
                    B[q][r] = s * bp + c * bq;
                }
                for (int r = 0; r < 3; ++r) {  // V <- V J
                    const double vp = V[r][p], vq = V[r][q];
                    V[r][p] = c * vp - s * vq;
                    V[r][q] = s * vp + c * vq;
                }
            }
    }
    // order eigenvalues descending -> i0 >= i1 >= i2
    int i0 = 0, i1 = 1, i2 = 2;
    if (B[i1][i1] > B[i0][i0]) { int t = i0; i0 = i1; i1 = t; }
    if (B[i2][i2] > B[i0][i0]) { int t = i0; i0 = i2; i2 = t; }
    if (B[i2][i2] > B[i1][i1]) { int t = i1; i1 = i2; i2 = t; }
    double v1[3] = {V[0][i0], V[1][i0], V[2][i0]}, v2[3] = {V[0][i1], V[1][i1], V[2][i1]};
    double u1[3], u2[3];
    for (int r = 0; r < 3; ++r) {
        u1[r] = S[r * 3 + 0] * v1[0] + S[r * 3 + 1] * v1[1] + S[r * 3 + 2] * v1[2];
        u2[r] = S[r * 3 + 0] * v2[0] + S[r * 3 + 1] * v2[1] + S[r * 3 + 2] * v2[2];
    }
    const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (!(n1 > 0.0)) {  // S == 0: every rotation is optimal; torch.svd returns U = V = I here
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    for (int r = 0; r < 3; ++r) u1[r] /= n1;
    const double dot = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    for (int r = 0; r < 3; ++r) u2[r] -= dot * u1[r];
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    if (!(n2 > 1e-14 * n1)) {
        // rank one: the rotation about u1 / v1 is not determined by the data (also ambiguous in the
        // reference's SVD); take the minimal rotation that maps u1 to v1's frame via an arbitrary normal
        const int a = fabs(u1[0]) < fabs(u1[1]) ? (fabs(u1[0]) < fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) < fabs(u1[2]) ? 1 : 2);
        double e[3] = {0, 0, 0};
        e[a] = 1.0;
        const double d = u1[a];
        for (int r = 0; r < 3; ++r) u2[r] = e[r] - d * u1[r];
        n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
        const int c = fabs(v1[0]) < fabs(v1[1]) ? (fabs(v1[0]) < fabs(v1[2]) ? 0 : 2) : (fabs(v1[1]) < fabs(v1[2]) ? 1 : 2);
        double f[3] = {0, 0, 0};
        f[c] = 1.0;
        const double dv = v1[c];
        double nv = 0;
        for (int r = 0; r < 3; ++r) { v2[r] = f[r] - dv * v1[r]; nv += v2[r] * v2[r]; }
        nv = sqrt(nv);
        for (int r = 0; r < 3; ++r) v2[r] /= nv;
    }
    for (int r = 0; r < 3; ++r) u2[r] /= n2;
    const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
    const double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j];
}

// One CTA per cloud.  mode bits: 1 = second operand is a flow (pc2 = pc + flow), else pc2 itself.
// Outputs (each optional): Rt (K,12) = R row-major then t; loss_pt (N); grad_mask (N,K).
__global__ void __launch_bounds__(kLossThreads, 1)
kabsch_loss_kernel(int n, int K, int second_is_flow, const float *__restrict__ pc, const float *__restrict__ second,
                   const float *__restrict__ mask, float *__restrict__ Rt_out, float *__restrict__ loss_pt,
                   float *__restrict__ grad_mask) {
    __shared__ double acc[kMaxSlots][9];
    __shared__ float mu[kMaxSlots][6];
    __shared__ float rt[kMaxSlots][12];
    const int tid = threadIdx.x, lane = tid & 31;
    const size_t bo = static_cast<size_t>(blockIdx.x);
    pc += bo * n * 3;
    second += bo * n * 3;
    mask += bo * n * K;

    auto load_p2 = [&](int i, const float p1[3], float p2[3]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float s = __ldg(second + i * 3 + c);
            p2[c] = second_is_flow ? __fadd_rn(p1[c], s) : s;
        }
    };

    // ---- pass 1: W = sum m, A = sum m p1, C = sum m p2  ------------------------------------------
    for (int i = tid; i < kMaxSlots * 9; i += blockDim.x) (&acc[0][0])[i] = 0.0;
    __syncthreads();
    for (int k0 = 0; k0 < K; k0 += 4) {
        float a[4][7];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < 7; ++c) a[u][c] = 0.f;
        for (int i = tid; i < n; i += blockDim.x) {
            float p1[3], p2[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) p1[c] = __ldg(pc + i * 3 + c);
            load_p2(i, p1, p2);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float m = (k0 + u < K) ? __ldg(mask + static_cast<size_t>(i) * K + k0 + u) : 0.f;
                a[u][0] += m;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    a[u][1 + c] = fmaf(m, p1[c], a[u][1 + c]);
                    a[u][4 + c] = fmaf(m, p2[c], a[u][4 + c]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const float s = warp_sum(a[u][c]);
                if (lane == 0 && k0 + u < K) atomicAdd(&acc[k0 + u][c], static_cast<double>(s));
            }
    }
    __syncthreads();
    if (tid < K) {
        // fp32 quotient, as einsum(...)/sum(mask) in the reference (:25-28)
        const float w = static_cast<float>(acc[tid][0]);
#pragma unroll
        for (int c = 0; c < 6; ++c) mu[tid][c] = static_cast<float>(acc[tid][1 + c]) / w;
    }
    __syncthreads();
    for (int i = tid; i < kMaxSlots * 9; i += blockDim.x) (&acc[0][0])[i] = 0.0;
    __syncthreads();

    // ---- pass 2: S = sum m (p1 - mu1)(p2 - mu2)^T  (rows from pc1, columns from pc2, :36) ---------
    for (int k0 = 0; k0 < K; k0 += 4) {
        float a[4][9];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < 9; ++c) a[u][c] = 0.f;
        for (int i = tid; i < n; i += blockDim.x) {
            float p1[3], p2[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) p1[c] = __ldg(pc + i * 3 + c);
            load_p2(i, p1, p2);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = min(k0 + u, K - 1);
                const float m = (k0 + u < K) ? __ldg(mask + static_cast<size_t>(i) * K + k) : 0.f;
                float x[3], y[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { x[c] = p1[c] - mu[k][c]; y[c] = m * (p2[c] - mu[k][3 + c]); }
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) a[u][r * 3 + c] = fmaf(x[r], y[c], a[u][r * 3 + c]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                const float s = warp_sum(a[u][c]);
                if (lane == 0 && k0 + u < K) atomicAdd(&acc[k0 + u][c], static_cast<double>(s));
            }
    }
    __syncthreads();

    // ---- per-segment rotation + translation ------------------------------------------------------
    if (tid < K) {
        double S[9], R[9];
        bool bad = false;
        for (int c = 0; c < 9; ++c) {
            S[c] = static_cast<double>(static_cast<float>(acc[tid][c]));  // the reference's S is fp32
            bad |= isnan(S[c]);
        }
        float Rf[9], tf[3];
        if (bad) {  // ill-posed segment -> identity (:40-42, :58-59)
            for (int c = 0; c < 9; ++c) Rf[c] = (c % 4 == 0) ? 1.f : 0.f;
            tf[0] = tf[1] = tf[2] = 0.f;
        } else {
            kabsch_rotation(S, R);
            for (int c = 0; c < 9; ++c) Rf[c] = static_cast<float>(R[c]);
            for (int r = 0; r < 3; ++r)  // t = mu2 - R mu1  (:56)
                tf[r] = mu[tid][3 + r] - (Rf[r * 3 + 0] * mu[tid][0] + Rf[r * 3 + 1] * mu[tid][1] + Rf[r * 3 + 2] * mu[tid][2]);
        }
        for (int c = 0; c < 9; ++c) rt[tid][c] = Rf[c];
        for (int c = 0; c < 3; ++c) rt[tid][9 + c] = tf[c];
        if (Rt_out)
            for (int c = 0; c < 12; ++c) Rt_out[(bo * K + tid) * 12 + c] = rt[tid][c];
    }
    if (!loss_pt && !grad_mask) return;
    __syncthreads();

    // ---- pass 3: loss_n = || sum_k m_k (R_k p + t_k) - p2 ||_2 ; grad_{n,k} = qhat . (R_k p + t_k) ----
    for (int i = tid; i < n; i += blockDim.x) {
        float p1[3], p2[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) p1[c] = __ldg(pc + i * 3 + c);
        load_p2(i, p1, p2);
        float q[3] = {0.f, 0.f, 0.f};
        for (int k = 0; k < K; ++k) {
            const float m = __ldg(mask + static_cast<size_t>(i) * K + k);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float tp = rt[k][r * 3 + 0] * p1[0] + rt[k][r * 3 + 1] * p1[1] + rt[k][r * 3 + 2] * p1[2] + rt[k][9 + r];
                q[r] += m * tp;
            }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) q[r] -= p2[r];
        const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
        if (loss_pt) loss_pt[bo * n + i] = nrm;
        if (grad_mask) {
            const float inv = nrm > 0.f ? 1.0f / nrm : 0.f;  // torch: subgradient 0 at the origin
            for (int k = 0; k < K; ++k) {
                float g = 0.f;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float tp = rt[k][r * 3 + 0] * p1[0] + rt[k][r * 3 + 1] * p1[1] + rt[k][r * 3 + 2] * p1[2] + rt[k][9 + r];
                    g += q[r] * inv * tp;
                }
                grad_mask[(bo * n + i) * K + k] = g;
            }
        }
    }
}

// out_n = sum_k m_k (R_k p_n + t_k) - p_n     (weighted_kabsch / object_aware_icp flow update,
// oa_icp.py:33-38, :78-83)
__global__ void __launch_bounds__(256)
apply_rigid_flow_kernel(int n, int K, const float *__restrict__ pc, const float *__restrict__ mask,
                        const float *__restrict__ Rt, float *__restrict__ flow_out) {
    __shared__ float rt[kMaxSlots][12];
    const size_t bo = blockIdx.y;
    for (int i = threadIdx.x; i < K * 12; i += blockDim.x) (&rt[0][0])[i] = __ldg(Rt + bo * K * 12 + i);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = pc + (bo * n + i) * 3;
    const float p1[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
    float q[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < K; ++k) {
        const float m = __ldg(mask + (bo * n + i) * K + k);
#pragma unroll
        for (int r = 0; r < 3; ++r)
            q[r] += m * (rt[k][r * 3 + 0] * p1[0] + rt[k][r * 3 + 1] * p1[1] + rt[k][r * 3 + 2] * p1[2] + rt[k][9 + r]);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) flow_out[(bo * n + i) * 3 + r] = q[r] - p1[r];
}

// ---- smoothness: one warp per point, lanes over the S neighbour slots -----------------------------
// loss_pt[n] = (1/S) sum_s sum_c | m[n,c] - m[nbr(n,s),c] |, nbr = idx[n,s] unless dist[n,s] > radius
// (then idx[n,0]).  grad_mask += coef * d(sum_n loss_pt)/d mask  (both the centre and the gathered
// term), accumulated with red.global.add like the reference's group_points_grad.
__global__ void __launch_bounds__(256)
neighbor_l1_kernel(int n, int K, int S, const float *__restrict__ mask, const int *__restrict__ idx,
                   const float *__restrict__ dist, float radius, int clip, float coef, float *__restrict__ loss_pt,
                   float *__restrict__ grad_mask) {
    const int lane = threadIdx.x & 31;
    const int pt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= n) return;
    const size_t bo = blockIdx.y;
    mask += bo * n * K;
    if (grad_mask) grad_mask += bo * n * K;
    const int *irow = idx + (bo * n + pt) * S;
    const float *drow = dist ? dist + (bo * n + pt) * S : nullptr;
    const int first = __ldg(irow);
    const float *crow = mask + static_cast<size_t>(pt) * K;
    float total = 0.f;
    const float gscale = coef / static_cast<float>(S);
    // K is small (8-16): loop over channels inside, neighbours across lanes.  The centre's own gradient is summed
    // across the warp (in units of gscale, exact small integers) and written with ONE atomic per channel; only the
    // gathered term scatters per neighbour.
    float cg[kMaxSlots];
#pragma unroll
    for (int c = 0; c < kMaxSlots; ++c) cg[c] = 0.f;
    for (int s0 = 0; s0 < S; s0 += 32) {
        const int s = s0 + lane;
        int j = -1;
        if (s < S) {
            j = __ldg(irow + s);
            if (clip && __ldg(drow + s) > radius) j = first;
        }
        if (j >= 0 && j != pt) {
            const float *nrow = mask + static_cast<size_t>(j) * K;
            if ((K & 1) == 0) {
                // even K: rows are 8-byte aligned -> 64-bit row loads and ONE vector reduction (red.v2.f32) per channel pair
                float *grow = grad_mask ? grad_mask + static_cast<size_t>(j) * K : nullptr;
#pragma unroll
                for (int c = 0; c < kMaxSlots; c += 2) {
                    if (c >= K) break;
                    const float2 a = __ldg(reinterpret_cast<const float2 *>(crow + c));
                    const float2 o = __ldg(reinterpret_cast<const float2 *>(nrow + c));
                    const float d0 = a.x - o.x, d1 = a.y - o.y;
                    total += fabsf(d0) + fabsf(d1);
                    if (grad_mask) {
                        const float s0 = d0 > 0.f ? 1.f : (d0 < 0.f ? -1.f : 0.f), s1 = d1 > 0.f ? 1.f : (d1 < 0.f ? -1.f : 0.f);
                        cg[c] += s0; cg[c + 1] += s1;
                        if (s0 != 0.f || s1 != 0.f)
                            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(grow + c), "f"(-s0 * gscale), "f"(-s1 * gscale) : "memory");
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < kMaxSlots; ++c) {
                    if (c >= K) break;
                    const float d = __ldg(crow + c) - __ldg(nrow + c);
                    total += fabsf(d);
                    if (grad_mask && d != 0.f) {
                        const float sg = d > 0.f ? 1.f : -1.f;
                        cg[c] += sg;
                        atomicAdd(grad_mask + static_cast<size_t>(j) * K + c, -sg * gscale);
                    }
                }
            }
        }
    }
    if (grad_mask) {
#pragma unroll
        for (int c = 0; c < kMaxSlots; ++c) {
            if (c >= K) break;
            const float v = warp_sum(cg[c]);
            if (lane == 0 && v != 0.f) atomicAdd(grad_mask + static_cast<size_t>(pt) * K + c, v * gscale);
        }
    }
    total = warp_sum(total);
    if (lane == 0 && loss_pt) loss_pt[bo * n + pt] = total / static_cast<float>(S);
}

// ---- invariance: hard-assignment contingency table -------------------------------------------------
// inter[b,g,p] = #{n : argmax m1[n] = g and argmax m2[n] = p}; row / column sums give the one-hot
// segment sizes.  argmax takes the FIRST maximum like torch.argmax.
__global__ void __launch_bounds__(256)
contingency_kernel(int n, int K, const float *__restrict__ m1, const float *__restrict__ m2, int *__restrict__ inter) {
    __shared__ int hist[kMaxSlots * kMaxSlots];
    const size_t bo = blockIdx.y;
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float *r1 = m1 + (bo * n + i) * K, *r2 = m2 + (bo * n + i) * K;
        int a1 = 0, a2 = 0;
        float b1 = __ldg(r1), b2 = __ldg(r2);
        for (int c = 1; c < K; ++c) {
            const float v1 = __ldg(r1 + c), v2 = __ldg(r2 + c);
            if (v1 > b1) { b1 = v1; a1 = c; }
            if (v2 > b2) { b2 = v2; a2 = c; }
        }
        atomicAdd(&hist[a1 * K + a2], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += blockDim.x)
        if (hist[i]) atomicAdd(inter + bo * K * K + i, hist[i]);
}

// loss_pt[n] = || m1[n] - m2[n, perm12] ||_2 + || m2[n] - m1[n, perm21] ||_2 ; targets are constants
// (detached, :270-273).  perm12[i] = column matched to row i (scipy col_ind).
__global__ void __launch_bounds__(256)
invariance_kernel(int n, int K, const float *__restrict__ m1, const float *__restrict__ m2,
                  const int *__restrict__ perm12, const int *__restrict__ perm21, float *__restrict__ loss_pt,
                  float *__restrict__ g1, float *__restrict__ g2) {
    __shared__ int p12[kMaxSlots], p21[kMaxSlots];
    const size_t bo = blockIdx.y;
    if (threadIdx.x < K) {
        p12[threadIdx.x] = perm12[bo * K + threadIdx.x];
        p21[threadIdx.x] = perm21[bo * K + threadIdx.x];
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *r1 = m1 + (bo * n + i) * K, *r2 = m2 + (bo * n + i) * K;
    float s1 = 0.f, s2 = 0.f;
    for (int c = 0; c < K; ++c) {
        const float d1 = __ldg(r1 + c) - __ldg(r2 + p12[c]);
        const float d2 = __ldg(r2 + c) - __ldg(r1 + p21[c]);
        s1 += d1 * d1;
        s2 += d2 * d2;
    }
    const float n1 = sqrtf(s1), n2 = sqrtf(s2);
    if (loss_pt) loss_pt[bo * n + i] = n1 + n2;
    const float i1 = n1 > 0.f ? 1.f / n1 : 0.f, i2 = n2 > 0.f ? 1.f / n2 : 0.f;
    for (int c = 0; c < K; ++c) {
        if (g1) g1[(bo * n + i) * K + c] = (__ldg(r1 + c) - __ldg(r2 + p12[c])) * i1;
        if (g2) g2[(bo * n + i) * K + c] = (__ldg(r2 + c) - __ldg(r1 + p21[c])) * i2;
    }
}

// One thread per (sample, direction): IoU of the hard assignments from the contingency counts exactly as
// losses/seg_loss_unsup.py:226-232 forms it in fp32 (intersection / clamp(|A| + |B| - intersection, 1e-10)),
// then the maximising assignment.  direction 0: rows = slots of mask1 (perm12), 1: the transpose (perm21).
__global__ void mask_match_kernel(int B, int K, const int *__restrict__ inter, int *__restrict__ perm12,
                                  int *__restrict__ perm21) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * B) return;
    const int b = t >> 1, dir = t & 1;
    const int *cnt = inter + static_cast<size_t>(b) * K * K;
    float rs[kLsapMax], cs[kLsapMax];
    for (int i = 0; i < K; ++i) { rs[i] = 0.f; cs[i] = 0.f; }
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            const float c = static_cast<float>(cnt[i * K + j]);
            rs[i] += c;
            cs[j] += c;
        }
    double cost[kLsapMax * kLsapMax];
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            const float in = static_cast<float>(dir == 0 ? cnt[i * K + j] : cnt[j * K + i]);
            const float un = dir == 0 ? (rs[i] + cs[j]) - in : (cs[i] + rs[j]) - in;   // same operand order as torch
            const float iou = __fdiv_rn(in, fmaxf(un, 1e-10f));
            cost[i * K + j] = -static_cast<double>(iou);
        }
    int col[kLsapMax];
    lsap_solve(K, cost, col);
    int *out = (dir == 0 ? perm12 : perm21) + static_cast<size_t>(b) * K;
    for (int i = 0; i < K; ++i) out[i] = col[i];
}

// Nuclear norm of the (N,K) soft mask of every sample: sum of singular values = sum sqrt(eig(M^T M)).
// Replaces RankLoss's `mask.norm(p='nuc', dim=(1,2))` (losses/seg_loss_unsup.py:313: a batched (N,K) SVD through
// cuSOLVER with a host sync, computed every step for a logged-only number).
//   gram_kernel : grid (chunks, B); each thread owns points n = tid, tid+256, ... of its chunk, reads the K-float
//                 row once and accumulates the K(K+1)/2 products (fp32 per thread, fp64 across threads/CTAs)
//   nuclear_from_gram_kernel : one thread per sample, cyclic Jacobi eigenvalues in fp64
constexpr int kGramMaxK = 16;     // register-resident pair accumulators; larger K uses the generic pair loop

template <int K>
__global__ void __launch_bounds__(256)
gram_kernel(int n, const float *__restrict__ mask, double *__restrict__ gram) {
    constexpr int NP = K * (K + 1) / 2;
    __shared__ double sg[NP];
    const int b = blockIdx.y;
    const float *m = mask + static_cast<size_t>(b) * n * K;
    for (int i = threadIdx.x; i < NP; i += blockDim.x) sg[i] = 0.0;
    __syncthreads();
    float acc[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i] = 0.f;
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(n, lo + per);
    for (int p = lo + threadIdx.x; p < hi; p += blockDim.x) {
        float r[K];
#pragma unroll
        for (int c = 0; c < K; ++c) r[c] = __ldg(m + static_cast<size_t>(p) * K + c);
        int q = 0;
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = i; j < K; ++j) { acc[q] = fmaf(r[i], r[j], acc[q]); ++q; }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(OGC_FULL_MASK, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sg[i], static_cast<double>(v));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NP; i += blockDim.x) atomicAdd(gram + static_cast<size_t>(b) * kMaxSlots * kMaxSlots + i, sg[i]);
}

// generic K (<= 32): warp per (i,j) pair
__global__ void __launch_bounds__(256)
gram_pairs_kernel(int n, int K, const float *__restrict__ mask, double *__restrict__ gram) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int b = blockIdx.y;
    const float *m = mask + static_cast<size_t>(b) * n * K;
    const int npairs = K * (K + 1) / 2;
    for (int pr = warp; pr < npairs; pr += nwarp) {
        int i = 0, rem = pr;
        while (rem >= K - i) { rem -= K - i; ++i; }
        const int j = i + rem;
        double acc = 0.0;
        for (int p = lane; p < n; p += 32)
            acc += static_cast<double>(__ldg(m + static_cast<size_t>(p) * K + i)) * static_cast<double>(__ldg(m + static_cast<size_t>(p) * K + j));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(OGC_FULL_MASK, acc, o);
        if (lane == 0) gram[static_cast<size_t>(b) * kMaxSlots * kMaxSlots + pr] = acc;
    }
}

// gram: per sample, the K(K+1)/2 upper-triangle entries in row-major pair order
__global__ void nuclear_from_gram_kernel(int B, int K, const double *__restrict__ gram, float *__restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double G[kMaxSlots][kMaxSlots];
    int q = 0;
    for (int i = 0; i < K; ++i)
        for (int j = i; j < K; ++j) {
            const double v = gram[static_cast<size_t>(b) * kMaxSlots * kMaxSlots + q++];
            G[i][j] = v;
            G[j][i] = v;
        }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < K; ++i) {
            diag += fabs(G[i][i]);
            for (int j = i + 1; j < K; ++j) off += fabs(G[i][j]);
        }
        if (off <= 1e-18 * diag || off == 0.0) break;
        for (int p = 0; p < K - 1; ++p)
            for (int r2 = p + 1; r2 < K; ++r2) {
                if (G[p][r2] == 0.0) continue;
                const double theta = (G[r2][r2] - G[p][p]) / (2.0 * G[p][r2]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s_ = t * c;
                for (int r = 0; r < K; ++r) {
                    const double gp = G[r][p], gq = G[r][r2];
                    G[r][p] = c * gp - s_ * gq;
                    G[r][r2] = s_ * gp + c * gq;
                }
                for (int r = 0; r < K; ++r) {
                    const double gp = G[p][r], gq = G[r2][r];
                    G[p][r] = c * gp - s_ * gq;
                    G[r2][r] = s_ * gp + c * gq;
                }
            }
    }
    double sum = 0.0;
    for (int i = 0; i < K; ++i) sum += sqrt(G[i][i] > 0.0 ? G[i][i] : 0.0);
    out[b] = static_cast<float>(sum);
}

}  // namespace ogc

extern "C" int ogc_weighted_kabsch(int b, int n, int k, int second_is_flow, const float *pc, const float *second,
                                   const float *mask, float *Rt, void *stream) {
    using namespace ogc;
    if (b < 0 || n <= 0 || k < 1 || k > kMaxSlots) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!pc || !second || !mask || !Rt) return OGC_ERR_INVALID_ARG;
    kabsch_loss_kernel<<<b, kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(n, k, second_is_flow, pc, second, mask, Rt,
                                                                                 nullptr, nullptr);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_dynamic_loss(int b, int n, int k, const float *pc, const float *flow, const float *mask,
                                float *loss_pt, float *grad_mask, float *Rt, void *stream) {
    using namespace ogc;
    if (b < 0 || n <= 0 || k < 1 || k > kMaxSlots) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!pc || !flow || !mask || !loss_pt) return OGC_ERR_INVALID_ARG;
    kabsch_loss_kernel<<<b, kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(n, k, 1, pc, flow, mask, Rt, loss_pt,
                                                                                 grad_mask);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_apply_rigid_flow(int b, int n, int k, const float *pc, const float *mask, const float *Rt,
                                    float *flow_out, void *stream) {
    using namespace ogc;
    if (b < 0 || n < 0 || k < 1 || k > kMaxSlots) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n == 0) return OGC_OK;
    if (!pc || !mask || !Rt || !flow_out) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((n + 255) / 256, b);
    apply_rigid_flow_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(n, k, pc, mask, Rt, flow_out);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_neighbor_l1(int b, int n, int k, int nsample, const float *mask, const int *idx, const float *dist,
                               float radius, float coef, float *loss_pt, float *grad_mask, void *stream) {
    using namespace ogc;
    if (b < 0 || n < 0 || k < 1 || k > kMaxSlots || nsample < 1) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n == 0) return OGC_OK;
    if (!mask || !idx) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((n + 7) / 8, b);
    neighbor_l1_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(n, k, nsample, mask, idx, dist, radius,
                                                                          dist != nullptr, coef, loss_pt, grad_mask);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_mask_contingency(int b, int n, int k, const float *mask1, const float *mask2, int *inter,
                                    void *stream) {
    using namespace ogc;
    if (b < 0 || n < 0 || k < 1 || k > kMaxSlots) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n == 0) return OGC_OK;
    if (!mask1 || !mask2 || !inter) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((n + 255) / 256, b);
    contingency_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(n, k, mask1, mask2, inter);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_invariance_loss(int b, int n, int k, const float *mask1, const float *mask2, const int *perm12,
                                   const int *perm21, float *loss_pt, float *grad1, float *grad2, void *stream) {
    using namespace ogc;
    if (b < 0 || n < 0 || k < 1 || k > kMaxSlots) return OGC_ERR_INVALID_ARG;
    if (b == 0 || n == 0) return OGC_OK;
    if (!mask1 || !mask2 || !perm12 || !perm21) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((n + 255) / 256, b);
    invariance_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(n, k, mask1, mask2, perm12, perm21, loss_pt,
                                                                         grad1, grad2);
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_mask_match(int b, int k, const int *inter, int *perm12, int *perm21, void *stream) {
    using namespace ogc;
    if (b < 0 || k < 1 || k > kLsapMax) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!inter || !perm12 || !perm21) return OGC_ERR_INVALID_ARG;
    mask_match_kernel<<<(2 * b + 31) / 32, 32, 0, static_cast<cudaStream_t>(stream)>>>(b, k, inter, perm12, perm21);
    OGC_RETURN_LAUNCH_STATUS();
}

// Host twin of the device assignment routine (same source, lsap.cuh) so that the port can be checked against
// scipy in the GPU-less test suite.  Not used by any product path.
extern "C" int ogc_lsap_maximize_host(int n, const double *score, int *col4row) {
    using namespace ogc;
    if (n < 1 || n > kLsapMax || !score || !col4row) return OGC_ERR_INVALID_ARG;
    double cost[kLsapMax * kLsapMax];
    for (int i = 0; i < n * n; ++i) cost[i] = -score[i];
    lsap_solve(n, cost, col4row);
    return OGC_OK;
}

extern "C" int ogc_mask_nuclear_norm(int b, int n, int k, const float *mask, float *out, double *gram_ws, void *stream) {
    using namespace ogc;
    if (b < 0 || n <= 0 || k < 1 || k > kMaxSlots) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!mask || !out || !gram_ws) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(gram_ws, 0, static_cast<size_t>(b) * kMaxSlots * kMaxSlots * sizeof(double), st);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int chunks = n >= 4096 ? 8 : 1;
    dim3 grid(chunks, b);
    switch (k) {
#define OGC_GRAM_CASE(KK) case KK: gram_kernel<KK><<<grid, 256, 0, st>>>(n, mask, gram_ws); break;
        OGC_GRAM_CASE(1) OGC_GRAM_CASE(2) OGC_GRAM_CASE(3) OGC_GRAM_CASE(4) OGC_GRAM_CASE(5) OGC_GRAM_CASE(6)
        OGC_GRAM_CASE(7) OGC_GRAM_CASE(8) OGC_GRAM_CASE(9) OGC_GRAM_CASE(10) OGC_GRAM_CASE(11) OGC_GRAM_CASE(12)
        OGC_GRAM_CASE(13) OGC_GRAM_CASE(14) OGC_GRAM_CASE(15) OGC_GRAM_CASE(16)
#undef OGC_GRAM_CASE
        default: gram_pairs_kernel<<<dim3(1, b), 256, 0, st>>>(n, k, mask, gram_ws); break;
    }
    nuclear_from_gram_kernel<<<(b + 31) / 32, 32, 0, st>>>(b, k, gram_ws, out);
    OGC_RETURN_LAUNCH_STATUS();
}
