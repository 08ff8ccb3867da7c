/*
 * ogc_b200.h -- C ABI of libogc_b200.so, the sm_100a (B200) implementation of the OGC
 * pointnet2 hot path.  This is the drop-in boundary: every entry point below replaces one
 * launcher of the reference `pointnet2_cuda` extension (vLAR-group/OGC, pointnet2/src),
 * with the same argument order and meaning, so the reference's five .cpp shims (or the
 * ctypes stub in INTEGRATION.md) can bind them one-for-one.
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only; no torch / ATen types.  All pointers are DEVICE pointers
 *     on the current CUDA device; fp32 data, int32 indices, contiguous, row-major with the
 *     shapes stated per function.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Work is
 *     enqueued on it; no entry point synchronises the host, allocates, frees or keeps state,
 *     so they are re-entrant (the autograd engine calls the *_grad functions from its own
 *     threads, pointnet2/pointnet2.py:69-76,169-185,214-228).
 *   - return value: 0 on success; a NEGATIVE ogc_status for rejected arguments (nothing
 *     launched); a POSITIVE cudaError_t if the launch failed.  The reference prints and calls
 *     exit(-1) instead (e.g. src/ball_query_gpu.cu:62-66); we never terminate the process.
 *   - outputs are always fully written (no pre-zero / pre-fill contract on the caller),
 *     EXCEPT the *_grad functions, which ACCUMULATE into a caller-zeroed buffer exactly like
 *     the reference's atomicAdd kernels.
 */
#ifndef OGC_B200_H
#define OGC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OGC_API __attribute__((visibility("default")))
#else
#define OGC_API
#endif

enum ogc_status {
    OGC_OK = 0,
    OGC_ERR_INVALID_ARG = -1,  /* negative size, NULL pointer, k out of range ...          */
    OGC_ERR_UNSUPPORTED = -2,  /* legal in principle but not implemented for these sizes   */
    OGC_ERR_WORKSPACE = -3     /* workspace missing or too small                           */
};

/* Library / build identification: "ogc_b200 <version> sm_100a". */
OGC_API const char *ogc_version(void);

/* ---------------------------------------------------------------------------------------
 * K1  furthest point sampling
 * replaces furthest_point_sampling_kernel_launcher(int b, int n, int m, const float *dataset,
 *          float *temp, int *idxs, cudaStream_t)            src/sampling_gpu.h:24-27,
 *          src/sampling_gpu.cu:93-253
 *   dataset (b,n,3) -> idxs (b,m).  idxs[.,0] = 0; ties between equal distances resolve
 *   exactly as the reference block reduction does for its block size 2^floor(log2 n) <= 1024
 *   (bit-exact indices).  `temp` (b,n) is the reference's running-min scratch: it is only
 *   used (and then clobbered) by the single-CTA fallback for n > 131072 (or when a thread-block
 *   cluster cannot be placed); pass it whenever n > 16384, otherwise it may be NULL and is untouched.
 *   It need not be pre-filled with 1e10.  16384 < n <= 131072 (raw scenes) runs on a cluster of 8 / 16 CTAs
 *   per cloud.
 * ------------------------------------------------------------------------------------- */
OGC_API int ogc_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                int *idxs, void *stream);

/* K2 / K3  gather_points(_grad)
 * replaces gather_points_kernel_launcher_fast / gather_points_grad_kernel_launcher_fast
 *          src/sampling_gpu.h:12-21, src/sampling_gpu.cu:8-84
 *   points (b,c,n), idx (b,npoints) -> out (b,c,npoints);   grad: grad_points += scatter */
OGC_API int ogc_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx,
                      float *out, void *stream);
OGC_API int ogc_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                           const int *idx, float *grad_points, void *stream);

/* K4  exact k nearest neighbours
 * replaces knn_kernel_launcher_fast(int b, int n, int m, int k, const float *unknown,
 *          const float *known, float *dist2, int *idx, cudaStream_t)
 *          src/interpolate_gpu.h:13-15, src/interpolate_gpu.cu:9-79
 *   unknown (b,n,3), known (b,m,3) -> dist2 (b,n,k) squared distances ascending, idx (b,n,k).
 *   Order: (distance, index) lexicographic; m < k leaves the tail at (+inf, 0).
 *   1 <= k <= 224 (the reference's stack arrays cap it at 200). */
OGC_API int ogc_knn(int b, int n, int m, int k, const float *unknown, const float *known, float *dist2,
            int *idx, void *stream);
/* Same, with the sqrt of pointnet2/pointnet2.py:103 fused into the store (dist = sqrtf(dist2),
 * IEEE round-to-nearest, identical to torch.sqrt). */
OGC_API int ogc_knn_sqrt(int b, int n, int m, int k, const float *unknown, const float *known,
                 float *dist, int *idx, void *stream);

/* K5  three nearest neighbours
 * replaces three_nn_kernel_launcher_fast   src/interpolate_gpu.h:17-19, .cu:81-146 */
OGC_API int ogc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                 int *idx, void *stream);

/* K6 / K7  three_interpolate(_grad)
 * replaces three_interpolate_kernel_launcher_fast / three_interpolate_grad_kernel_launcher_fast
 *          src/interpolate_gpu.h:22-34, .cu:149-236
 *   points (b,c,m), idx (b,n,3), weight (b,n,3) -> out (b,c,n)
 *   out = fma(w2,p2, fma(w0,p0, w1*p1))  -- the rounding order of the reference build.
 *   grad: grad_points (b,c,m) += grad_out (b,c,n) * weight, scattered by idx. */
OGC_API int ogc_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, void *stream);
OGC_API int ogc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                               const int *idx, const float *weight, float *grad_points,
                               void *stream);

/* K8 / K9  group_points(_grad)
 * replaces group_points_kernel_launcher_fast / group_points_grad_kernel_launcher_fast
 *          src/group_points_gpu.h:13-20, src/group_points_gpu.cu:8-89
 *   points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample) */
OGC_API int ogc_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out, void *stream);
OGC_API int ogc_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int *idx, float *grad_points, void *stream);

/* K10  ball query
 * replaces ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
 *          const float *new_xyz, const float *xyz, int *idx, cudaStream_t)
 *          src/ball_query_gpu.h:12-13 (definition order: src/ball_query_gpu.cu:48-49)
 *   xyz (b,n,3), new_xyz (b,m,3) -> idx (b,m,nsample): the first nsample indices (ascending)
 *   with d2 < radius*radius (fp32); remaining slots repeat the first hit; no hit -> zeros.
 *   Unlike the reference, idx need not be pre-zeroed. */
OGC_API int ogc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                   const float *xyz, int *idx, void *stream);

/* =======================================================================================
 * Fused OGC-loss entry points.  These replace torch-level code of the reference (there is no
 * reference CUDA kernel for them); each cites the Python it restates.  K = n_slot <= 32.
 * ===================================================================================== */

/* Per-segment weighted Kabsch: replaces fit_motion_svd_batch(pc1, pc2, mask)
 * losses/seg_loss_unsup.py:10-61 for the call pattern of DynamicLoss (:81-87), weighted_kabsch
 * (oa_icp.py:16-38) and object_aware_icp (oa_icp.py:77-79), where pc1/pc2 are one cloud repeated
 * over the K segments.
 *   pc (b,n,3); second (b,n,3) = flow (second_is_flow=1: pc2 = pc + flow) or pc2 itself;
 *   mask (b,n,k) soft assignment  ->  Rt (b,k,12): R (3x3 row-major) then t (3).
 *   Segments whose covariance contains NaN get R = I, t = 0 (:40-42,58-59). */
OGC_API int ogc_weighted_kabsch(int b, int n, int k, int second_is_flow, const float *pc,
                                const float *second, const float *mask, float *Rt, void *stream);

/* DynamicLoss.forward + backward: losses/seg_loss_unsup.py:72-98.
 *   loss_pt (b,n) = || sum_k m_k (R_k p + t_k) - (p + flow) ||_2   (the caller takes the mean)
 *   grad_mask (b,n,k) = d (sum_n loss_pt) / d mask with R,t constant (they are .detach()ed, :91);
 *   may be NULL.  Rt (b,k,12) optional output. */
OGC_API int ogc_dynamic_loss(int b, int n, int k, const float *pc, const float *flow, const float *mask,
                             float *loss_pt, float *grad_mask, float *Rt, void *stream);

/* flow_out (b,n,3) = sum_k m_k (R_k p + t_k) - p      oa_icp.py:33-38, :80-83 */
OGC_API int ogc_apply_rigid_flow(int b, int n, int k, const float *pc, const float *mask, const float *Rt,
                                 float *flow_out, void *stream);

/* KnnLoss / BallQLoss with loss_norm = 1: losses/seg_loss_unsup.py:112-129, :143-158, forward +
 * backward.  mask (b,n,k); idx (b,n,nsample) neighbour indices; dist (b,n,nsample) = sqrt'ed k-NN
 * distances or NULL: when given, neighbours with dist > radius are replaced by idx[.,.,0] (:121-122).
 *   loss_pt (b,n) = mean_s sum_c | m[n,c] - m[nbr(n,s),c] |      (may be NULL)
 *   grad_mask (b,n,k) += coef * d (sum_n loss_pt) / d mask       (ACCUMULATES; may be NULL) */
OGC_API int ogc_neighbor_l1(int b, int n, int k, int nsample, const float *mask, const int *idx,
                            const float *dist, float radius, float coef, float *loss_pt, float *grad_mask,
                            void *stream);

/* match_mask_by_iou's hard-assignment statistics: losses/seg_loss_unsup.py:221-232.
 *   inter (b,k,k) int32 += #{n : argmax mask1[n] = g, argmax mask2[n] = p}   (ACCUMULATES: zero it first)
 *   intersection = inter, segment sizes = its row / column sums, so the caller forms the IoU. */
OGC_API int ogc_mask_contingency(int b, int n, int k, const float *mask1, const float *mask2, int *inter,
                                 void *stream);

/* InvarianceLoss.forward + backward given the Hungarian matches: losses/seg_loss_unsup.py:252-280.
 *   perm12 (b,k): column of mask2 matched to slot i of mask1 (scipy col_ind); perm21 the converse.
 *   loss_pt (b,n) = ||m1 - m2[perm12]||_2 + ||m2 - m1[perm21]||_2 ; grad1/grad2 (b,n,k) the gradients
 *   of sum_n loss_pt w.r.t. mask1 / mask2 (targets are constants). */
OGC_API int ogc_invariance_loss(int b, int n, int k, const float *mask1, const float *mask2, const int *perm12,
                                const int *perm21, float *loss_pt, float *grad1, float *grad2, void *stream);

/* =======================================================================================
 * Training-step plumbing on one flat fp32 buffer (data-parallel step).
 * ===================================================================================== */

/* counter[0] += (number of warps that saw a NaN in grad[0..n)); replaces the per-parameter
 * `torch.any(torch.isnan(param.grad))` scan of train_seg.py:81-83 without a host sync. */
OGC_API int ogc_count_nan(long long n, const float *grad, float *counter, void *stream);

/* torch.optim.Adam step (train_seg.py:85, :328) over flat buffers: g = grad*grad_scale (+ wd*p);
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
 * `step` = t >= 1.  If skip_counter != NULL and *skip_counter != 0 nothing is updated (the
 * reference returns before optimizer.step() when a gradient holds a NaN). */
OGC_API int ogc_adam_step(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                          float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                          float grad_scale, const float *skip_counter, void *stream);

/* =======================================================================================
 * Fused set-abstraction MLP (csrc/mlp.cu, csrc/mlp_bwd.cu).  Replaces the torch-level stack of the
 * reference SA module: QueryAndGroup's two grouping_operation calls + concat
 * (pointnet2/pointnet2.py:287-294), SharedMLP = (Conv2d 1x1, GroupNorm(4), ReLU) x L
 * (utils/nn_util.py:151-168) and max_pool2d over nsample (utils/pointnet2_util.py:38-42), forward
 * and backward.  P = m*nsample positions per cloud, position p = centre*nsample + slot.
 * Per layer l: y_l = W_l a_{l-1} (pre-norm, stored once), a_l = relu(scale*y_l + shift).
 * ===================================================================================== */

/* One layer forward.  gather=1 (layer 1): a_0 is built on the fly from xyz (b,n,3), new_xyz (b,m,3),
 * point-major features feat_pm (b,n,cin-3) and neighbour indices idx (b,m,nsample): channels
 * [xyz_j - centre, feat_j].  gather=0: a_{l-1} = relu(ss_prev[.,0]*y_prev + ss_prev[.,1]).
 * wt = W^T (cin,cout).  y (b,cout,P) may be NULL.  sums (b,4,2) fp64 += [sum y, sum y^2] per
 * GroupNorm group (zero it first).  last=1 (nsample == 64): also ymax/ymin (b,cout,m) = max/min of y
 * over the nsample axis and amax/amin their first positions. */
OGC_API int ogc_sa_mlp_layer_fwd(int b, int n, int m, int nsample, int cin, int cout, int gather, int last,
                                 const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                                 const float *y_prev, const float *ss_prev, const float *wt, float *y,
                                 double *sums, float *ymax, float *ymin, unsigned char *amax,
                                 unsigned char *amin, void *stream);

/* GroupNorm(4) statistics -> per-channel affine: scale_shift (b,c,2) = [gamma*rstd, beta - mean*gamma*rstd],
 * mean_rstd (b,4,2).  count_per_group = (c/4)*P.  Biased variance, eps 1e-5 (nn.GroupNorm defaults). */
OGC_API int ogc_gn_finalize(int b, int c, long long count_per_group, const double *sums, const float *gamma,
                            const float *beta, float *scale_shift, float *mean_rstd, void *stream);

/* Pooled output of the last layer: out[b, c_offset+c, m] = relu(max_s(scale*y+shift)) written into a
 * (b,c_total,m) tensor (multi-scale concat) and optionally its point-major twin out_pm (b,m,c_total);
 * sel (b,c,m) = winning slot (255: clamped by the ReLU), ysel = its pre-norm value. */
OGC_API int ogc_sa_finish(int b, int c, int m, const float *ymax, const float *ymin, const unsigned char *amax,
                          const unsigned char *amin, const float *scale_shift, float *out, float *out_pm,
                          int c_total, int c_offset, unsigned char *sel, float *ysel, void *stream);

/* Backward, last layer: GroupNorm-backward sums from the sparse pooled gradient go (b,go_ctotal,m).
 * ab (b,4,2) fp64 += [sum gamma dz, sum gamma dz yhat]; dgamma/dbeta (c) += their per-channel parts. */
OGC_API int ogc_sa_last_stats(int b, int c, int m, const float *go, int go_ctotal, int go_coff,
                              const unsigned char *sel, const float *ysel, const float *mean_rstd,
                              const float *gamma, double *ab, float *dgamma, float *dbeta, void *stream);

/* coef (b,c,4) = [rstd*gamma, rstd*A/n, rstd^2*B/n, mean]: dY = coef0*dz - coef1 - (y - coef3)*coef2 */
OGC_API int ogc_gn_bwd_coef(int b, int c, long long count_per_group, const double *ab, const float *mean_rstd,
                            const float *gamma, float *coef, void *stream);

/* Input gradient of layer l: rows [row_off,row_off+rows) of W^T dY_l with W (cout,cin_full).
 * dY_l is rebuilt from dz (b,cout,P) -- or, when dz == NULL (last layer), from go/sel -- plus y and coef.
 * Dense mode (dfeat_pm == NULL): dz_prev (b,rows,P) = relu'(.) * result, and the GroupNorm-backward sums
 * of layer l-1 (ab_prev, dgamma_prev, dbeta_prev) are accumulated.
 * Scatter mode (layer 1): result is scatter-added through idx into dfeat_pm (b,n,dfeat_stride) at
 * channel offset dfeat_off (replaces group_points_grad, src/group_points_gpu.cu:8-25); rows <= 128 there.
 * Dense mode accepts any rows % 16 == 0 (one launch per 128-row block). */
OGC_API int ogc_sa_mlp_layer_dx(int b, int n, int m, int nsample, int cout, int cin_full, int row_off, int rows,
                                const float *dz, const float *go, int go_ctotal, int go_coff,
                                const unsigned char *sel, const float *y, const float *coef, const float *w,
                                const float *y_prev, const float *ss_prev, const float *mean_rstd_prev,
                                const float *gamma_prev, float *dz_prev, double *ab_prev, float *dgamma_prev,
                                float *dbeta_prev, const int *idx, float *dfeat_pm, int dfeat_stride,
                                int dfeat_off, void *stream);

/* Weight gradient of layer l: dw (cout,cin) += sum_{b,p} dY_l a_{l-1}^T (a_{l-1} as in the forward). */
OGC_API int ogc_sa_mlp_layer_dw(int b, int n, int m, int nsample, int cout, int cin, int gather, const float *dz,
                                const float *go, int go_ctotal, int go_coff, const unsigned char *sel,
                                const float *y, const float *coef, const float *y_prev, const float *ss_prev,
                                const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                                float *dw, void *stream);

/* ---- Uniform-grid acceleration of the radius-bounded searches (csrc/grid.cu) -------------------------------------------
 * ogc_grid_build counting-sorts every cloud xyz (b,m,3) into a grid whose cells are >= radius wide (at most
 * 32 x 8 x 32 cells): sorted (b,m,4) fp32 records (x, y, z, original index as int bits), cell_start
 * (ogc_grid_table_bytes(b) bytes), params (b,16) fp32.  ogc_knn_grid / ogc_ball_query_grid then visit only the 27
 * cells around a query and return EXACTLY what ogc_knn_bounded / ogc_ball_query return (interpolate_gpu.cu:9-57 order
 * rule (distance, index); ball_query_gpu.cu:9-45 first-nsample-in-index-order rule), for max_dist / radius <= the
 * radius the grid was built with.  unknown / new_xyz == NULL: the queries are the cloud itself (n == m). */
OGC_API long long ogc_grid_sorted_bytes(int b, int m);
OGC_API long long ogc_grid_table_bytes(int b);
OGC_API long long ogc_grid_params_bytes(int b);
OGC_API int ogc_grid_build(int b, int m, float radius, const float *xyz, void *sorted, int *cell_start, float *params,
                           void *stream);
OGC_API int ogc_knn_grid(int b, int n, int m, int k, float max_dist, const float *unknown, const void *sorted,
                         const int *cell_start, const float *params, float *dist, int *idx, void *stream);
OGC_API int ogc_ball_query_grid(int b, int n, int m, float radius, int nsample, const float *new_xyz, const void *sorted,
                                const int *cell_start, const float *params, int *idx, void *stream);

/* ---- Chained set-abstraction MLP (round 2; csrc/sa_chain_fwd.cu) ----------------------------------------------------
 * Replaces, for one grouper of a PointNet++ SA level, grouping_operation x2 + concat + (conv1x1, GroupNorm(4), ReLU) x L
 * (+ max over nsample) of utils/pointnet2_util.py:33-44 / pointnet2/pointnet2.py:283-294 / utils/nn_util.py:151-168
 * with positions on the tensor-core M axis and the activations of up to three layers chained through tensor memory.
 * One launch computes layers 1..nl for every position and reduces the LAST of them: GroupNorm sums, optionally the
 * stored pre-norm tensor y_out (b,c,m*64), optionally the max / min over each half centre (b,2m,c).  A grouper of L
 * layers is L launches (nl = 1..L); nothing but `sums` and the pooled extremes is written.
 *   gather != 0: input = [xyz[idx] - new_xyz (3) | feat_pm[idx] (cf)], w1 (c1, 3 + cf) row-major
 *   gather == 0: input = relu(scale * y_in + shift), y_in (b, cf, m*64), ss_in (b, cf, 2), w1 (c1, cf)
 * widths[l]: output channels of layer l (multiples of 32, <= 256).  ss1 / ss2: (b,c_l,2) scale/shift of layers 1 / 2.
 * nsample == 64, m even.  OGC_ERR_UNSUPPORTED when the shape does not fit (see ogc_sa_chain_fits). */
/* Diagnostics (not thread-safe): buf = device int64[3*64*8] or NULL; subsequent ogc_sa_chain_fwd launches record the
 * SM-clock timeline of CTA (0,0,0): [producer | MMA issuer | epilogue][tile][event]. */
OGC_API int ogc_sa_chain_debug(long long *buf);
OGC_API int ogc_sa_chain_fits(int m, int nsample, int cf, int gather, int nl, const int *widths);
OGC_API int ogc_sa_chain_fwd(int b, int n, int m, int nsample, int cf, int gather, int nl, const int *widths,
                             const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                             const float *y_in, const float *ss_in, const float *w1, const float *w2, const float *w3,
                             const float *ss1, const float *ss2, double *sums, float *y_out, float *ymax_h,
                             float *ymin_h, unsigned char *amax_h, unsigned char *amin_h, void *stream);
/* Combines the half-centre extremes of ogc_sa_chain_fwd and applies GroupNorm + ReLU of the last layer: same outputs
 * as ogc_sa_finish (out (b,c_total,m) at channel offset c_offset, optional point-major twin out_pm (b,m,c_total),
 * sel / ysel (b,c,m): winning position (255 = clamped by the ReLU) and its pre-norm value). */
OGC_API int ogc_sa_pool_finish(int b, int c, int m, const float *ymax_h, const float *ymin_h,
                               const unsigned char *amax_h, const unsigned char *amin_h, const float *scale_shift,
                               float *out, float *out_pm, int c_total, int c_offset, unsigned char *sel, float *ysel,
                               void *stream);

/* Tensor-core variant of ogc_sa_mlp_layer_fwd: tcgen05.mma kind::tf32 with the 3xTF32 split (fp32-grade), fp32
 * accumulators in TMEM, warp-specialised loader / MMA / epilogue (csrc/mlp_tc.cu).  Same arguments, except that
 * `w` is W (cout,cin) row-major (not transposed).  nsample == 64; returns OGC_ERR_UNSUPPORTED for shapes it does
 * not cover (K = cin (dense) or cin-3 (gather) must be in [8,160]) -- callers then use the SIMT entry point. */
OGC_API int ogc_sa_mlp_layer_fwd_tc(int b, int n, int m, int nsample, int cin, int cout, int gather, int last,
                                    const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                                    const float *y_prev, const float *ss_prev, const float *w, float *y,
                                    double *sums, float *ymax, float *ymin, unsigned char *amax,
                                    unsigned char *amin, void *stream);

/* Hungarian matching on the device: replaces the per-sample `iou[b].cpu().numpy()` + scipy
 * linear_sum_assignment(maximize=True) loop of match_mask_by_iou (losses/seg_loss_unsup.py:226-240).
 * inter (b,k,k) int32 from ogc_mask_contingency -> perm12 (b,k): slot of mask2 matched to each slot of mask1,
 * perm21 (b,k): the converse.  Same algorithm and tie-breaking as scipy (csrc/lsap.cuh).  k <= 32. */
OGC_API int ogc_mask_match(int b, int k, const int *inter, int *perm12, int *perm21, void *stream);

/* Host twin of the assignment routine used by ogc_mask_match (test hook: checked against scipy on CPU).
 * score (n,n) row-major doubles, maximised; col4row[i] = column matched to row i. */
OGC_API int ogc_lsap_maximize_host(int n, const double *score, int *col4row);

/* RankLoss (losses/seg_loss_unsup.py:300-314): out (b) = nuclear norm of each (n,k) mask, via the fp64 Gram
 * matrix and a Jacobi eigen-solve on the device (no cuSOLVER SVD, no host sync). */
OGC_API int ogc_mask_nuclear_norm(int b, int n, int k, const float *mask, float *out, double *gram_ws,
                                  void *stream);   /* gram_ws: caller scratch, b*32*32 doubles */

/* ogc_adam_step with the step counter and learning rate in device memory (state[0] = t, state[1] = lr), so a
 * CUDA-graph-captured training step can be replayed: t advances on the device, lr is refreshed by a host->device
 * copy before each replay.  Skipped (t unchanged) when *skip_counter != 0. */
OGC_API int ogc_adam_step_dev(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                              float *state, float beta1, float beta2, float eps, float weight_decay,
                              float grad_scale, const float *skip_counter, void *stream);

/* Tensor-core variants of ogc_sa_mlp_layer_dx / ogc_sa_mlp_layer_dw (csrc/mlp_tc_bwd.cu): identical arguments and
 * results to fp32 accuracy (3xTF32), OGC_ERR_UNSUPPORTED for shapes outside nsample == 64, rows <= 128,
 * 32 <= cout <= 256 (dx: scatter mode cout <= 128; dw: 16 <= cin <= 160). */
OGC_API int ogc_sa_mlp_layer_dx_tc(int b, int n, int m, int nsample, int cout, int cin_full, int row_off, int rows,
                                   const float *dz, const float *go, int go_ctotal, int go_coff,
                                   const unsigned char *sel, const float *y, const float *coef, const float *w,
                                   const float *y_prev, const float *ss_prev, const float *mean_rstd_prev,
                                   const float *gamma_prev, float *dz_prev, double *ab_prev, float *dgamma_prev,
                                   float *dbeta_prev, const int *idx, float *dfeat_pm, int dfeat_stride,
                                   int dfeat_off, void *stream);
/* ogc_sa_mlp_layer_dx_tc in the round-2 orientation (csrc/sa_chain_bwd.cu: positions on the MMA's M axis, dY built
 * straight from global memory into tensor memory, no shared-memory staging of activations).  Same arguments and
 * results; dense mode additionally needs `chan_sums`, a caller-ZEROED (b, prev_total, 2) fp32 workspace, and may cover
 * a slice [prev_off, prev_off + rows) of a layer of prev_total channels (0, 0 = the whole layer; w column = row_off + k,
 * y_prev / dz_prev / ss_prev / gamma_prev / dgamma_prev / dbeta_prev are those of the whole layer).
 * nsample == 64, m even, cout % 32 == 0 (<= 256), rows % 16 == 0 (<= 128; dense mode: % 32 == 0), else
 * OGC_ERR_UNSUPPORTED.  Replaces utils/nn_util.py:151-168 + pointnet2/pointnet2.py:283-294 backward (autograd). */
OGC_API int ogc_sa_chain_dx(int b, int n, int m, int nsample, int cout, int cin_full, int row_off, int rows,
                            const float *dz, const float *go, int go_ctotal, int go_coff, const unsigned char *sel,
                            const float *y, const float *coef, const float *w, const float *y_prev,
                            const float *ss_prev, const float *mean_rstd_prev, const float *gamma_prev,
                            float *dz_prev, double *ab_prev, float *dgamma_prev, float *dbeta_prev, const int *idx,
                            float *dfeat_pm, int dfeat_stride, int dfeat_off, float *chan_sums, int prev_total, int prev_off,
                            void *stream);
/* Diagnostics for ogc_sa_chain_dx: a device buffer of 8 x 32 int64; the k-th launch after this call fills
 * slot k with per-role cycle sums of one CTA; NULL = off. */
OGC_API int ogc_sa_chain_dx_debug(long long *buf);
OGC_API int ogc_sa_mlp_layer_dw_tc(int b, int n, int m, int nsample, int cout, int cin, int gather, const float *dz,
                                   const float *go, int go_ctotal, int go_coff, const unsigned char *sel,
                                   const float *y, const float *coef, const float *y_prev, const float *ss_prev,
                                   const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                                   float *dw, void *stream);

/* ogc_knn_sqrt restricted to candidates with distance <= max_dist: exact AFTER the radius clipping the callers
 * apply (QueryAndGroup pointnet2/pointnet2.py:284-286, KnnLoss losses/seg_loss_unsup.py:121-122), far cheaper in
 * sparse clouds.  Slots without a candidate inside the bound hold (+inf, 0).  Requires each query's nearest
 * neighbour to lie inside the bound (queries drawn from `known`). */
OGC_API int ogc_knn_bounded(int b, int n, int m, int k, float max_dist, const float *unknown, const float *known,
                            float *dist, int *idx, void *stream);

/* One correspondence step of object_aware_icp (oa_icp.py:64-75), without any N x N tensor:
 *   flow_out[m] = sum_n softmax_n(-|pc1[m]+flow[m] - pc2[n]| / temperature) c[m,n] pc2[n]
 *                 / max(sum_n softmax_n(.) c[m,n], 1e-10)  -  pc1[m],      c = mask1 mask2^T (oa_icp.py:57)
 * pc1, flow (b,n1,3); pc2 (b,n2,3); mask1 (b,n1,k); mask2 (b,n2,k) already permuted to mask1's slot order; k <= 16. */
OGC_API int ogc_icp_correspond(int b, int n1, int n2, int k, float temperature, const float *pc1, const float *flow,
                               const float *pc2, const float *mask1, const float *mask2, float *flow_out,
                               void *stream);

/* =====================================================================================
 * Fused feature-propagation block (csrc/fp_mlp.cu) -- replaces the torch-level stack of the reference FP
 * module (utils/pointnet2_util.py:91-120): inverse-distance weights (:98-101), three_interpolate (:103), skip
 * concat (:108-112) and SharedMLP = (Conv2d 1x1, GroupNorm(4), ReLU) x L (utils/nn_util.py:151-168) over the
 * p = n points of the finer level, forward and backward.
 * ===================================================================================== */

/* x (b,c2+c1,n) = [three_interpolate(known_feats (b,c2,m), idx, w) ; skip (b,c1,n)] with
 * w_j = (1/(sqrt(dist2_j)+1e-8)) / sum_j(.) written to weight (b,n,3).  idx, dist2 (b,n,3) from ogc_three_nn. */
OGC_API int ogc_fp_interp_concat(int b, int c2, int m, int c1, int n, const float *known_feats, const int *idx,
                                 const float *dist2, const float *skip, float *x, float *weight, void *stream);

/* One pointwise layer y (b,cout,p) = W act(x), x (b,cin,p); act = relu(ss_prev[.,0]*x + ss_prev[.,1]) or identity
 * when ss_prev == NULL (layer 0).  wt = W^T (cin,cout), any cin (streamed in chunks).  sums (b,4,2) fp64 +=
 * GroupNorm sums as ogc_sa_mlp_layer_fwd.  cout multiple of 16, <= 256. */
OGC_API int ogc_pw_mlp_layer_fwd(int b, int p, int cin, int cout, const float *x, const float *ss_prev,
                                 const float *wt, float *y, double *sums, void *stream);

/* out (b,c,p) = relu(scale*y + shift): the block's output. */
OGC_API int ogc_gn_relu_apply(int b, int c, int p, const float *y, const float *scale_shift, float *out,
                              void *stream);

/* Backward entry: dz (b,c,p) = (scale*y+shift > 0) ? dout : 0; ab / dgamma / dbeta as ogc_sa_last_stats. */
OGC_API int ogc_gn_relu_bwd_stats(int b, int c, int p, const float *dout, const float *y, const float *scale_shift,
                                  const float *mean_rstd, const float *gamma, float *dz, double *ab, float *dgamma,
                                  float *dbeta, void *stream);

/* Gradient of a layer-0 input (no norm / activation in front of it): dx[b, dx_coff + r, :] = rows
 * [row_off,row_off+rows) of W^T dY, W (cout,cin_full), dY rebuilt from dz, y, coef.  rows <= 128.
 * (The layer-to-layer and weight gradients of the FP block use ogc_sa_mlp_layer_dx / _dw with nsample = 1;
 * ogc_sa_mlp_layer_dw accepts ss_prev == NULL for a raw layer input.) */
OGC_API int ogc_pw_mlp_input_grad(int b, int p, int cout, int cin_full, int row_off, int rows, const float *dz,
                                  const float *y, const float *coef, const float *w, float *dx, int dx_ctotal,
                                  int dx_coff, void *stream);

/* =====================================================================================
 * Narrow SharedMLP layers (csrc/mlp_narrow.cu): the 32 -> 32 / 32 -> 64 layers of the first SA level
 * (models/segnet_kitti.py:27-33) in a warp-per-centre / channels-in-registers mapping.  Same operands and results
 * as ogc_sa_mlp_layer_fwd / _dx / _dw (dense, gather = 0); nsample == 64, cin (cprev) == 32, cout in {32, 64}.
 * w is W (cout,cin) row-major (NOT transposed).
 * ===================================================================================== */
/* last = 1: gamma (cout) = this layer's GroupNorm weight; only the extreme the pooling will pick (max for
 * gamma >= 0, min otherwise) is computed and written to BOTH ymax/ymin (and amax/amin). */
OGC_API int ogc_sa_mlp_narrow_fwd(int b, int m, int nsample, int cin, int cout, int last, const float *y_prev,
                                  const float *ss_prev, const float *w, const float *gamma, float *y, double *sums,
                                  float *ymax, float *ymin, unsigned char *amax, unsigned char *amin, void *stream);
OGC_API int ogc_sa_mlp_narrow_dx(int b, int m, int nsample, int cout, int cprev, const float *dz, const float *go,
                                 int go_ctotal, int go_coff, const unsigned char *sel, const float *y,
                                 const float *coef, const float *w, const float *y_prev, const float *ss_prev,
                                 const float *mean_rstd_prev, const float *gamma_prev, float *dz_prev,
                                 double *ab_prev, float *dgamma_prev, float *dbeta_prev, void *stream);
OGC_API int ogc_sa_mlp_narrow_dw(int b, int m, int nsample, int cout, int cin, const float *dz, const float *go,
                                 int go_ctotal, int go_coff, const unsigned char *sel, const float *y,
                                 const float *coef, const float *y_prev, const float *ss_prev, float *dw,
                                 void *stream);

/* Forward of a DENSE SharedMLP layer with tensor-map TMA staging (csrc/sa_fwd_tma.cu): weights stationary in tensor
 * memory, relu(GN(y_prev)) formed in place in the TMA tile (MN-major tcgen05 operand), thread = output channel in the
 * epilogue.  Outputs as the gather == 0 mode of ogc_sa_mlp_layer_fwd_tc; w is (cout, cin) row-major.
 * nsample == 64, m even, cin % 32 == 0 (<= 128), cout % 32 == 0 (<= 256). */
OGC_API int ogc_sa_fwd_tma(int b, int m, int nsample, int cin, int cout, int last, const float *y_prev, const float *ss_prev,
                           const float *w, float *y, double *sums, float *ymax, float *ymin, unsigned char *amax,
                           unsigned char *amin, void *stream);

/* Input gradient of a DENSE SharedMLP layer, channel-major with tensor-map TMA staging (csrc/sa_dx_tma.cu): the dense
 * mode of ogc_sa_mlp_layer_dx_tc (same results).  W^T stationary in tensor memory, dY formed in place in the TMA tiles of
 * y / dz (MN-major operand), thread = input channel in the epilogue, TMA store of dz_prev.  cout == 256 runs as two
 * launches over halves of the contraction.  nsample == 64, rows % 32 == 0 (<= 128), cout % 32 == 0 (<= 128, or 256). */
OGC_API int ogc_sa_dx_tma(int b, int m, int nsample, int cout, int cin_full, int row_off, int rows, const float *dz,
                          const float *go, int go_ctotal, int go_coff, const unsigned char *sel, const float *y,
                          const float *coef, const float *w, const float *y_prev, const float *ss_prev,
                          const float *mean_rstd_prev, const float *gamma_prev, float *dz_prev, double *ab_prev,
                          float *dgamma_prev, float *dbeta_prev, void *stream);

/* Weight gradient of a DENSE SharedMLP layer with tensor-map TMA staging (csrc/sa_dw_tma.cu): the stored (b,c,p) tensors
 * land in shared memory as 128-byte-swizzled K-major tcgen05 operands, are turned into dY / relu(GN(y_prev)) and their
 * TF32 residuals in place, and contracted over positions on the tensor cores (3xTF32).  Same arguments and result as
 * ogc_sa_mlp_narrow_dw.  nsample == 64, m even, cout % 32 == 0 (<= 256), cin % 32 == 0 (<= 128). */
OGC_API int ogc_sa_dw_tma(int b, int m, int nsample, int cout, int cin, const float *dz, const float *go, int go_ctotal,
                          int go_coff, const unsigned char *sel, const float *y, const float *coef, const float *y_prev,
                          const float *ss_prev, float *dw, void *stream);

/* =====================================================================================
 * Mask head (csrc/mask_head.cu) -- replaces models/segnet_kitti.py:85-88:
 *   mask (b,n,k) = softmax_k( <feats[:,n]/max(|feats[:,n]|,1e-12), slots_hat[:,k]> * inv_temperature )
 * feats (b,d,n) channel-major, slots_hat (b,d,k) ALREADY normalised over d (tiny; stays in torch).  d == 64, k <= 16.
 * Backward: dfeats (b,d,n) written, dslots_hat (b,d,k) accumulated atomically (zero it first).
 * ===================================================================================== */
OGC_API int ogc_mask_head_fwd(int b, int d, int n, int k, float inv_temperature, const float *feats,
                              const float *slots_hat, float *mask, void *stream);
OGC_API int ogc_mask_head_bwd(int b, int d, int n, int k, float inv_temperature, const float *feats,
                              const float *slots_hat, const float *mask, const float *dmask, float *dfeats,
                              float *dslots_hat, void *stream);

/* Soft correspondence transfer of the multi-frame voting (vote.py:17-28 softmax(-cdist/T) and its use corr @ mask,
 * vote.py:121; chained correspondences vote.py:50-57 by repeated application -- see csrc/icp.cu):
 *   out[m,:] = sum_n softmax_n(-|query[m] - key[n]| / temperature) val[n,:]
 * query (b,n1,3) (= pc_src + flow), key (b,n2,3), val (b,n2,k), out (b,n1,k); k <= 16.  No N x N tensor. */
OGC_API int ogc_softmax_transfer(int b, int n1, int n2, int k, float temperature, const float *query, const float *key,
                                 const float *val, float *out, void *stream);

/* =====================================================================================
 * BatchNorm shared MLP of the FlowStep3D blocks (csrc/bn_mlp.cu) -- replaces the torch-level stack of the reference's
 * PointNetSetAbstraction / FlowEmbedding MLPs (utils/flowstep3d_util.py:52-64, :126-137): per layer Conv2d 1x1,
 * BatchNorm2d in training mode (statistics over batch x npoint x nsample), ReLU; torch.max over nsample at the end.
 * The contractions are ogc_pw_mlp_layer_fwd / ogc_sa_mlp_layer_dw / ogc_sa_mlp_layer_dx / ogc_pw_mlp_input_grad (their
 * per-(sample, channel) tables hold the per-channel BatchNorm value for every sample); the entry points below make
 * the statistics, the tables, the pooling and the backward entry.  p = m * nsample positions per sample.
 * ===================================================================================== */
/* sums (c,2) fp64 += [sum y, sum y^2] over (b, p) per channel (zero it first). */
OGC_API int ogc_bn_stats(int b, int c, int p, const float *y, double *sums, void *stream);
/* scale_shift (b,c,2) = [gamma*rstd, beta - mean*gamma*rstd] for every sample, mean_rstd (c,2); count = b*p; biased
 * variance, eps 1e-5.  running_mean / running_var (c), when given, are updated like nn.BatchNorm2d does in training
 * mode: (1-momentum)*old + momentum*new with the UNBIASED batch variance. */
OGC_API int ogc_bn_finalize(int b, int c, long long count, const double *sums, const float *gamma, const float *beta,
                            float *scale_shift, float *mean_rstd, float *running_mean, float *running_var,
                            float momentum, void *stream);
/* out (b,c,m) = max over nsample of relu(scale*y+shift) (scale_shift != NULL) or of y (bare convolution block);
 * sel (b,c,m) = first winning slot, 255 = clamped by the ReLU.  nsample < 255. */
OGC_API int ogc_bn_pool(int b, int c, int m, int nsample, const float *y, const float *scale_shift, float *out,
                        unsigned char *sel, void *stream);
/* dz (b,c,p) = go (b,c,m) at the winning slot, 0 elsewhere; with mean_rstd / ab given also the last layer's
 * BatchNorm-backward sums ab (c,2) fp64 += [sum dz, sum dz*yhat] (both NULL for a bare convolution block). */
OGC_API int ogc_bn_pool_bwd(int b, int c, int m, int nsample, const float *go, const unsigned char *sel, const float *y,
                            const float *mean_rstd, float *dz, double *ab, void *stream);
/* ab (c,2) fp64 += [sum dz, sum dz*yhat] of an inner layer from its ReLU-masked dz (b,c,p) and pre-norm y. */
OGC_API int ogc_bn_bwd_stats(int b, int c, int p, const float *dz, const float *y, const float *mean_rstd, double *ab,
                             void *stream);
/* coef (b,c,4) = [gamma rstd, gamma rstd A/count, gamma rstd^2 Bx/count, mean] for every sample (the dY table of
 * ogc_sa_mlp_layer_dx / _dw); dgamma (c) += Bx, dbeta (c) += A. */
OGC_API int ogc_bn_bwd_coef(int b, int c, long long count, const double *ab, const float *mean_rstd, const float *gamma,
                            float *coef, float *dgamma, float *dbeta, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OGC_B200_H */
