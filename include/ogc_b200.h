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
 *   REQUIRED (and then clobbered) when n > 16384; otherwise it may be NULL and is untouched.
 *   It need not be pre-filled with 1e10.
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

#ifdef __cplusplus
}
#endif
#endif /* OGC_B200_H */
