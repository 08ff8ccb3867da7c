/*
 * oracle/pointnet2_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the ten kernels of the reference `pointnet2_cuda` extension
 * (vLAR-group/OGC @ 52c9836, pointnet2/src/ *.cu).  Each function states which
 * reference lines it follows.  It is the checker for the sm_100a kernels in
 * ogc_b200/csrc and the "port" CPU baseline of bench.py; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (ogc_b200, pointnet2/) never does.
 *
 * Parity status: the reference ships NO tests, golden vectors or CPU path
 * (SURVEY.md section 4, 8c), so this restatement is pinned differentially: against
 * the reference extension itself built from /root/reference into oracle/_ref and
 * run on a B200 (tests/test_gpu_vs_refext.py), and against brute-force numpy
 * definitions (tests/test_oracle.py).
 *
 * Floating point: the reference's distance expression
 *     (a-b)*(a-b) + (c-d)*(c-d) + (e-f)*(e-f)
 * is contracted by nvcc -O2 (fmad on) into  fma(dz,dz, fma(dx,dx, dy*dy))  (checked in
 * the SASS of the reference build, see oracle/ref_build.py --sass).  We spell that out
 * with fmaf() and compile with -ffp-contract=off so gcc adds no contraction of its own.
 *
 * Build: see oracle/build.py  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* fp32 squared distance with the contraction nvcc emits for the reference kernels. */
static inline float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* pointnet2/src/cuda_utils.h:10-14  opt_n_threads(): 2^floor(log2 n) clamped to [1,1024],
 * evaluated with the same double-precision log quotient. */
ORACLE_API int oracle_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* pointnet2/src/sampling_gpu.cu:86-209  furthest_point_sampling_kernel<block_size>.
 * One "block" per cloud; we simulate the block literally: every thread's strided scan
 * (:129-138, strict '>' keeps the first maximum in its stride), then the shared-memory
 * tree (:143-203) whose __update (:86-91) keeps the lower slot on ties.
 * temp must hold 1e10 on entry exactly as the Python wrapper fills it
 * (pointnet2/pointnet2.py:33); it is clobbered like in the reference. */
ORACLE_API void oracle_furthest_point_sampling(int b, int n, int m, const float *dataset,
                                               float *temp, int *idxs) {
    if (m <= 0 || n <= 0) return;
    const int bs = oracle_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi) {
        const float *pts = dataset + (size_t)bi * n * 3;
        float *tmp = temp + (size_t)bi * n;
        int *out = idxs + (size_t)bi * m;
        float *dists = (float *)malloc(sizeof(float) * bs);
        int *dists_i = (int *)malloc(sizeof(int) * bs);
        int old = 0;
        out[0] = old;
        for (int j = 1; j < m; ++j) {
            const float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < n; k += bs) {
                    float d = sqdist(pts[k * 3 + 0], pts[k * 3 + 1], pts[k * 3 + 2], x1, y1, z1);
                    float d2 = fminf(d, tmp[k]); /* CUDA min(float,float) == fminf */
                    tmp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int half = bs >> 1; half >= 1; half >>= 1) {
                for (int tid = 0; tid < half; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + half];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + half];
                    dists[tid] = fmaxf(v1, v2); /* CUDA max(float,float) */
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
        free(dists);
        free(dists_i);
    }
}

/* pointnet2/src/sampling_gpu.cu:8-24  gather_points_kernel_fast */
ORACLE_API void oracle_gather_points(int b, int c, int n, int m, const float *points,
                                     const int *idx, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            const int *id = idx + (size_t)bi * m;
            float *dst = out + ((size_t)bi * c + ci) * m;
            for (int p = 0; p < m; ++p) dst[p] = src[id[p]];
        }
}

/* pointnet2/src/sampling_gpu.cu:46-63  gather_points_grad_kernel_fast (atomicAdd scatter;
 * here summed in ascending point order -- the reference order is unspecified). */
ORACLE_API void oracle_gather_points_grad(int b, int c, int n, int m, const float *grad_out,
                                          const int *idx, float *grad_points) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *go = grad_out + ((size_t)bi * c + ci) * m;
            const int *id = idx + (size_t)bi * m;
            float *gp = grad_points + ((size_t)bi * c + ci) * n;
            for (int p = 0; p < m; ++p) gp[id[p]] += go[p];
        }
}

/* pointnet2/src/interpolate_gpu.cu:9-57  knn_kernel_fast.
 * best[] is double (1e40 sentinel), the candidate distance is fp32; strict '<' insertion
 * scanning i ascending => result is the k smallest under (d, index) lexicographic order.
 * k <= 200 as in the reference (:30-31). dist2 is written back as float ((float)1e40 = inf). */
ORACLE_API int oracle_knn(int b, int n, int m, int k, const float *unknown, const float *known,
                          float *dist2, int *idx) {
    if (k > 200 || k < 0) return -1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < n; ++p) {
            const float *u = unknown + ((size_t)bi * n + p) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            const float ux = u[0], uy = u[1], uz = u[2];
            double best[200];
            int besti[200];
            for (int i = 0; i < k; ++i) {
                best[i] = 1e40;
                besti[i] = 0;
            }
            for (int i = 0; i < m; ++i) {
                float d = sqdist(ux, uy, uz, kn[i * 3 + 0], kn[i * 3 + 1], kn[i * 3 + 2]);
                for (int j = 0; j < k; ++j) {
                    if ((double)d < best[j]) {
                        for (int l = k - 1; l > j; --l) {
                            best[l] = best[l - 1];
                            besti[l] = besti[l - 1];
                        }
                        best[j] = d;
                        besti[j] = i;
                        break;
                    }
                }
            }
            float *od = dist2 + ((size_t)bi * n + p) * k;
            int *oi = idx + ((size_t)bi * n + p) * k;
            for (int i = 0; i < k; ++i) {
                oi[i] = besti[i];
                od[i] = (float)best[i];
            }
        }
    return 0;
}

/* pointnet2/src/interpolate_gpu.cu:81-124  three_nn_kernel_fast (strict '<' cascade). */
ORACLE_API void oracle_three_nn(int b, int n, int m, const float *unknown, const float *known,
                                float *dist2, int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < n; ++p) {
            const float *u = unknown + ((size_t)bi * n + p) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            const float ux = u[0], uy = u[1], uz = u[2];
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                float d = sqdist(ux, uy, uz, kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; besti3 = besti2;
                    best2 = best1; besti2 = besti1;
                    best1 = d; besti1 = k;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2;
                    best2 = d; besti2 = k;
                } else if (d < best3) {
                    best3 = d; besti3 = k;
                }
            }
            float *od = dist2 + ((size_t)bi * n + p) * 3;
            int *oi = idx + ((size_t)bi * n + p) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
        }
}

/* pointnet2/src/interpolate_gpu.cu:149-169  three_interpolate_kernel_fast.
 * nvcc contracts w0*p0 + w1*p1 + w2*p2 into fma(w2,p2, fma(w0,p0, w1*p1)). */
ORACLE_API void oracle_three_interpolate(int b, int c, int m, int n, const float *points,
                                         const int *idx, const float *weight, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * m;
            float *dst = out + ((size_t)bi * c + ci) * n;
            for (int p = 0; p < n; ++p) {
                const float *w = weight + ((size_t)bi * n + p) * 3;
                const int *id = idx + ((size_t)bi * n + p) * 3;
                dst[p] = fmaf(w[2], src[id[2]], fmaf(w[0], src[id[0]], w[1] * src[id[1]]));
            }
        }
}

/* pointnet2/src/interpolate_gpu.cu:192-214  three_interpolate_grad_kernel_fast. */
ORACLE_API void oracle_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                              const int *idx, const float *weight,
                                              float *grad_points) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *go = grad_out + ((size_t)bi * c + ci) * n;
            float *gp = grad_points + ((size_t)bi * c + ci) * m;
            for (int p = 0; p < n; ++p) {
                const float *w = weight + ((size_t)bi * n + p) * 3;
                const int *id = idx + ((size_t)bi * n + p) * 3;
                gp[id[0]] += go[p] * w[0];
                gp[id[1]] += go[p] * w[1];
                gp[id[2]] += go[p] * w[2];
            }
        }
}

/* pointnet2/src/group_points_gpu.cu:47-66  group_points_kernel_fast. */
ORACLE_API void oracle_group_points(int b, int c, int n, int npoints, int nsample,
                                    const float *points, const int *idx, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            const int *id = idx + (size_t)bi * npoints * nsample;
            float *dst = out + ((size_t)bi * c + ci) * npoints * nsample;
            for (size_t e = 0; e < (size_t)npoints * nsample; ++e) dst[e] = src[id[e]];
        }
}

/* pointnet2/src/group_points_gpu.cu:8-25  group_points_grad_kernel_fast. */
ORACLE_API void oracle_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                         const float *grad_out, const int *idx,
                                         float *grad_points) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *go = grad_out + ((size_t)bi * c + ci) * npoints * nsample;
            const int *id = idx + (size_t)bi * npoints * nsample;
            float *gp = grad_points + ((size_t)bi * c + ci) * n;
            for (size_t e = 0; e < (size_t)npoints * nsample; ++e) gp[id[e]] += go[e];
        }
}

/* pointnet2/src/ball_query_gpu.cu:9-45  ball_query_kernel_fast.
 * radius2 = radius*radius in fp32 (:23); strict '<' (:34); the first hit fills every slot
 * (:35-39); stop at nsample (:42).  idx must be pre-zeroed by the caller
 * (pointnet2/pointnet2.py:251). */
ORACLE_API void oracle_ball_query(int b, int n, int m, float radius, int nsample,
                                  const float *new_xyz, const float *xyz, int *idx) {
    const float radius2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < m; ++p) {
            const float *q = new_xyz + ((size_t)bi * m + p) * 3;
            const float *pts = xyz + (size_t)bi * n * 3;
            int *o = idx + ((size_t)bi * m + p) * nsample;
            const float nx = q[0], ny = q[1], nz = q[2];
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist(nx, ny, nz, pts[k * 3 + 0], pts[k * 3 + 1], pts[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
}
