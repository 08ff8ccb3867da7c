// Fused set-abstraction MLP layer on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 split).
//
// Same contract as mlp_fwd_kernel in mlp.cu (one SharedMLP layer: y = W a, GroupNorm statistics, optional
// max/min over nsample), for the layers where the contraction dominates (C_out >= 64, K a multiple of 32
// after taking the three xyz input channels of a gathered layer out of the GEMM).
//
//   D[co][p] = sum_k W[co][k] a[k][p]      M = 128 output channels (TMEM lanes), N = 64 positions (one centre,
//                                          TMEM columns), K = input channels, fp32 accumulators in TMEM
//   fp32-grade accuracy from tf32 tensor cores: x = hi + lo (hi = rna_tf32(x), lo = rna_tf32(x - hi));
//   D += W_hi a_hi + W_hi a_lo + W_lo a_hi      (error ~3x an fp32 GEMM: tests/test_gpu_tcgen05.py)
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-3  epilogue: tcgen05.ld their 32 TMEM lanes (= 32 channels) x 64 columns; thread = channel, so
//              the GroupNorm sums, the max/min/arg over the 64 samples and the xyz contribution of a gathered
//              layer are plain per-thread loops (no shuffles); stores y channel-major (16 x 16 B per thread)
//   warps 4-7  loader: build the activation tile (64 positions x K) in the 128B-swizzled K-major layout,
//              hi and lo copies: gather feature rows through idx (coalesced 128 B row reads), or read the
//              previous layer's pre-norm output and apply GroupNorm+ReLU on the fly
//   warp  8    one thread issues the MMAs and commits to mbarriers; owns the TMEM allocation
// The weight tile (hi+lo, up to 128 KB) stays in shared memory for the CTA's lifetime; TMEM is double
// buffered (2 x 64 columns) so the epilogue of tile t overlaps the load + MMA of tile t+1.
#include "mlp_common.cuh"
#include "tcgen05.cuh"

namespace ogc {

constexpr int kTcLoaderWarps = 8;
constexpr int kTcLoaders = kTcLoaderWarps * 32;
constexpr int kTcMmaWarp = 4 + kTcLoaderWarps;
constexpr int kTcThreads = (kTcMmaWarp + 1) * 32;
constexpr int kTcNT = 64;          // positions per tile == nsample
constexpr int kTcM = 128;          // output channels per CTA (one M block)

struct MlpTcParams {
    MlpFwdParams f;
    const float *W;       // (Cout, Cin) row-major (NOT transposed)
    int k_off, K;         // tensor-core K range: W columns [k_off, k_off+K); gather: k_off = 3, K = Cf
};

template <bool GATHER, bool LAST>
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_fwd_tc_kernel(MlpTcParams q) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full, bar_empty, bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ double gs[kGnGroups][2];
    __shared__ float rel[2][kTcNT][4];   // centred xyz of the tile's positions (gather), double buffered

    const MlpFwdParams &f = q.f;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, mb = blockIdx.z;
    const int K = q.K, KB = (K + 31) / 32;
    const int Cout = f.Cout, Cin = f.Cin, P = f.P;
    const int ntiles = P / kTcNT;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t w_bytes = static_cast<uint32_t>(KB) * kTcM * 128u, a_bytes = static_cast<uint32_t>(KB) * kTcNT * 128u;
    uint8_t *w_hi = smem, *w_lo = w_hi + w_bytes, *a_hi = w_lo + w_bytes, *a_lo = a_hi + a_bytes;

    if (warp == kTcMmaWarp) tc::tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) {
        mbar_init(&bar_full, kTcLoaders);
        mbar_init(&bar_empty, 1);
        mbar_init(&bar_tfull[0], 1); mbar_init(&bar_tfull[1], 1);
        mbar_init(&bar_tempty[0], 128); mbar_init(&bar_tempty[1], 128);
        mbar_fence_init();
    }
    if (tid < kGnGroups * 2) (&gs[0][0])[tid] = 0.0;
    // weight tile: rows = this M block's output channels, K-major, hi + lo
    for (int e = tid; e < kTcM * KB * 32; e += kTcThreads) {
        const int r = e / (KB * 32), k = e - r * (KB * 32);
        const int co = mb * kTcM + r;
        const float v = (co < Cout && k < K) ? __ldg(q.W + static_cast<size_t>(co) * Cin + q.k_off + k) : 0.f;
        const float hi = tc::tf32_hi(v);
        const uint32_t off = static_cast<uint32_t>(k >> 5) * (kTcM * 128u) + tc::sw128_offset(r, k & 31);
        *reinterpret_cast<float *>(w_hi + off) = hi;
        *reinterpret_cast<float *>(w_lo + off) = tc::tf32_hi(v - hi);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp >= 4 && warp < kTcMmaWarp) {
        // ================================ loader ================================
        const int lt = tid - 128;           // 0..kTcLoaders-1
        const int lw = warp - 4;
        int use = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++use) {
            const int p0 = t * kTcNT;
            mbar_wait(&bar_empty, (use & 1) ^ 1);          // the MMAs that read the previous tile are done
            if (GATHER) {
                // rel[use & 1] was last read by the epilogue of tile use-2, which arrives on bar_tempty AFTER that read
                mbar_wait(&bar_tempty[use & 1], ((use >> 1) & 1) ^ 1);
                // the tile's 64 neighbour indices: lane l holds positions l and l + 32
                const int j_lo = __ldg(f.idx + static_cast<size_t>(b) * P + p0 + lane);
                const int j_hi = __ldg(f.idx + static_cast<size_t>(b) * P + p0 + 32 + lane);
                // one warp per (position, column block): 32 lanes = 32 consecutive feature channels of one point
                // row (coalesced 128 B); 8 independent row reads in flight per lane before any is consumed
                const int nitems = kTcNT * KB;
                for (int it0 = lw; it0 < nitems; it0 += kTcLoaderWarps * 8) {
                    float vals[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int it = it0 + kTcLoaderWarps * u;
                        const int p = it / KB, kb = it - p * KB;
                        const int j = __shfl_sync(OGC_FULL_MASK, p < 32 ? j_lo : j_hi, p & 31);
                        const int c = kb * 32 + lane;
                        vals[u] = (it < nitems && c < K) ? __ldg(f.feat_pm + (static_cast<size_t>(b) * f.N + j) * f.Cf + c) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int it = it0 + kTcLoaderWarps * u;
                        if (it < nitems) {
                            const int p = it / KB, kb = it - p * KB;
                            const float hi = tc::tf32_hi(vals[u]);
                            const uint32_t off = static_cast<uint32_t>(kb) * (kTcNT * 128u) + tc::sw128_offset(p, lane);
                            *reinterpret_cast<float *>(a_hi + off) = hi;
                            *reinterpret_cast<float *>(a_lo + off) = tc::tf32_hi(vals[u] - hi);
                        }
                    }
                }
                if (lt < kTcNT) {
                    const int j = __ldg(f.idx + static_cast<size_t>(b) * P + p0 + lt);
                    const int m = (p0 + lt) / f.S;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        rel[use & 1][lt][c] = __ldg(f.xyz + (static_cast<size_t>(b) * f.N + j) * 3 + c) -
                                              __ldg(f.new_xyz + (static_cast<size_t>(b) * f.M + m) * 3 + c);
                }
            } else {
                // channel-major source: item = (column block, position quad, channel in block); a warp covers the
                // 32 channels of one block for one quad -> conflict-free swizzled row writes.  8 independent 16 B
                // loads in flight per thread before any is consumed (the loader is latency-, not issue-bound).
                const int nitems = KB * 16 * 32;
                for (int it0 = lt; it0 < nitems; it0 += kTcLoaders * 8) {
                    float4 raw[8];
                    float scv[8], shv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int it = it0 + kTcLoaders * u;
                        const int cl = it & 31, pq = (it >> 5) & 15, kb = it >> 9;
                        const int c = kb * 32 + cl;
                        raw[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        scv[u] = shv[u] = 0.f;
                        if (it < nitems && c < K) {
                            scv[u] = __ldg(f.ss_prev + (static_cast<size_t>(b) * Cin + c) * 2);
                            shv[u] = __ldg(f.ss_prev + (static_cast<size_t>(b) * Cin + c) * 2 + 1);
                            raw[u] = __ldg(reinterpret_cast<const float4 *>(f.y_prev + (static_cast<size_t>(b) * Cin + c) * P + p0 + pq * 4));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int it = it0 + kTcLoaders * u;
                        if (it >= nitems) continue;
                        const int cl = it & 31, pq = (it >> 5) & 15, kb = it >> 9;
                        const float vv[4] = {fmaxf(fmaf(scv[u], raw[u].x, shv[u]), 0.f), fmaxf(fmaf(scv[u], raw[u].y, shv[u]), 0.f),
                                             fmaxf(fmaf(scv[u], raw[u].z, shv[u]), 0.f), fmaxf(fmaf(scv[u], raw[u].w, shv[u]), 0.f)};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float hi = tc::tf32_hi(vv[j]);
                            const uint32_t off = static_cast<uint32_t>(kb) * (kTcNT * 128u) + tc::sw128_offset(pq * 4 + j, cl);
                            *reinterpret_cast<float *>(a_hi + off) = hi;
                            *reinterpret_cast<float *>(a_lo + off) = tc::tf32_hi(vv[j] - hi);
                        }
                    }
                }
            }
            tc::fence_proxy_async();
            mbar_arrive(&bar_full);
        }
    } else if (warp == kTcMmaWarp) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_tf32(kTcM, kTcNT, 0, 0);
            int use = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++use) {
                const int buf = use & 1;
                mbar_wait(&bar_full, use & 1);
                mbar_wait(&bar_tempty[buf], ((use >> 1) & 1) ^ 1);
                tc::fence_after_sync();
                const uint32_t d = tmem_base + static_cast<uint32_t>(buf * kTcNT);
                uint32_t acc = 0;
                for (int s = 0; s < KB * 4; ++s) {
                    const uint32_t wo = static_cast<uint32_t>(s >> 2) * (kTcM * 128u) + static_cast<uint32_t>(s & 3) * 32u;
                    const uint32_t ao = static_cast<uint32_t>(s >> 2) * (kTcNT * 128u) + static_cast<uint32_t>(s & 3) * 32u;
                    const uint64_t whd = tc::make_desc_sw128(smem_u32(w_hi) + wo, 16, 1024);
                    const uint64_t wld = tc::make_desc_sw128(smem_u32(w_lo) + wo, 16, 1024);
                    const uint64_t ahd = tc::make_desc_sw128(smem_u32(a_hi) + ao, 16, 1024);
                    const uint64_t ald = tc::make_desc_sw128(smem_u32(a_lo) + ao, 16, 1024);
                    tc::mma_tf32(d, whd, ahd, idesc, acc);
                    tc::mma_tf32(d, whd, ald, idesc, 1);
                    tc::mma_tf32(d, wld, ahd, idesc, 1);
                    acc = 1;
                }
                tc::mma_commit(&bar_empty);        // operand tile may be overwritten
                tc::mma_commit(&bar_tfull[buf]);   // accumulator ready
            }
        }
    } else {
        // ================================ epilogue (warps 0-3) ================================
        const int co = mb * kTcM + tid;             // this thread's output channel
        const bool valid = co < Cout;
        float wx[3] = {0.f, 0.f, 0.f};
        if (GATHER && valid)
            for (int c = 0; c < 3; ++c) wx[c] = __ldg(q.W + static_cast<size_t>(co) * Cin + c);
        double ds = 0.0, dq = 0.0;
        int use = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++use) {
            const int buf = use & 1;
            const int p0 = t * kTcNT;
            mbar_wait(&bar_tfull[buf], (use >> 1) & 1);
            tc::fence_after_sync();
            float v[kTcNT];
            {
                float h[32];
                const uint32_t ta = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(buf * kTcNT);
                tc::tmem_ld32(ta, h);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = h[j];
                tc::tmem_ld32(ta + 32, h);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[32 + j] = h[j];
            }
            if (GATHER) {
                // the three xyz input channels stay on the CUDA cores
#pragma unroll
                for (int j = 0; j < kTcNT; ++j)
                    v[j] = fmaf(wx[2], rel[buf][j][2], fmaf(wx[1], rel[buf][j][1], fmaf(wx[0], rel[buf][j][0], v[j])));
            }
            tc::fence_before_sync();
            mbar_arrive(&bar_tempty[buf]);          // TMEM buffer (and rel[buf]) free for tile t+2
            if (valid) {
                float s = 0.f, sq = 0.f;
#pragma unroll
                for (int j = 0; j < kTcNT; ++j) { s += v[j]; sq = fmaf(v[j], v[j], sq); }
                ds += static_cast<double>(s);
                dq += static_cast<double>(sq);
                if (f.y) {
                    float4 *dst = reinterpret_cast<float4 *>(f.y + (static_cast<size_t>(b) * Cout + co) * P + p0);
#pragma unroll
                    for (int j = 0; j < kTcNT / 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if (LAST) {
                    float vmx = v[0], vmn = v[0];
                    int imx = 0, imn = 0;
#pragma unroll
                    for (int j = 1; j < kTcNT; ++j) {
                        if (v[j] > vmx) { vmx = v[j]; imx = j; }
                        if (v[j] < vmn) { vmn = v[j]; imn = j; }
                    }
                    const size_t o = (static_cast<size_t>(b) * Cout + co) * f.M + p0 / kTcNT;
                    f.ymax[o] = vmx; f.ymin[o] = vmn;
                    f.amax[o] = static_cast<unsigned char>(imx);
                    f.amin[o] = static_cast<unsigned char>(imn);
                }
            }
        }
        if (valid) {
            const int g = co / (Cout / kGnGroups);
            atomicAdd(&gs[g][0], ds);
            atomicAdd(&gs[g][1], dq);
        }
        named_bar_sync(1, 128);
        if (tid < kGnGroups * 2) atomicAdd(f.sums + static_cast<size_t>(b) * kGnGroups * 2 + tid, (&gs[0][0])[tid]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kTcMmaWarp) tc::tmem_dealloc(tmem_base, 128);
}

}  // namespace ogc

// Tensor-core variant of ogc_sa_mlp_layer_fwd (same meaning of every argument; `w` is W (cout,cin), not W^T).
// Supported: nsample == 64, cout a multiple of 16 with cout/4 groups aligned, K = (gather ? cin-3 : cin) a
// multiple of 4 and <= 160.  Returns OGC_ERR_UNSUPPORTED otherwise (callers fall back to the SIMT kernel).
extern "C" int ogc_sa_mlp_layer_fwd_tc(int b, int n, int m, int nsample, int cin, int cout, int gather, int last,
                                       const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                                       const float *y_prev, const float *ss_prev, const float *w, float *y,
                                       double *sums, float *ymax, float *ymin, unsigned char *amax,
                                       unsigned char *amin, void *stream) {
    using namespace ogc;
    if (b < 0 || m <= 0 || nsample <= 0 || cin <= 0 || cout <= 0 || !w || !sums) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    const int K = gather ? cin - 3 : cin;
    if (nsample != kTcNT || cout % 16 != 0 || cout > 256 || K < 8 || K > 160 || b > 65535) return OGC_ERR_UNSUPPORTED;
    if (last && (!ymax || !ymin || !amax || !amin)) return OGC_ERR_INVALID_ARG;
    if (gather && (!xyz || !new_xyz || !idx || !feat_pm)) return OGC_ERR_INVALID_ARG;
    if (!gather && (!y_prev || !ss_prev)) return OGC_ERR_INVALID_ARG;
    MlpTcParams q;
    q.f.Cin = cin; q.f.Cout = cout; q.f.P = m * nsample; q.f.S = nsample; q.f.M = m; q.f.N = n; q.f.Cf = cin - 3;
    q.f.xyz = xyz; q.f.new_xyz = new_xyz; q.f.feat_pm = feat_pm; q.f.idx = idx; q.f.y_prev = y_prev; q.f.ss_prev = ss_prev;
    q.f.Wt = nullptr; q.f.y = y; q.f.sums = sums; q.f.ymax = ymax; q.f.ymin = ymin; q.f.amax = amax; q.f.amin = amin;
    q.W = w; q.k_off = gather ? 3 : 0; q.K = K;
    const int KB = (K + 31) / 32;
    const size_t smem = static_cast<size_t>(KB) * (kTcM + kTcNT) * 128 * 2 + 1024;
    if (smem > static_cast<size_t>(kMaxSmemPerCta) - 2048) return OGC_ERR_UNSUPPORTED;
    const int ntiles = m;
    const int mblocks = (cout + kTcM - 1) / kTcM;
    int per_sample = kNumSMs / (b * mblocks);      // one CTA per SM (shared memory): never more CTAs than SMs if avoidable
    per_sample = per_sample > ntiles ? ntiles : (per_sample < 1 ? 1 : per_sample);
    dim3 grid(per_sample, b, mblocks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define OGC_TC_LAUNCH(G, L)                                                                                           \
    do {                                                                                                              \
        cudaError_t e = cudaFuncSetAttribute(mlp_fwd_tc_kernel<G, L>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                             static_cast<int>(smem));                                                 \
        if (e != cudaSuccess) return static_cast<int>(e);                                                             \
        mlp_fwd_tc_kernel<G, L><<<grid, kTcThreads, smem, st>>>(q);                                                   \
    } while (0)
    if (gather && last) OGC_TC_LAUNCH(true, true);
    else if (gather) OGC_TC_LAUNCH(true, false);
    else if (last) OGC_TC_LAUNCH(false, true);
    else OGC_TC_LAUNCH(false, false);
#undef OGC_TC_LAUNCH
    OGC_RETURN_LAUNCH_STATUS();
}
