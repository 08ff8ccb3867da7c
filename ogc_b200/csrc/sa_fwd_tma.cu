// Forward of a dense SharedMLP layer of the set-abstraction block, operand staged by tensor-map TMA.
//
//   y_l[b][co][p] = sum_k W[co][k] a[k][p]       a = relu(scale y_{l-1} + shift)   (GroupNorm + ReLU of layer l-1)
//   M = C_out rows (weights: the STATIONARY operand, hi / lo in tensor memory for the CTA's life),
//   N = positions (64 = one centre, or 128), K = C_in.
//
// The moving operand a is MN-major with the channel as the K row -- the layout of the stored (B,C,P) tensor.  A TMA
// tile [channels][32 positions] with the 32-byte-atom 128-byte swizzle is what an MN-major tf32 shared-memory descriptor
// reads, so its preparation is elementwise and in place: y_{l-1} tile <- a (the tensor core truncates it to TF32: the hi
// operand), second tile <- a - trunc(a) (exact lo operand).  The accumulator row of a thread is ONE output channel over
// the tile's positions: GroupNorm statistics, the max / min / arg over a centre's 64 slots and the stores of y_l need
// no cross-thread exchange at all.
// Replaces mlp_fwd_tc_kernel (mlp_tc.cu) / narrow_fwd_kernel for dense layers; utils/nn_util.py:151-168,
// utils/pointnet2_util.py:38-42 (max over nsample).
#include "mlp_common.cuh"
#include "tcgen05.cuh"
#include "tma.cuh"
#include <cstring>

namespace ogc {
namespace fwt {

constexpr int kEpiWarps = 8, kEpi = kEpiWarps * 32, kEpiGroup = 128;     // two epilogue groups of 4 warps alternate tiles
constexpr int kSplitWarp0 = kEpiWarps, kSplitWarps = 4, kSplit = kSplitWarps * 32;
constexpr int kMmaWarp = kSplitWarp0 + kSplitWarps, kLoadWarp = kMmaWarp + 1;
constexpr int kThreads = (kLoadWarp + 1) * 32;
constexpr int kMaxStages = 8;

struct Params {
    int C, Ctot, co_off, Cin, P, M, last;
    const float *ss_prev;                // (B, Cin, 2)
    const float *W;                      // (Ctot, Cin)
    float *y;                            // (B, Ctot, P)
    double *sums;                        // (B, 4, 2) += [sum y, sum y^2] per GroupNorm group of the WHOLE layer
    float *ymax, *ymin;                  // last: (B, Ctot, M)
    unsigned char *amax, *amin;
    int stages, nb, nout;                // ring depth; 32-position column blocks per stage (2 or 4); output staging buffers
    uint32_t stage_bytes, blk_bytes, off_lo, off_tab, off_out, out_bytes;
};

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void __launch_bounds__(kThreads, 1)
sa_fwd_tma_kernel(const __grid_constant__ Params q, const __grid_constant__ CUtensorMap tm_yp, const __grid_constant__ CUtensorMap tm_y) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kMaxStages], bar_ready[kMaxStages], bar_free[kMaxStages], bar_acc[2], bar_accfree[2];
    __shared__ double gs[kGnGroups * 2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int C = q.C, Cin = q.Cin, P = q.P, NB = q.nb, N = 32 * NB, NS = q.stages;
    const int ntiles = P / N;
    const int n_my = ntiles > static_cast<int>(blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float2 *tab_ss = reinterpret_cast<float2 *>(smem + q.off_tab);              // [Cin]

    if (warp == kMmaWarp) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_ready[i], kSplit); mbar_init(&bar_free[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc[i], 1); mbar_init(&bar_accfree[i], kEpiGroup); }
        mbar_fence_init();
    }
    if (tid < kGnGroups * 2) gs[tid] = 0.0;
    for (int c = tid; c < Cin; c += kThreads)
        tab_ss[c] = __ldg(reinterpret_cast<const float2 *>(q.ss_prev) + static_cast<size_t>(b) * Cin + c);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t col_acc = static_cast<uint32_t>(2 * Cin);        // two accumulators of N columns behind the weights
    auto tile_of = [&](int u) { return static_cast<int>(blockIdx.x) + u * static_cast<int>(gridDim.x); };

    if (warp < 4) {
        // ---- stationary operand: W[co][:] hi / lo into tensor memory, lane = output channel ----
        const int co = warp * 32 + lane;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        for (int k0 = 0; k0 < Cin; k0 += 32) {
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float w = co < C ? __ldg(q.W + static_cast<size_t>(q.co_off + co) * Cin + k0 + j) : 0.f;
                hi[j] = w;
                lo[j] = w - trunc_tf32(w);
            }
            tc::tmem_st32(trow + static_cast<uint32_t>(k0), hi);
            tc::tmem_st32(trow + static_cast<uint32_t>(Cin + k0), lo);
        }
        tc::fence_before_sync();
    }
    if (warp < 4 || warp == kMmaWarp)
        asm volatile("bar.sync 2, %0;" ::"r"(kEpiGroup + 32) : "memory");  // first epilogue group + the MMA warp: weights are in place

    if (warp == kLoadWarp) {
        // ============================================ TMA loader ============================================
        if (lane == 0) tma::prefetch_map(&tm_yp);
        const uint32_t bytes = q.blk_bytes * static_cast<uint32_t>(NB);
        for (int s = 0; s < n_my; ++s) {
            const int st = s % NS;
            mbar_wait(&bar_free[st], ((s / NS) & 1) ^ 1);
            if (lane == 0) {
                uint8_t *base = smem + static_cast<size_t>(st) * q.stage_bytes;
                const int p0 = tile_of(s) * N;
                mbar_arrive_expect_tx(&bar_full[st], bytes);
                for (int j = 0; j < NB; ++j) tma::load_2d(base + static_cast<size_t>(j) * q.blk_bytes, &tm_yp, p0 + 32 * j, b * Cin, &bar_full[st]);
            }
            __syncwarp();
        }
    } else if (warp == kMmaWarp) {
        // ============================================ MMA issuer ============================================
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc_tf32(128, N, 0, 1);            // A from tensor memory (K-major), B MN-major
        for (int s = 0; s < n_my; ++s) {
            const int st = s % NS, buf = s & 1;
            mbar_wait(&bar_accfree[buf], ((s >> 1) & 1) ^ 1);
            mbar_wait(&bar_ready[st], (s / NS) & 1);
            tc::fence_after_sync();
            const uint32_t base = smem_u32(smem + static_cast<size_t>(st) * q.stage_bytes);
            const uint32_t d = tmem_base + col_acc + static_cast<uint32_t>(buf * N);
            for (int ks = 0; ks < Cin / 8; ++ks) {
                const uint64_t ba = tc::make_desc(base + static_cast<uint32_t>(ks) * 1024u, q.blk_bytes, 512, tc::kLayoutSw128Base32);
                const uint64_t bl = tc::make_desc(base + q.off_lo + static_cast<uint32_t>(ks) * 1024u, q.blk_bytes, 512, tc::kLayoutSw128Base32);
                const uint32_t ah = tmem_base + static_cast<uint32_t>(ks * 8), al = ah + static_cast<uint32_t>(Cin);
                tc::mma_tf32_ts_elect(d, ah, ba, idesc, ks ? 1u : 0u);
                tc::mma_tf32_ts_elect(d, ah, bl, idesc, 1u);
                tc::mma_tf32_ts_elect(d, al, ba, idesc, 1u);
            }
            tc::mma_commit_elect(&bar_free[st]);
            tc::mma_commit_elect(&bar_acc[buf]);
        }
    } else if (warp >= kSplitWarp0) {
        // ============================================ splitters: a and its residual, elementwise in place ============================================
        const int t = tid - kSplitWarp0 * 32;
        const int q4 = t & 7, r0 = t >> 3;
        for (int s = 0; s < n_my; ++s) {
            const int st = s % NS;
            uint8_t *base = smem + static_cast<size_t>(st) * q.stage_bytes;
            mbar_wait(&bar_full[st], (s / NS) & 1);
            for (int r = r0; r < Cin; r += kSplit / 8) {
                const float2 ss = tab_ss[r];
                for (int j = 0; j < NB; ++j) {
                    const uint32_t off = static_cast<uint32_t>(j) * q.blk_bytes + static_cast<uint32_t>(r) * 128u + static_cast<uint32_t>(q4) * 16u;
                    float4 *pa = reinterpret_cast<float4 *>(base + off), *pl = reinterpret_cast<float4 *>(base + q.off_lo + off);
                    const float4 y4 = *pa;
                    const float a0 = fmaxf(fmaf(ss.x, y4.x, ss.y), 0.f), a1 = fmaxf(fmaf(ss.x, y4.y, ss.y), 0.f);
                    const float a2 = fmaxf(fmaf(ss.x, y4.z, ss.y), 0.f), a3 = fmaxf(fmaf(ss.x, y4.w, ss.y), 0.f);
                    *pa = make_float4(a0, a1, a2, a3);
                    *pl = make_float4(a0 - trunc_tf32(a0), a1 - trunc_tf32(a1), a2 - trunc_tf32(a2), a3 - trunc_tf32(a3));
                }
            }
            tc::fence_proxy_async();
            mbar_arrive(&bar_ready[st]);
        }
    }
    if (warp < kEpiWarps) {
        // ============================================ epilogue: thread = output channel ============================================
        const int eg = warp >> 2, ngroups = q.nout;      // group eg takes the tiles s % ngroups == eg (its own accumulator + staging buffer)
        const int co = (warp & 3) * 32 + lane;
        const bool valid = co < C;
        const bool leader = (warp & 3) == 0 && lane == 0;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + col_acc;
        double dsum = 0.0, dsq = 0.0;
        const size_t row = static_cast<size_t>(b) * q.Ctot + q.co_off + co;
        for (int s = eg; s < n_my && eg < ngroups; s += ngroups) {
            const int buf = s & 1;
            const int p0 = tile_of(s) * N;
            uint8_t *obuf = smem + q.off_out + static_cast<size_t>(eg) * q.out_bytes;
            if (leader) tma::store_wait_read<0>();       // this group's previous TMA store has finished reading the staging buffer
            asm volatile("bar.sync %0, %1;" ::"r"(3 + eg), "r"(kEpiGroup) : "memory");
            mbar_wait(&bar_acc[buf], (s >> 1) & 1);
            tc::fence_after_sync();
            for (int cen = 0; cen < N / 64; ++cen) {
                float v[64];
                {
                    float h[32];
                    tc::tmem_ld32(trow + static_cast<uint32_t>(buf * N + cen * 64), h);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = h[j];
                    tc::tmem_ld32(trow + static_cast<uint32_t>(buf * N + cen * 64 + 32), h);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[32 + j] = h[j];
                }
                if (cen == N / 64 - 1) {
                    tc::fence_before_sync();
                    mbar_arrive(&bar_accfree[buf]);
                }
                float su = 0.f, sq = 0.f;
#pragma unroll
                for (int j = 0; j < 64; ++j) { su += v[j]; sq = fmaf(v[j], v[j], sq); }
                dsum += static_cast<double>(su);
                dsq += static_cast<double>(sq);
                // the row goes into the 128-byte-swizzled staging tile (column blocks [128 rows][32 positions]); per quarter-warp
                // the 8 rows hit 8 different 16-byte chunks: conflict-free
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int blk = cen * 2 + (j >> 3), c = j & 7;
                    *reinterpret_cast<float4 *>(obuf + static_cast<uint32_t>(blk) * (128u * 128u) + static_cast<uint32_t>(co) * 128u +
                                                (static_cast<uint32_t>(c ^ (co & 7)) << 4)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if (!valid) continue;
                if (q.last) {
                    float mx = v[0], mn = v[0];
                    int ax = 0, an = 0;
#pragma unroll
                    for (int j = 1; j < 64; ++j) {
                        if (v[j] > mx) { mx = v[j]; ax = j; }
                        if (v[j] < mn) { mn = v[j]; an = j; }
                    }
                    const size_t o = row * q.M + p0 / 64 + cen;
                    q.ymax[o] = mx; q.ymin[o] = mn;
                    q.amax[o] = static_cast<unsigned char>(ax); q.amin[o] = static_cast<unsigned char>(an);
                }
            }
            tc::fence_proxy_async();            // generic writes of the staging tile -> the TMA store's reads
            asm volatile("bar.sync %0, %1;" ::"r"(3 + eg), "r"(kEpiGroup) : "memory");
            if (leader) {
                for (int j = 0; j < NB; ++j)
                    tma::store_2d(&tm_y, p0 + 32 * j, b * q.Ctot + q.co_off, obuf + static_cast<size_t>(j) * (128u * 128u));
                tma::store_commit();
            }
        }
        if (leader) tma::store_wait_all();
        if (valid && n_my > 0) {
            const int g = (q.co_off + co) / (q.Ctot / kGnGroups);
            atomicAdd(&gs[2 * g], dsum);
            atomicAdd(&gs[2 * g + 1], dsq);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (tid < kGnGroups * 2 && n_my > 0) atomicAdd(q.sums + static_cast<size_t>(b) * kGnGroups * 2 + tid, gs[tid]);
    if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace fwt
}  // namespace ogc

// Dense-layer forward: the same outputs as the gather == 0 mode of ogc_sa_mlp_layer_fwd_tc (y, sums, and for last != 0
// ymax / ymin / amax / amin).  w is (cout, cin) row-major.  nsample == 64, m even, cin % 32 == 0 (<= 128),
// cout % 32 == 0 (<= 256; 256 runs as two 128-row launches); OGC_ERR_UNSUPPORTED otherwise.
extern "C" int ogc_sa_fwd_tma(int b, int m, int nsample, int cin, int cout, int last, const float *y_prev, const float *ss_prev,
                              const float *w, float *y, double *sums, float *ymax, float *ymin, unsigned char *amax,
                              unsigned char *amin, void *stream) {
    using namespace ogc;
    using namespace ogc::fwt;
    if (b < 0 || m <= 0 || cin <= 0 || cout <= 0 || !y_prev || !ss_prev || !w || !y || !sums) return OGC_ERR_INVALID_ARG;
    if (last && (!ymax || !ymin || !amax || !amin)) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (nsample != 64 || (m & 1) || b > 65535 || cin % 32 != 0 || cin > 128 || cout % 32 != 0 || cout > 256 || (cout / kGnGroups) == 0)
        return OGC_ERR_UNSUPPORTED;
    Params q{};
    q.Ctot = cout; q.Cin = cin; q.P = m * nsample; q.M = m; q.last = last;
    q.ss_prev = ss_prev; q.W = w; q.y = y; q.sums = sums; q.ymax = ymax; q.ymin = ymin; q.amax = amax; q.amin = amin;
    q.nb = (cin <= 32 || (cin <= 64 && cout <= 64)) ? 4 : 2;   // 128-position tiles for the small layers (per-stage hand-shake amortised)
    q.blk_bytes = static_cast<uint32_t>(cin) * 128u;
    q.off_lo = q.blk_bytes * static_cast<uint32_t>(q.nb);
    q.stage_bytes = 2 * q.off_lo;
    const uint32_t tab_bytes = static_cast<uint32_t>(cin) * 8u;
    q.out_bytes = static_cast<uint32_t>(q.nb) * 128u * 128u;                 // nb column blocks of [128 rows][128 B]
    const long long budget = static_cast<long long>(kMaxSmemPerCta) - 2048 - tab_bytes;
    q.nout = budget - 2ll * q.out_bytes >= 3ll * q.stage_bytes ? 2 : 1;
    int stages = static_cast<int>((budget - static_cast<long long>(q.nout) * q.out_bytes) / q.stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return OGC_ERR_UNSUPPORTED;
    q.stages = stages;
    q.off_out = static_cast<uint32_t>(stages) * q.stage_bytes;
    q.off_tab = q.off_out + static_cast<uint32_t>(q.nout) * q.out_bytes;
    const size_t smem = static_cast<size_t>(q.off_tab) + tab_bytes + 1024;
    CUtensorMap tm_yp;
    if (!tma::make_2d_f32(&tm_yp, y_prev, static_cast<uint64_t>(q.P), static_cast<uint64_t>(b) * cin, 32, cin, 2)) return OGC_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(sa_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    const int ntiles = q.P / (32 * q.nb);
    int per_sample = kNumSMs / b;
    per_sample = per_sample > ntiles ? ntiles : (per_sample < 1 ? 1 : per_sample);
    dim3 grid(per_sample, b);
    for (int off = 0; off < cout; off += 128) {
        q.co_off = off;
        q.C = cout - off < 128 ? cout - off : 128;
        CUtensorMap tm_y;          // stores of [C rows][32 positions] column blocks, 128-byte swizzle
        if (!tma::make_2d_f32(&tm_y, y, static_cast<uint64_t>(q.P), static_cast<uint64_t>(b) * cout, 32, q.C, 1)) return OGC_ERR_UNSUPPORTED;
        sa_fwd_tma_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(q, tm_yp, tm_y);
    }
    OGC_RETURN_LAUNCH_STATUS();
}
