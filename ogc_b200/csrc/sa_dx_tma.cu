// Input gradient of a dense SharedMLP layer of the set-abstraction block, channel-major with tensor-map TMA staging
// (the backward twin of sa_fwd_tma.cu).
//
//   dz_{l-1}[k][p] = [relu input of layer l-1 > 0] * sum_co W[co][k] dY[co][p]      dY = k1 dz - k2 - (y - mean) k3r
//   M = input channels k (W^T: the STATIONARY operand, hi / lo in tensor memory), N = 64 positions (one centre), K = C_out.
//
// The moving operand dY is MN-major with the channel as the K row = the stored (B,C,P) layout: the TMA tiles of y_l (and
// dz_l) with the 32-byte-atom swizzle are turned into dY and its TF32 residual elementwise IN PLACE.  A thread of the
// operand builders owns (channel row, 4 positions): GroupNorm-backward coefficients, and for the last layer the arg-max
// slot and the pooled gradient, are per-row constants -- the per-element shuffles of the positions-on-M kernel
// (sa_chain_bwd.cu) are gone.  In the epilogue a thread owns ONE input channel: ReLU mask from its row of the y_{l-1}
// tile, dz_{l-1} through a swizzled staging tile and a TMA store, and the per-channel sums (d gamma, d beta, the group
// sums of the next GroupNorm backward) in two registers, no reduction.  C_out = 256 runs as two launches over halves of
// the contraction: the first stores the raw partial sum, the second adds it back in its epilogue and finishes.  The y_{l-1} tile doubles as
// the staging tile of the output store.
// Replaces mlp_dx_tc_kernel / sa_dx_kernel for dense layers; utils/nn_util.py:151-168 autograd.
#include "mlp_dy.cuh"
#include "tcgen05.cuh"
#include "tma.cuh"
#include <cstring>

namespace ogc {
namespace dxt {

constexpr int kEpiWarps = 4, kEpi = kEpiWarps * 32;
constexpr int kSplitWarp0 = kEpiWarps, kSplitWarps = 8, kSplit = kSplitWarps * 32;
constexpr int kMmaWarp = kSplitWarp0 + kSplitWarps, kLoadWarp = kMmaWarp + 1;
constexpr int kThreads = (kLoadWarp + 1) * 32;
constexpr int kMaxStages = 4;
constexpr int kN = 64;                       // positions per tile = one centre

struct Params {
    int Kc, Ctot, co_off;                    // contraction channels of this launch, of the layer, first one
    int rows, cin_full, row_off;             // W (Ctot, cin_full); output channels = W columns [row_off, row_off + rows)
    int P, M, synth, first, final;
    const float *go;                         // (B, go_ctotal, M)
    const unsigned char *sel;                // (B, Ctot, M)
    int go_ctotal, go_coff;
    const float *coef;                       // (B, Ctot, 4)
    const float *W;
    const float *ss_prev, *mean_rstd_prev, *gamma_prev;
    double *ab_prev;
    float *dgamma_prev, *dbeta_prev;
    const float *partial;                    // second half of a split contraction: (B, rows, P) raw sums of the first half
    int stages;
    uint32_t dy_stage, yp_stage, off_yp, off_tab;
};

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <bool SYNTH>
__global__ void __launch_bounds__(kThreads, 1)
sa_dx_tma_kernel(const __grid_constant__ Params q, const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_dz,
                 const __grid_constant__ CUtensorMap tm_yp, const __grid_constant__ CUtensorMap tm_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kMaxStages], bar_ready[kMaxStages], bar_free[kMaxStages];
    __shared__ __align__(8) uint64_t bar_ypfull[kMaxStages], bar_ypfree[kMaxStages], bar_acc[2], bar_accfree[2];
    __shared__ double gs[kGnGroups * 2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int Kc = q.Kc, rows = q.rows, M = q.M, NS = q.stages;
    const int ntiles = M;                                           // one centre per tile
    const int n_my = ntiles > static_cast<int>(blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float4 *tab_cf = reinterpret_cast<float4 *>(smem + q.off_tab);              // [Kc]
    float2 *tab_sg = reinterpret_cast<float2 *>(tab_cf + Kc);                   // synth: [2 tiles][Kc]: (slot, pooled gradient)
    const uint32_t blk_dy = static_cast<uint32_t>(Kc) * 128u;                   // one [Kc][32 positions] column block
    const uint32_t blk_yp = static_cast<uint32_t>(rows) * 128u;

    if (warp == kMmaWarp) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&bar_full[i], 1); mbar_init(&bar_ready[i], kSplit); mbar_init(&bar_free[i], 1);
            mbar_init(&bar_ypfull[i], 1); mbar_init(&bar_ypfree[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc[i], 1); mbar_init(&bar_accfree[i], kEpi); }
        mbar_fence_init();
    }
    if (tid < kGnGroups * 2) gs[tid] = 0.0;
    for (int c = tid; c < Kc; c += kThreads)
        tab_cf[c] = __ldg(reinterpret_cast<const float4 *>(q.coef) + static_cast<size_t>(b) * q.Ctot + q.co_off + c);
    auto tile_of = [&](int u) { return static_cast<int>(blockIdx.x) + u * static_cast<int>(gridDim.x); };
    auto load_sg = [&](int u, int c) {
        const int m = tile_of(u);
        // (the slot stays an integer bit pattern until it is stored into the table: see sa_dw_tma.cu)
        return make_float2(__int_as_float(static_cast<int>(__ldg(q.sel + (static_cast<size_t>(b) * q.Ctot + q.co_off + c) * M + m))),
                           __ldg(q.go + (static_cast<size_t>(b) * q.go_ctotal + q.go_coff + q.co_off + c) * M + m));
    };
    auto sg_entry = [](float2 raw) { return make_float2(static_cast<float>(__float_as_int(raw.x)), raw.y); };
    if (SYNTH && n_my > 0)
        for (int c = tid; c < Kc; c += kThreads) tab_sg[c] = sg_entry(load_sg(0, c));
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t col_acc = static_cast<uint32_t>(2 * Kc);

    if (warp < kEpiWarps) {
        // ---- stationary operand: W^T[k][co] hi / lo into tensor memory, lane = input channel k ----
        const int k = warp * 32 + lane;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        for (int c0 = 0; c0 < Kc; c0 += 32) {
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float w = k < rows ? __ldg(q.W + static_cast<size_t>(q.co_off + c0 + j) * q.cin_full + q.row_off + k) : 0.f;
                hi[j] = w;
                lo[j] = w - trunc_tf32(w);
            }
            tc::tmem_st32(trow + static_cast<uint32_t>(c0), hi);
            tc::tmem_st32(trow + static_cast<uint32_t>(Kc + c0), lo);
        }
        tc::fence_before_sync();
    }
    if (warp < kEpiWarps || warp == kMmaWarp)
        asm volatile("bar.sync 2, %0;" ::"r"(kEpi + 32) : "memory");

    if (warp == kLoadWarp) {
        // ============================================ TMA loader ============================================
        if (lane == 0) {
            tma::prefetch_map(&tm_y);
            if (!SYNTH) tma::prefetch_map(&tm_dz);
            if (q.final) tma::prefetch_map(&tm_yp);
        }
        const uint32_t dy_bytes = (SYNTH ? 1u : 2u) * 2u * blk_dy, yp_bytes = 2u * blk_yp;
        for (int s = 0; s < n_my; ++s) {
            const int st = s % NS;
            const int p0 = tile_of(s) * kN;
            mbar_wait(&bar_free[st], ((s / NS) & 1) ^ 1);
            if (lane == 0) {
                uint8_t *base = smem + static_cast<size_t>(st) * q.dy_stage;
                mbar_arrive_expect_tx(&bar_full[st], dy_bytes);
                for (int j = 0; j < 2; ++j) {
                    tma::load_2d(base + static_cast<size_t>(j) * blk_dy, &tm_y, p0 + 32 * j, b * q.Ctot + q.co_off, &bar_full[st]);
                    if (!SYNTH) tma::load_2d(base + 2u * blk_dy + static_cast<size_t>(j) * blk_dy, &tm_dz, p0 + 32 * j, b * q.Ctot + q.co_off, &bar_full[st]);
                }
            }
            __syncwarp();
            if (q.final) {
                mbar_wait(&bar_ypfree[st], ((s / NS) & 1) ^ 1);
                if (lane == 0) {
                    uint8_t *base = smem + q.off_yp + static_cast<size_t>(st) * q.yp_stage;
                    mbar_arrive_expect_tx(&bar_ypfull[st], yp_bytes);
                    for (int j = 0; j < 2; ++j) tma::load_2d(base + static_cast<size_t>(j) * blk_yp, &tm_yp, p0 + 32 * j, b * rows, &bar_ypfull[st]);
                }
                __syncwarp();
            }
        }
    } else if (warp == kMmaWarp) {
        // ============================================ MMA issuer ============================================
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc_tf32(128, kN, 0, 1);           // A from tensor memory, B MN-major
        for (int s = 0; s < n_my; ++s) {
            const int st = s % NS, buf = s & 1;
            mbar_wait(&bar_accfree[buf], ((s >> 1) & 1) ^ 1);
            mbar_wait(&bar_ready[st], (s / NS) & 1);
            tc::fence_after_sync();
            const uint32_t base = smem_u32(smem + static_cast<size_t>(st) * q.dy_stage);
            const uint32_t d = tmem_base + col_acc + static_cast<uint32_t>(buf * kN);
            for (int ks = 0; ks < Kc / 8; ++ks) {
                const uint64_t bh = tc::make_desc(base + static_cast<uint32_t>(ks) * 1024u, blk_dy, 512, tc::kLayoutSw128Base32);
                const uint64_t bl = tc::make_desc(base + 2u * blk_dy + static_cast<uint32_t>(ks) * 1024u, blk_dy, 512, tc::kLayoutSw128Base32);
                const uint32_t ah = tmem_base + static_cast<uint32_t>(ks * 8), al = ah + static_cast<uint32_t>(Kc);
                tc::mma_tf32_ts_elect(d, ah, bh, idesc, ks ? 1u : 0u);
                tc::mma_tf32_ts_elect(d, ah, bl, idesc, 1u);
                tc::mma_tf32_ts_elect(d, al, bh, idesc, 1u);
            }
            tc::mma_commit_elect(&bar_free[st]);
            tc::mma_commit_elect(&bar_acc[buf]);
        }
    } else if (warp >= kSplitWarp0) {
        // ============================================ operand builders: dY and its residual, in place ============================================
        const int t = tid - kSplitWarp0 * 32;
        const int q4 = t & 7, r0 = t >> 3;
        float2 sg_next[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        for (int s = 0; s < n_my; ++s) {
            const int st = s % NS;
            uint8_t *base = smem + static_cast<size_t>(st) * q.dy_stage;
            if (SYNTH && s + 1 < n_my && q4 == 0) {      // next tile's (slot, gradient) of this thread's rows: used one tile later
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (r0 + 32 * i < Kc) sg_next[i] = load_sg(s + 1, r0 + 32 * i);
            }
            mbar_wait(&bar_full[st], (s / NS) & 1);
            const float2 *sg = tab_sg + (s & 1) * Kc;
            for (int r = r0; r < Kc; r += 32) {
                const float4 cf = tab_cf[r];
                // logical positions of this 16-byte piece inside the 32-byte-atom swizzled row: 8-float group (chunk ^ (r & 3))
                const int pos0 = (((q4 >> 1) ^ (r & 3)) << 3) + ((q4 & 1) << 2);
                float2 e = make_float2(255.f, 0.f);
                if (SYNTH) e = sg[r];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t off = static_cast<uint32_t>(j) * blk_dy + static_cast<uint32_t>(r) * 128u + static_cast<uint32_t>(q4) * 16u;
                    float4 *pv = reinterpret_cast<float4 *>(base + off), *pl = reinterpret_cast<float4 *>(base + 2u * blk_dy + off);
                    const float4 y4 = *pv;
                    float z[4];
                    if (SYNTH) {
                        const int sl = static_cast<int>(e.x) - (32 * j + pos0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) z[i] = sl == i ? e.y : 0.f;
                    } else {
                        const float4 z4 = *pl;
                        z[0] = z4.x; z[1] = z4.y; z[2] = z4.z; z[3] = z4.w;
                    }
                    const float yy[4] = {y4.x, y4.y, y4.z, y4.w};
                    float v[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[i] = fmaf(cf.x, z[i], -cf.y) - (yy[i] - cf.w) * cf.z;
                        lo[i] = v[i] - trunc_tf32(v[i]);
                    }
                    *pv = make_float4(v[0], v[1], v[2], v[3]);
                    *pl = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            tc::fence_proxy_async();
            mbar_arrive(&bar_ready[st]);
            if (SYNTH && s + 1 < n_my) {                 // the other table half was last read one tile ago by every builder
                asm volatile("bar.sync 1, %0;" ::"r"(kSplit) : "memory");
                if (q4 == 0) {
                    float2 *dst = tab_sg + ((s + 1) & 1) * Kc;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (r0 + 32 * i < Kc) dst[r0 + 32 * i] = sg_entry(sg_next[i]);
                }
                asm volatile("bar.sync 1, %0;" ::"r"(kSplit) : "memory");
            }
        }
    }
    if (warp < kEpiWarps) {
        // ============================================ epilogue: thread = input channel ============================================
        const int k = warp * 32 + lane;
        const bool valid = k < rows;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + col_acc;
        float sc = 0.f, sh = 0.f, mu = 0.f, rs = 0.f;
        if (valid && q.final) {
            sc = __ldg(q.ss_prev + (static_cast<size_t>(b) * rows + k) * 2);
            sh = __ldg(q.ss_prev + (static_cast<size_t>(b) * rows + k) * 2 + 1);
            const int g = k / (rows / kGnGroups);
            mu = __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2);
            rs = __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2 + 1);
        }
        double d0 = 0.0, d1 = 0.0;
        const uint32_t orow = static_cast<uint32_t>(k) * 128u;
        const float *prow = q.partial ? q.partial + (static_cast<size_t>(b) * rows + (valid ? k : 0)) * q.P : nullptr;
        for (int s = 0; s < n_my; ++s) {
            const int buf = s & 1, st = s % NS;
            const int p0 = tile_of(s) * kN;
            // the y_{l-1} tile of this stage doubles as the staging tile of the TMA store: it goes back to the loader once the
            // store has finished reading it (one tile later)
            uint8_t *tile = smem + q.off_yp + static_cast<size_t>(st) * q.yp_stage;
            if (tid == 0 && s > 0) {
                tma::store_wait_read<0>();
                if (q.final) mbar_arrive(&bar_ypfree[(s - 1) % NS]);
            }
            if (!q.final) asm volatile("bar.sync 3, %0;" ::"r"(kEpi) : "memory");      // staging reuse without the loader's hand-shake
            float v[64];
            if (prow) {                                  // the first half's sums: requested BEFORE the wait for this tile's MMAs
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 o = __ldg(reinterpret_cast<const float4 *>(prow + p0) + j);
                    v[4 * j] = o.x; v[4 * j + 1] = o.y; v[4 * j + 2] = o.z; v[4 * j + 3] = o.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 64; ++j) v[j] = 0.f;
            }
            mbar_wait(&bar_acc[buf], (s >> 1) & 1);
            tc::fence_after_sync();
            {
                float h[32];
                tc::tmem_ld32(trow + static_cast<uint32_t>(buf * kN), h);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += h[j];
                tc::tmem_ld32(trow + static_cast<uint32_t>(buf * kN + 32), h);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[32 + j] += h[j];
            }
            tc::fence_before_sync();
            mbar_arrive(&bar_accfree[buf]);
            float s0 = 0.f, s1 = 0.f;
            if (q.final) mbar_wait(&bar_ypfull[st], (s / NS) & 1);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (!valid) break;                   // rows beyond the layer's width do not exist in the tile
                float4 *pt = reinterpret_cast<float4 *>(tile + static_cast<uint32_t>(j >> 3) * blk_yp + orow + (static_cast<uint32_t>((j & 7) ^ (k & 7)) << 4));
                if (q.final) {
                    const float4 y4 = *pt;
                    const float yy[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float gz = fmaf(sc, yy[i], sh) > 0.f ? v[4 * j + i] : 0.f;
                        v[4 * j + i] = gz;
                        s0 += gz;
                        s1 = fmaf(gz, (yy[i] - mu) * rs, s1);
                    }
                }
                *pt = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            d0 += static_cast<double>(s0);
            d1 += static_cast<double>(s1);
            tc::fence_proxy_async();
            asm volatile("bar.sync 3, %0;" ::"r"(kEpi) : "memory");
            if (tid == 0) {
                for (int j = 0; j < 2; ++j) tma::store_2d(&tm_out, p0 + 32 * j, b * rows, tile + static_cast<size_t>(j) * blk_yp);
                tma::store_commit();
            }
        }
        if (tid == 0) tma::store_wait_all();
        if (valid && q.final && n_my > 0) {
            atomicAdd(q.dbeta_prev + k, static_cast<float>(d0));
            atomicAdd(q.dgamma_prev + k, static_cast<float>(d1));
            const int g = k / (rows / kGnGroups);
            const double gm = static_cast<double>(__ldg(q.gamma_prev + k));
            atomicAdd(&gs[2 * g], gm * d0);
            atomicAdd(&gs[2 * g + 1], gm * d1);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (q.final && tid < kGnGroups * 2 && n_my > 0) atomicAdd(q.ab_prev + static_cast<size_t>(b) * kGnGroups * 2 + tid, gs[tid]);
    if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace dxt
}  // namespace ogc

// Dense mode of ogc_sa_mlp_layer_dx_tc (same arguments minus the scatter ones, same results): dz_prev (b, rows, P) written,
// ab_prev / dgamma_prev / dbeta_prev accumulated.  The output covers ALL channels of layer l-1 (rows = its width,
// w columns [row_off, row_off + rows)).  nsample == 64, rows % 32 == 0 (<= 128), cout % 32 == 0 (<= 256).
extern "C" int ogc_sa_dx_tma(int b, int m, int nsample, int cout, int cin_full, int row_off, int rows, const float *dz,
                             const float *go, int go_ctotal, int go_coff, const unsigned char *sel, const float *y, const float *coef,
                             const float *w, const float *y_prev, const float *ss_prev, const float *mean_rstd_prev,
                             const float *gamma_prev, float *dz_prev, double *ab_prev, float *dgamma_prev, float *dbeta_prev,
                             void *stream) {
    using namespace ogc;
    using namespace ogc::dxt;
    if (b < 0 || m <= 0 || cout <= 0 || rows <= 0 || row_off < 0 || row_off + rows > cin_full || !w || !y || !coef) return OGC_ERR_INVALID_ARG;
    if (!dz && (!go || !sel)) return OGC_ERR_INVALID_ARG;
    if (!y_prev || !ss_prev || !mean_rstd_prev || !gamma_prev || !dz_prev || !ab_prev || !dgamma_prev || !dbeta_prev) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (nsample != kN || b > 65535 || rows % 32 != 0 || rows > 128 || cout % 32 != 0 || cout > 256) return OGC_ERR_UNSUPPORTED;
    const bool synth = dz == nullptr;
    const int nhalf = cout > 128 ? 2 : 1;
    if (nhalf == 2 && cout != 256) return OGC_ERR_UNSUPPORTED;
    Params q{};
    q.Kc = cout / nhalf; q.Ctot = cout; q.rows = rows; q.cin_full = cin_full; q.row_off = row_off;
    q.P = m * nsample; q.M = m; q.synth = synth;
    q.go = go; q.sel = sel; q.go_ctotal = go_ctotal; q.go_coff = go_coff; q.coef = coef; q.W = w;
    q.ss_prev = ss_prev; q.mean_rstd_prev = mean_rstd_prev; q.gamma_prev = gamma_prev; q.ab_prev = ab_prev;
    q.dgamma_prev = dgamma_prev; q.dbeta_prev = dbeta_prev;
    // shared memory: dY ring (hi | lo, two column blocks each) | y_prev ring | staging tile | tables
    const uint32_t blk_dy = static_cast<uint32_t>(q.Kc) * 128u, blk_yp = static_cast<uint32_t>(rows) * 128u;
    q.dy_stage = 4u * blk_dy;
    q.yp_stage = (2u * blk_yp + 1023u) & ~1023u;
    const uint32_t tab_bytes = static_cast<uint32_t>(q.Kc) * 16u + (synth ? 2u * q.Kc * 8u : 0u);
    const long long budget = static_cast<long long>(kMaxSmemPerCta) - 2048 - tab_bytes;
    int stages = static_cast<int>(budget / (q.dy_stage + q.yp_stage));
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return OGC_ERR_UNSUPPORTED;
    q.stages = stages;
    q.off_yp = static_cast<uint32_t>(stages) * q.dy_stage;
    q.off_tab = q.off_yp + static_cast<uint32_t>(stages) * q.yp_stage;
    const size_t smem = static_cast<size_t>(q.off_tab) + tab_bytes + 1024;
    CUtensorMap tm_y, tm_dz, tm_yp, tm_out;
    memset(&tm_dz, 0, sizeof(tm_dz));
    const uint64_t p64 = static_cast<uint64_t>(q.P);
    bool ok = tma::make_2d_f32(&tm_y, y, p64, static_cast<uint64_t>(b) * cout, 32, q.Kc, 2);
    if (ok && !synth) ok = tma::make_2d_f32(&tm_dz, dz, p64, static_cast<uint64_t>(b) * cout, 32, q.Kc, 2);
    if (ok) ok = tma::make_2d_f32(&tm_yp, y_prev, p64, static_cast<uint64_t>(b) * rows, 32, rows, 1);
    if (ok) ok = tma::make_2d_f32(&tm_out, dz_prev, p64, static_cast<uint64_t>(b) * rows, 32, rows, 1);
    if (!ok) return OGC_ERR_UNSUPPORTED;
    int per_sample = kNumSMs / b;
    per_sample = per_sample > m ? m : (per_sample < 1 ? 1 : per_sample);
    dim3 grid(per_sample, b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int h = 0; h < nhalf; ++h) {
        q.co_off = h * q.Kc; q.first = h == 0; q.final = h == nhalf - 1;
        q.partial = h == 0 ? nullptr : dz_prev;
        cudaError_t e;
        if (synth) {
            e = cudaFuncSetAttribute(sa_dx_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return static_cast<int>(e);
            sa_dx_tma_kernel<true><<<grid, kThreads, smem, st>>>(q, tm_y, tm_dz, tm_yp, tm_out);
        } else {
            e = cudaFuncSetAttribute(sa_dx_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return static_cast<int>(e);
            sa_dx_tma_kernel<false><<<grid, kThreads, smem, st>>>(q, tm_y, tm_dz, tm_yp, tm_out);
        }
    }
    OGC_RETURN_LAUNCH_STATUS();
}
