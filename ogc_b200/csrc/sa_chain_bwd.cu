// Backward of a SharedMLP layer of the set-abstraction block in the round-2 orientation (positions on the MMA's M axis,
// see sa_chain.cuh): input gradient.
//
//   dA[p][k] = sum_co dY[p][co] W[co][k]        M = 128 positions, N = rows (input channels k), K = C_out
//   dY[p][co] = k1 dz - k2 - (y - mean) k3r     (GroupNorm backward folded into per-(sample, channel) coefficients,
//                                                mlp_dy.cuh; for the last layer dz is synthesised from the pooled
//                                                gradient and the recorded arg-max position)
//   dz_prev[p][k] = dA[p][k] * [relu input of layer l-1 > 0]            (dense layers), or
//   dfeat[idx[p]][k] += dA[p][k]                                          (first layer: scatter into the point-major
//                                                                          feature gradient, group_points_grad)
//
// Same skeleton as the chained forward (sa_chain_fwd.cu): producer warps (two per lane quadrant, alternating 32-channel
// K chunks) build dY for THEIR position straight from global memory -- for a fixed channel the 32 lanes read 32
// consecutive positions of the channel-major tensors, one 128-byte segment per instruction, no staging -- split it and
// store the A operand into tensor memory; one warp issues the MMAs chunk by chunk (B = W^T resident in shared memory);
// epilogue warps read the accumulator row of their position, apply the ReLU mask, store dz_prev the same coalesced
// way, and reduce the per-channel GroupNorm-backward sums through a per-warp shared-memory transpose.
// Replaces mlp_dx_tc_kernel (mlp_tc_bwd.cu) for C_out <= 256, rows <= 128; utils/nn_util.py:151-168 autograd.
#include "mlp_dy.cuh"
#include "sa_chain.cuh"
#include "tma.cuh"
#include <cstring>

namespace ogc {
namespace chain {

#ifndef OGC_DX_TIMELINE
#define OGC_DX_TIMELINE 0       // 1 (OGC_NVCC_FLAGS=-DOGC_DX_TIMELINE=1): compile the per-role cycle counters in
#endif
constexpr bool kRec = OGC_DX_TIMELINE != 0;

constexpr int kLoadWarpP = kMmaWarp + 1;       // warp 17: bulk-copy loader of the dY inputs (y_l, dz_l rows)
constexpr int kLoadWarpE = kMmaWarp + 2;       // warp 18: bulk-copy loader of the ReLU-mask inputs (y_{l-1} rows)
constexpr int kDxThreads = (kLoadWarpE + 1) * 32;
constexpr int kMaxPStages = 12, kMaxEStages = 4;
constexpr uint32_t kRowBytes = kTile * 4;      // one channel row of a tile: 128 positions, 512 B, contiguous in HBM

struct DxParams {
    DySrc dy;                      // layer l: dy.C = C_out = contraction length
    int cin_full, row_off, rows;   // W (C_out, cin_full); output columns = W columns [row_off, row_off + rows)
    const float *W;
    const float *y_prev, *ss_prev, *mean_rstd_prev;     // dense: (B,rows,P), (B,rows,2), (B,4,2)
    float *dz_prev;                                     // (B,rows,P)
    float *chan_sums;                                   // (B,rows,2): sum_p dz_prev, sum_p dz_prev * xhat_prev
    const int *idx;                                     // scatter: (B,M,64)
    float *dfeat_pm;
    int N, dfeat_stride, dfeat_off;
    int prev_total, prev_off;      // dense mode on a slice of a wider layer: channels [prev_off, prev_off + rows) of prev_total
    uint32_t off_w, off_tab, off_scr, off_pring, off_ering;
    uint32_t col_acc[2];
    long long *dbg;                // optional: per-role cycle accumulators of CTA (0,0), see ogc_sa_chain_dx_debug
    int ring;                      // A-operand ring slots in tensor memory (32-channel chunks, hi + lo = 64 columns each)
    int p_stages, e_stages;        // shared-memory rings: 16-channel half-chunks of (y_l [, dz_l]); 32-channel chunks of y_{l-1}
};

template <bool SYNTH, bool SCATTER>
__global__ void __launch_bounds__(kDxThreads, 1)
sa_dx_kernel(const __grid_constant__ DxParams q, const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_dz,
             const __grid_constant__ CUtensorMap tm_yp) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_acc[2], bar_accfree[2], bar_kfull[8], bar_kfree[8];
    __shared__ __align__(8) uint64_t bar_pfull[kMaxPStages], bar_pfree[kMaxPStages], bar_efull[kMaxEStages], bar_efree[kMaxEStages];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int C = q.dy.C, P = q.dy.P, M = q.dy.M, rows = q.rows;
    const int ntiles = M / 2;
    const int n_my = ntiles > static_cast<int>(blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nchunks = C >> 5, ring = q.ring, PS = q.p_stages, ES = q.e_stages;
    constexpr uint32_t kPStageBytes = (SYNTH ? 1u : 2u) * 16u * kRowBytes;      // 16 channels of y_l (+ 16 of dz_l)
    constexpr uint32_t kEStageBytes = 32u * kRowBytes;                           // 32 channels of y_{l-1}

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float4 *tab_cf = reinterpret_cast<float4 *>(smem + q.off_tab);              // [C]: k1, k2, k3r, mean of layer l
    float4 *tab_ss = tab_cf + C;                                                // [rows]: scale, shift, mean, rstd of layer l-1
    float *csum = reinterpret_cast<float *>(tab_ss + rows);                     // [8 epilogue warps][rows]: per-warp channel sums
    float *scr = reinterpret_cast<float *>(smem + q.off_scr);                   // scatter: [8 epilogue warps][32][33]
    uint8_t *pring = smem + q.off_pring, *ering = smem + q.off_ering;

    if (warp == kMmaWarp) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc[i], 1); mbar_init(&bar_accfree[i], kEpi); }
        for (int c = 0; c < 8; ++c) { mbar_init(&bar_kfull[c], 256); mbar_init(&bar_kfree[c], 1); }
        for (int i = 0; i < kMaxPStages; ++i) { mbar_init(&bar_pfull[i], 1); mbar_init(&bar_pfree[i], 256); }
        for (int i = 0; i < kMaxEStages; ++i) { mbar_init(&bar_efull[i], 1); mbar_init(&bar_efree[i], kEpi); }
        mbar_fence_init();
    }
    // resident B operand: rows k < rows, K index = co: element (k, co) = W[co][row_off + k]
    {
        const uint32_t blk = 2u * static_cast<uint32_t>(rows) * 128u;
        for (int e = tid; e < rows * C; e += kDxThreads) {
            const int co = e / rows, k = e - co * rows;                        // consecutive threads: consecutive W columns
            const float v = __ldg(q.W + static_cast<size_t>(co) * q.cin_full + q.row_off + k);
            const float hi = tc::tf32_hi(v), lo = tc::tf32_hi(v - hi);
            const uint32_t off = static_cast<uint32_t>(co >> 5) * blk + tc::sw128_offset(k, co & 31);
            *reinterpret_cast<float *>(smem + q.off_w + off) = hi;
            *reinterpret_cast<float *>(smem + q.off_w + off + static_cast<uint32_t>(rows) * 128u) = lo;
        }
    }
    for (int c = tid; c < C; c += kDxThreads)
        tab_cf[c] = __ldg(reinterpret_cast<const float4 *>(q.dy.coef) + static_cast<size_t>(b) * C + c);
    if (!SCATTER) {
        for (int c = tid; c < rows; c += kDxThreads) {
            const float2 s2 = __ldg(reinterpret_cast<const float2 *>(q.ss_prev) + static_cast<size_t>(b) * q.prev_total + q.prev_off + c);
            const int g = (q.prev_off + c) / (q.prev_total / kGnGroups);
            tab_ss[c] = make_float4(s2.x, s2.y, __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2),
                                    __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2 + 1));
        }
    }
    if (!SCATTER)
        for (int c = tid; c < kEpiWarps * rows; c += kDxThreads) csum[c] = 0.f;
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    auto tile_of = [&](int u) { return static_cast<int>(blockIdx.x) + u * static_cast<int>(gridDim.x); };

    if (warp == kLoadWarpP) {
        // ============================================ loader of the dY inputs ============================================
        // one TMA tile per tensor and half-chunk: [16 channels][128 positions] of y_l (and of dz_l).  The loop is
        // warp-uniform (every lane waits, lane 0 issues): the warp must reach the final __syncthreads converged.
        if (lane == 0) {
            tma::prefetch_map(&tm_y);
            if (!SYNTH) tma::prefetch_map(&tm_dz);
        }
        int hseq = 0;
        for (int u = 0; u < n_my; ++u) {
            const int pos0 = tile_of(u) * kTile;
            for (int h = 0; h < 2 * nchunks; ++h, ++hseq) {
                const int stage = hseq % PS;
                mbar_wait(&bar_pfree[stage], (((hseq / PS) & 1) ^ 1));
                if (lane == 0) {
                    uint8_t *dst = pring + static_cast<size_t>(stage) * kPStageBytes;
                    mbar_arrive_expect_tx(&bar_pfull[stage], kPStageBytes);
                    tma::load_2d(dst, &tm_y, pos0, b * C + 16 * h, &bar_pfull[stage]);
                    if (!SYNTH) tma::load_2d(dst + 16u * kRowBytes, &tm_dz, pos0, b * C + 16 * h, &bar_pfull[stage]);
                }
                __syncwarp();
            }
        }
    } else if (warp == kLoadWarpE) {
        // ============================================ loader of the ReLU-mask inputs ============================================
        if (!SCATTER) {
            if (lane == 0) tma::prefetch_map(&tm_yp);
            const int nech = rows >> 5;
            int eseq = 0;
            for (int u = 0; u < n_my; ++u) {
                const int pos0 = tile_of(u) * kTile;
                for (int i = 0; i < nech; ++i, ++eseq) {
                    const int stage = eseq % ES;
                    mbar_wait(&bar_efree[stage], (((eseq / ES) & 1) ^ 1));
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&bar_efull[stage], kEStageBytes);
                        tma::load_2d(ering + static_cast<size_t>(stage) * kEStageBytes, &tm_yp, pos0, b * q.prev_total + q.prev_off + 32 * i, &bar_efull[stage]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= kProdWarp0 && warp < kMmaWarp) {
        // ============================================ producer: dY chunks ============================================
        // All 8 warps consume every ring stage IN ORDER (lane quadrant = warp & 3, channel half = pg): with a single
        // logical consumer a waiter is never two phases ahead of a stage's barrier (two groups on alternating chunks
        // were: the parity test of a wait for use k passes while use k-1 is still in flight).
        const int pw = warp & 3, pg = (warp - kProdWarp0) >> 2;
        const int pt = pw * 32 + lane;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(pw * 32) << 16);
        const bool rec = kRec && q.dbg && blockIdx.x == 0 && blockIdx.y == 0 && warp == kProdWarp0 && lane == 0;
        long long a_kfree = 0, a_pfull = 0, a_work = 0, t_begin = clock64(), t0 = 0, t1 = 0;
        const int s_own = pt & 63;
        // one 32-channel chunk: wait for the tensor-memory slot, then the two half-chunk stages
        auto do_chunk = [&](int u, int c, int sel_l, float go_l) {
            const int chunk_seq = u * nchunks + c;             // position in the CTA's chunk sequence
            const int slot = chunk_seq % ring;
            const uint32_t use = static_cast<uint32_t>(chunk_seq / ring);      // how many times the slot was used before
            if (rec) t0 = clock64();
            mbar_wait(&bar_kfree[slot], (use & 1) ^ 1);
            tc::fence_after_sync();
            if (rec) { t1 = clock64(); a_kfree += t1 - t0; }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int hseq = chunk_seq * 2 + hh;
                const int stage = hseq % PS;
                if (rec) t0 = clock64();
                mbar_wait(&bar_pfull[stage], (hseq / PS) & 1);
                if (rec) { t1 = clock64(); a_pfull += t1 - t0; }
                const float *st_y = reinterpret_cast<const float *>(pring + static_cast<size_t>(stage) * kPStageBytes) + (8 * pg) * kTile + pt;
                float yv[8], zv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) yv[j] = st_y[j * kTile];
                if (SYNTH) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int sl = __shfl_sync(OGC_FULL_MASK, sel_l, 16 * hh + 8 * pg + j);
                        const float g = __shfl_sync(OGC_FULL_MASK, go_l, 16 * hh + 8 * pg + j);
                        zv[j] = sl == s_own ? g : 0.f;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) zv[j] = st_y[(16 + j) * kTile];
                }
                float hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 cf = tab_cf[32 * c + 16 * hh + 8 * pg + j];
                    const float v = fmaf(cf.x, zv[j], -cf.y) - (yv[j] - cf.w) * cf.z;
                    tc::tf32_split(v, hi[j], lo[j]);
                }
                // release the stage only after the loaded values have been CONSUMED: an arrive issued right behind the
                // shared-memory loads is not ordered behind their completion (measured: the next TMA tile overwrote rows
                // a warp was still reading, once per few hundred launches)
                mbar_arrive(&bar_pfree[stage]);
                tc::tmem_st8_nowait(trow + slot * 64 + 16 * hh + 8 * pg, hi);
                tc::tmem_st8_nowait(trow + slot * 64 + 32 + 16 * hh + 8 * pg, lo);
                if (rec) a_work += clock64() - t1;
            }
            tc::tmem_st_wait();
            tc::fence_before_sync();
            mbar_arrive(&bar_kfull[slot]);
        };
        if (SYNTH) {
            // lane j of chunk c: arg-max slot and pooled gradient of channel 32 c + j at this warp's centre.  The values of
            // a chunk are reloaded for the NEXT tile as soon as they are used: a whole tile of prefetch distance.
            const unsigned char *selb = q.dy.sel + (static_cast<size_t>(b) * C + lane) * M + (pt >> 6);
            const float *gob = q.dy.go + (static_cast<size_t>(b) * q.dy.go_ctotal + q.dy.go_coff + lane) * M + (pt >> 6);
            int sel_a[kMaxC / 32];
            float go_a[kMaxC / 32];
#pragma unroll
            for (int c = 0; c < kMaxC / 32; ++c) {
                sel_a[c] = 255; go_a[c] = 0.f;
                if (c < nchunks && n_my > 0) {
                    sel_a[c] = __ldg(selb + static_cast<size_t>(32 * c) * M + 2 * tile_of(0));
                    go_a[c] = __ldg(gob + static_cast<size_t>(32 * c) * M + 2 * tile_of(0));
                }
            }
            for (int u = 0; u < n_my; ++u) {
                const bool more = u + 1 < n_my;
                const int m_next = 2 * tile_of(more ? u + 1 : u);
#pragma unroll
                for (int c = 0; c < kMaxC / 32; ++c) {
                    if (c < nchunks) {
                        const int sl = sel_a[c];
                        const float g = go_a[c];
                        if (more) {
                            sel_a[c] = __ldg(selb + static_cast<size_t>(32 * c) * M + m_next);
                            go_a[c] = __ldg(gob + static_cast<size_t>(32 * c) * M + m_next);
                        }
                        do_chunk(u, c, sl, g);
                    }
                }
            }
        } else {
            for (int u = 0; u < n_my; ++u)
                for (int c = 0; c < nchunks; ++c) do_chunk(u, c, 0, 0.f);
        }
        if (rec) { q.dbg[0] = a_pfull; q.dbg[1] = a_kfree; q.dbg[2] = a_work; q.dbg[3] = clock64() - t_begin; }
    } else if (warp == kMmaWarp) {
        // ============================================ MMA issuer (warp-uniform) ============================================
        const uint32_t idesc = tc::make_idesc_tf32(kTile, rows, 0, 0);
        const uint32_t blk16 = (2u * static_cast<uint32_t>(rows) * 128u) >> 4, lo16 = (static_cast<uint32_t>(rows) * 128u) >> 4;
        const uint64_t d0 = tc::make_desc_sw128(smem_u32(smem + q.off_w), 16, 1024);
        int chunk_seq = 0;
        const bool rec = kRec && q.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
        long long a_accfree = 0, a_kfull = 0, a_issue = 0, t_begin = clock64(), t0 = 0, t1 = 0;
        for (int u = 0; u < n_my; ++u) {
            const int buf = u & 1;
            if (rec) t0 = clock64();
            mbar_wait(&bar_accfree[buf], ((u >> 1) & 1) ^ 1);
            if (rec) a_accfree += clock64() - t0;
            const uint32_t d = tmem_base + q.col_acc[buf];
            for (int c = 0; c < nchunks; ++c, ++chunk_seq) {
                const int slot = chunk_seq % ring;
                const uint32_t use = static_cast<uint32_t>(chunk_seq / ring);
                if (rec) t0 = clock64();
                mbar_wait(&bar_kfull[slot], use & 1);
                tc::fence_after_sync();
                if (rec) t1 = clock64();
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const uint64_t bh = d0 + (static_cast<uint32_t>(c) * blk16 + static_cast<uint32_t>(s) * 2u), bl = bh + lo16;
                    const uint32_t ah = tmem_base + static_cast<uint32_t>(slot * 64 + s * 8), al = ah + 32u;
                    tc::mma_tf32_ts_elect(d, ah, bh, idesc, (c | s) ? 1u : 0u);
                    tc::mma_tf32_ts_elect(d, ah, bl, idesc, 1u);
                    tc::mma_tf32_ts_elect(d, al, bh, idesc, 1u);
                }
                tc::mma_commit_elect(&bar_kfree[slot]);
                if (rec) { a_kfull += t1 - t0; a_issue += clock64() - t1; }
            }
            tc::mma_commit_elect(&bar_acc[buf]);
        }
        if (rec) { q.dbg[4] = a_accfree; q.dbg[5] = a_kfull; q.dbg[6] = a_issue; q.dbg[7] = clock64() - t_begin; }
    } else {
        // ============================================ epilogue ============================================
        const int eq = warp & 3, eg = warp >> 2;
        const int et = eq * 32 + lane;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(eq * 32) << 16);
        float *wsum = csum + warp * rows;
        const bool rec = kRec && q.dbg && blockIdx.x == 0 && blockIdx.y == 0 && warp == 0 && lane == 0;
        long long a_acc = 0, a_efull = 0, a_work = 0, t_begin = clock64(), t0 = 0, t1 = 0;
        long long a_ld = 0, a_stg = 0, a_tr = 0, a_fence = 0, t2 = 0;
        int eseq = 0;
        for (int u = 0; u < n_my; ++u) {
            const int t = tile_of(u), buf = u & 1;
            const size_t pos = static_cast<size_t>(t) * kTile + et;
            if (SCATTER) {
                float *sc = scr + warp * (32 * 33);
                const int nchunk_mine = rows > eg * 32 ? (rows - eg * 32 + 63) / 64 : 0;
                const int jpt = __ldg(q.idx + static_cast<size_t>(b) * P + pos);
                if (rec) t0 = clock64();
                mbar_wait(&bar_acc[buf], (u >> 1) & 1);
                tc::fence_after_sync();
                if (rec) { t1 = clock64(); a_acc += t1 - t0; }
                for (int n = 0; n < nchunk_mine; ++n) {
                    const int c0 = eg * 32 + n * 64;
                    float v[32];
                    tc::tmem_ld32(trow + q.col_acc[buf] + c0, v);
                    if (n == nchunk_mine - 1) {
                        tc::fence_before_sync();
                        mbar_arrive(&bar_accfree[buf]);
                    }
                    // transpose the 32 positions x 32 channels block: lane = channel, one coalesced red per position
#pragma unroll
                    for (int j = 0; j < 32; ++j) sc[j * 33 + lane] = v[j];
                    __syncwarp();
                    float *dst = q.dfeat_pm + static_cast<size_t>(b) * q.N * q.dfeat_stride + q.dfeat_off + c0 + lane;
                    const bool act = c0 + lane < rows;
#pragma unroll 8
                    for (int p = 0; p < 32; ++p) {
                        const int jp = __shfl_sync(OGC_FULL_MASK, jpt, p);
                        if (act) atomicAdd(dst + static_cast<size_t>(jp) * q.dfeat_stride, sc[lane * 33 + p]);
                    }
                    __syncwarp();
                }
                if (nchunk_mine == 0) {          // a column group without a chunk still takes part in the hand-shake
                    tc::fence_before_sync();
                    mbar_arrive(&bar_accfree[buf]);
                }
                if (rec) a_work += clock64() - t1;
                continue;
            }
            // dense: per 32-channel chunk of y_{l-1} (one ring stage) this column group owns 16 channels
            const int nech = rows >> 5;
            for (int i = 0; i < nech; ++i, ++eseq) {
                const int stage = eseq % ES;
                const int c0 = 32 * i + 16 * eg;
                if (rec) t0 = clock64();
                mbar_wait(&bar_efull[stage], (eseq / ES) & 1);
                if (rec) { t1 = clock64(); a_efull += t1 - t0; }
                // this warp's [16 channels][32 positions] block of the stage
                const float *blk = reinterpret_cast<const float *>(ering + static_cast<size_t>(stage) * kEStageBytes) + (16 * eg) * kTile + eq * 32;
                float yp[16], v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) yp[j] = blk[j * kTile + lane];
                if (i == 0) {
                    if (rec) t0 = clock64();
                    mbar_wait(&bar_acc[buf], (u >> 1) & 1);
                    tc::fence_after_sync();
                    if (rec) { t1 = clock64(); a_acc += t1 - t0; }
                }
                if (rec) t2 = clock64();
                tc::tmem_ld16(trow + q.col_acc[buf] + c0, v);
                if (i == nech - 1) {
                    tc::fence_before_sync();
                    mbar_arrive(&bar_accfree[buf]);
                }
                if (rec) { const long long t = clock64(); a_ld += t - t2; t2 = t; }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 s4 = tab_ss[c0 + j];
                    const float gz = fmaf(s4.x, yp[j], s4.y) > 0.f ? v[j] : 0.f;
                    v[j] = gz;
                    yp[j] = gz * ((yp[j] - s4.z) * s4.w);      // dz_prev * xhat_prev
                }
                mbar_arrive(&bar_efree[stage]);            // the loaded values have been consumed: the stage may be refilled
                {
                    float *dzp = q.dz_prev + (static_cast<size_t>(b) * q.prev_total + q.prev_off + c0) * P + pos;
                    float *yo[4] = {dzp, dzp + P, dzp + 2 * static_cast<size_t>(P), dzp + 3 * static_cast<size_t>(P)};
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        *yo[j & 3] = v[j];
                        if (j < 12) yo[j & 3] += static_cast<size_t>(4) * P;
                    }
                }
                if (rec) { const long long t = clock64(); a_stg += t - t2; t2 = t; }
                // per-channel sums over this warp's 32 positions: 31-shuffle transpose-reduction of the 16 + 16 values;
                // lane L < 16 ends with sum dz_prev of channel c0 + L, lane 16 + L with sum dz_prev * xhat of the same channel
                {
                    float w[32];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { w[j] = v[j]; w[16 + j] = yp[j]; }
                    const float tot = warp_transpose_sum32(w, lane);
                    wsum[(i * 16 + (lane & 15)) * 2 + (lane >> 4)] += tot;      // this warp's private accumulators
                }
                if (rec) { const long long t = clock64(); a_tr += t - t2; t2 = t; }
                if (rec) { const long long t = clock64(); a_fence += t - t2; a_work += t - t1; }
            }
        }
        if (rec) { q.dbg[8] = a_acc; q.dbg[9] = a_work; q.dbg[10] = clock64() - t_begin; q.dbg[11] = n_my; q.dbg[12] = a_efull;
                   q.dbg[13] = a_ld; q.dbg[14] = a_stg; q.dbg[15] = a_tr; q.dbg[16] = a_fence; }
        if (!SCATTER && n_my > 0) {
            __syncwarp();
            for (int e = lane; e < rows; e += 32) {        // e = (piece i, channel within the piece, which sum)
                const int ch = 32 * (e >> 5) + 16 * eg + ((e >> 1) & 15);
                atomicAdd(q.chan_sums + (static_cast<size_t>(b) * q.prev_total + q.prev_off + ch) * 2 + (e & 1), wsum[e]);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, 512);
}

// chan_sums (B,rows,2) -> dgamma / dbeta (+=), ab (B,4,2) (+=): the inputs of ogc_gn_bwd_coef for layer l-1
__global__ void dx_finalize_kernel(int B, int rows, int prev_total, int prev_off, const float *__restrict__ chan_sums,
                                   const float *__restrict__ gamma, double *__restrict__ ab, float *__restrict__ dgamma,
                                   float *__restrict__ dbeta) {
    const int c = prev_off + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= prev_off + rows) return;
    const int g = c / (prev_total / kGnGroups);
    const double gm = static_cast<double>(gamma[c]);
    float sb = 0.f, sg = 0.f;
    for (int b = 0; b < B; ++b) {
        const float s0 = chan_sums[(static_cast<size_t>(b) * prev_total + c) * 2], s1 = chan_sums[(static_cast<size_t>(b) * prev_total + c) * 2 + 1];
        sb += s0; sg += s1;
        atomicAdd(ab + (b * kGnGroups + g) * 2, gm * s0);
        atomicAdd(ab + (b * kGnGroups + g) * 2 + 1, gm * s1);
    }
    atomicAdd(dbeta + c, sb);
    atomicAdd(dgamma + c, sg);
}

}  // namespace chain
}  // namespace ogc

static long long *g_dx_dbg = nullptr;
static int g_dx_dbg_count = 0;
// Diagnostics: a device buffer of 8 x 32 int64 that the NEXT launches fill with per-role cycle sums of CTA (0,0)
// (producer: load issue / slot wait / rebuild+store / total; issuer: accumulator wait / operand wait / issue / total;
// epilogue: accumulator wait / work / total / tiles); launch k after this call writes at buf + 32 k (k < 8).  NULL = off.
extern "C" int ogc_sa_chain_dx_debug(long long *buf) {
    g_dx_dbg = buf;
    g_dx_dbg_count = 0;
    return OGC_OK;
}

// Drop-in replacement of ogc_sa_mlp_layer_dx_tc (same arguments, same outputs) in the positions-on-M orientation, plus
// `chan_sums`: a caller-zeroed (b, prev_total, 2) fp32 workspace (dense mode), and `prev_total` / `prev_off`: the launch
// covers channels [prev_off, prev_off + rows) of a layer of prev_total channels (0 / 0 = the whole layer; a 256 -> 128
// layer whose resident weights exceed one SM runs as two 64-row launches).  nsample == 64, m even, cout a multiple of 32
// (<= 256), rows a multiple of 16 (<= 128; dense: a multiple of 32).  OGC_ERR_UNSUPPORTED otherwise.
extern "C" int ogc_sa_chain_dx(int b, int n, int m, int nsample, int cout, int cin_full, int row_off, int rows,
                               const float *dz, const float *go, int go_ctotal, int go_coff, const unsigned char *sel,
                               const float *y, const float *coef, const float *w, const float *y_prev,
                               const float *ss_prev, const float *mean_rstd_prev, const float *gamma_prev,
                               float *dz_prev, double *ab_prev, float *dgamma_prev, float *dbeta_prev, const int *idx,
                               float *dfeat_pm, int dfeat_stride, int dfeat_off, float *chan_sums, int prev_total, int prev_off,
                               void *stream) {
    using namespace ogc;
    using namespace ogc::chain;
    if (b < 0 || m <= 0 || cout <= 0 || rows <= 0 || row_off < 0 || row_off + rows > cin_full || !w) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    const bool scatter = dfeat_pm != nullptr;
    if (nsample != kNS || (m & 1) || b > 65535 || cout % 32 != 0 || cout > kMaxC || rows > 128 || rows % 16 != 0)
        return OGC_ERR_UNSUPPORTED;
    DxParams q{};
    int rc = fill_dy(q.dy, cout, m, nsample, dz, go, go_ctotal, go_coff, sel, y, coef);
    if (rc != OGC_OK) return rc;
    if (scatter) {
        if (!idx) return OGC_ERR_INVALID_ARG;
    } else {
        if (!y_prev || !ss_prev || !mean_rstd_prev || !gamma_prev || !dz_prev || !ab_prev || !dgamma_prev || !dbeta_prev || !chan_sums)
            return OGC_ERR_INVALID_ARG;
        if (rows % 32 != 0) return OGC_ERR_UNSUPPORTED;
    }
    q.cin_full = cin_full; q.row_off = row_off; q.rows = rows; q.W = w;
    q.y_prev = y_prev; q.ss_prev = ss_prev; q.mean_rstd_prev = mean_rstd_prev; q.dz_prev = dz_prev; q.chan_sums = chan_sums;
    q.idx = idx; q.dfeat_pm = dfeat_pm; q.N = n; q.dfeat_stride = dfeat_stride; q.dfeat_off = dfeat_off;
    if (prev_total <= 0) { prev_total = rows; prev_off = 0; }
    if (!scatter && (prev_off < 0 || prev_off + rows > prev_total || prev_total % kGnGroups != 0)) return OGC_ERR_INVALID_ARG;
    q.prev_total = prev_total; q.prev_off = prev_off;
    // shared memory: W^T tile | tables | scatter transpose scratch | dY-input ring | mask-input ring
    const bool synth = dz == nullptr;
    const uint32_t w_bytes = w_tile_bytes(rows, cout);
    q.off_w = 0;
    q.off_tab = w_bytes;
    const uint32_t tab_bytes = static_cast<uint32_t>(cout) * 16u + static_cast<uint32_t>(rows) * 16u + kEpiWarps * static_cast<uint32_t>(rows) * 4u;
    q.off_scr = (q.off_tab + tab_bytes + 127u) & ~127u;
    q.off_pring = q.off_scr + (scatter ? kEpiWarps * 32u * 33u * 4u : 0u);
    q.off_pring = (q.off_pring + 127u) & ~127u;
    const uint32_t p_stage = (synth ? 1u : 2u) * 16u * kRowBytes, e_stage = 32u * kRowBytes;
    const long long budget = static_cast<long long>(kMaxSmemPerCta) - 1024 - 1024 - q.off_pring;   // alignment slack, static barriers
    int es = scatter ? 0 : 2;
    long long left = budget - static_cast<long long>(es) * e_stage;
    int ps = static_cast<int>(left / p_stage);
    if (ps > kMaxPStages) ps = kMaxPStages;
    if (ps < 3) return OGC_ERR_UNSUPPORTED;
    if (!scatter && left - static_cast<long long>(ps) * p_stage >= e_stage) es = 3;
    q.p_stages = ps; q.e_stages = es;
    q.off_ering = q.off_pring + static_cast<uint32_t>(ps) * p_stage;
    const size_t smem = static_cast<size_t>(q.off_ering) + static_cast<size_t>(es) * e_stage + 1024;
    // tensor memory: A ring (64 columns per 32-channel chunk) + two accumulators of `rows` columns
    const int acc = align_up(rows, 32);
    int ring = (512 - 2 * acc) / 64;
    if (ring > 8) ring = 8;
    if (ring < 2) return OGC_ERR_UNSUPPORTED;
    q.ring = ring;
    q.dbg = (g_dx_dbg && g_dx_dbg_count < 8) ? g_dx_dbg + 32 * g_dx_dbg_count : nullptr;
    ++g_dx_dbg_count;
    q.col_acc[0] = static_cast<uint32_t>(ring * 64);
    q.col_acc[1] = q.col_acc[0] + static_cast<uint32_t>(acc);
    int per_sample = kNumSMs / b;
    per_sample = per_sample > m / 2 ? m / 2 : (per_sample < 1 ? 1 : per_sample);
    dim3 grid(per_sample, b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUtensorMap tm_y, tm_dz, tm_yp;
    memset(&tm_dz, 0, sizeof(tm_dz));
    memset(&tm_yp, 0, sizeof(tm_yp));
    const uint64_t p64 = static_cast<uint64_t>(m) * nsample;
    bool ok = tma::make_2d_f32(&tm_y, y, p64, static_cast<uint64_t>(b) * cout, kTile, 16);
    if (ok && !synth) ok = tma::make_2d_f32(&tm_dz, dz, p64, static_cast<uint64_t>(b) * cout, kTile, 16);
    if (ok && !scatter) ok = tma::make_2d_f32(&tm_yp, y_prev, p64, static_cast<uint64_t>(b) * prev_total, kTile, 32);
    if (!ok) return OGC_ERR_UNSUPPORTED;
#define OGC_DX_LAUNCH(S, SC)                                                                                          \
    do {                                                                                                              \
        cudaError_t e = cudaFuncSetAttribute(sa_dx_kernel<S, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                             static_cast<int>(smem));                                                 \
        if (e != cudaSuccess) return static_cast<int>(e);                                                             \
        sa_dx_kernel<S, SC><<<grid, kDxThreads, smem, st>>>(q, tm_y, tm_dz, tm_yp);                                                         \
    } while (0)
    if (synth && scatter) OGC_DX_LAUNCH(true, true);
    else if (synth) OGC_DX_LAUNCH(true, false);
    else if (scatter) OGC_DX_LAUNCH(false, true);
    else OGC_DX_LAUNCH(false, false);
#undef OGC_DX_LAUNCH
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    if (!scatter) {
        dx_finalize_kernel<<<(rows + 127) / 128, 128, 0, st>>>(b, rows, prev_total, prev_off, chan_sums, gamma_prev, ab_prev, dgamma_prev,
                                                               dbeta_prev);
    }
    OGC_RETURN_LAUNCH_STATUS();
}
