// Chained set-abstraction forward on the 5th-generation tensor cores: up to three SharedMLP layers
// (conv1x1 -> GroupNorm(4) -> ReLU) of one grouper evaluated for a tile of 128 grouped positions WITHOUT any
// activation leaving the SM (utils/pointnet2_util.py:33-44, utils/nn_util.py:151-168, pointnet2/pointnet2.py:283-294).
//
// GroupNorm needs the statistics of a whole sample before the next layer can be normalised, so a grouper of L layers
// runs L passes of this kernel: pass p recomputes layers 1..p from the gathered input (nothing is stored), and
// accumulates the group sums of layer p; the last pass also reduces max / min (+ positions) over the 64 neighbours
// of every centre.  Where the weights of all layers do not fit one SM's shared memory (SA level 3: 2 x 4 B x 82 K
// weights), the same kernel runs one layer per launch with the pre-norm output stored once (y_out) and read back
// through the dense producer.
//
// Per persistent CTA (one per SM), tile = 2 centres x 64 neighbours:
//   producer warps 8-15 (two per lane quadrant, alternating K chunks): cp.async the tile's raw input into a shared-memory stage one tile ahead (gathered
//        feature rows, or rows of the stored previous layer), then thread = position turns its row into the layer-1
//        A operand in TENSOR MEMORY (GroupNorm+ReLU of the stored layer, hi/lo split)
//   MMA warp 16: one thread issues the layer's k-steps (A from tensor memory, B = resident weights), commits to mbarriers
//   epilogue warps 0-7 (two per lane quadrant, alternating 32-column chunks): thread = position reads its accumulator row, adds the three relative-coordinate input
//        channels (kept on the CUDA cores: exact fp32, no extra k-step), then either writes GroupNorm+ReLU of it as the
//        next layer's A operand back into tensor memory, or -- last computed layer -- accumulates the group sums,
//        stores y, reduces the per-centre extremes (redux.sync over the warp's 32 positions)
#include <cstdlib>

#include "sa_chain.cuh"

namespace ogc {
namespace chain {

struct FwdParams {
    int B, N, M, Cf, K1, nl, gather, small;
    int C[3];               // widths of the computed layers (the last one: this launch's channel slice)
    int c_total;            // full width of the last computed layer
    const float *xyz, *new_xyz, *feat_pm;
    const int *idx;
    const float *y_in, *ss_in;          // dense input: (B,K1,P) pre-norm + (B,K1,2) scale/shift
    const float *W[3];
    int ldw[3];
    const float *ss[2];                 // (B,C_l,2) GroupNorm scale/shift of computed layers 1, 2
    double *sums;                       // (B,4,2) of the last computed layer
    float *y_out;                       // (B,c_total,P) or null
    float *ymax, *ymin;                 // (B,2M,c_total) extremes per half centre, or null
    unsigned char *amax, *amin;
    uint32_t off_w[3], off_raw, off_tab, raw_pitch;
    uint32_t col_a23, col_acc[2];       // tensor-memory columns (A1 at 0)
    int dual;                           // two accumulator regions (see the MMA issuer)
    long long *dbg;                     // optional timeline of CTA (0,0,0): [role 3][tile 64][event 8] SM clocks
};

#define OGC_DBG(role, u, ev)                                                                                     \
    do {                                                                                                          \
        if (q.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (u) < 64)               \
            q.dbg[((role) * 64 + (u)) * 8 + (ev)] = clock64();                                                     \
    } while (0)

__global__ void __launch_bounds__(kThreads, 1)
sa_chain_fwd_kernel(FwdParams q) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_a[3], bar_acc[3], bar_accfree[2];
    __shared__ __align__(8) uint64_t bar_kfull[8], bar_kfree[8];      // layer-1 operand, per 32-column K chunk
    __shared__ uint32_t tmem_base_s;
    __shared__ double gs[kGnGroups * 2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, slice = blockIdx.z;
    const int P = q.M * kNS, ntiles = q.M / 2;
    const int n_my = ntiles > static_cast<int>(blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nl = q.nl, K1 = q.K1;
    const int c_last = q.C[nl - 1], c_off = slice * c_last;
    const bool gather = q.gather != 0;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float2 *tab_in = reinterpret_cast<float2 *>(smem + q.off_tab);            // [K1] dense input scale/shift
    float2 *tab_ss0 = tab_in + kMaxC, *tab_ss1 = tab_ss0 + kMaxC;             // layers 1, 2
    float4 *tab_wx = reinterpret_cast<float4 *>(tab_ss1 + kMaxC);            // xyz columns of W1, per channel quad: wx[4] wy[4] wz[4]
    float *pool_s = reinterpret_cast<float *>(tab_wx + kMaxC);                // [8 epilogue warps][8][33] pooling transpose

    // tensor-memory columns: A1 (hi | lo) at 0, A2 / A3 (hi | lo) at col_a23, accumulator regions 0 / 1.
    // Chained layers: layer 1 accumulates in region 1, layers 2.. in region 0, so the next tile's first layer runs
    // underneath the last (heaviest) epilogue of the current tile.  One layer per launch: tiles alternate regions.
    const uint32_t colA1 = 0, colA23 = q.col_a23;
    const bool dual = q.dual != 0, single = nl == 1;
    if (warp == kMmaWarp) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int l = 1; l < 3; ++l) mbar_init(&bar_a[l], kEpi);
        for (int c = 0; c < 8; ++c) { mbar_init(&bar_kfull[c], 128); mbar_init(&bar_kfree[c], 1); }
        for (int l = 0; l < 3; ++l) mbar_init(&bar_acc[l], 1);
        mbar_init(&bar_accfree[0], kEpi);
        mbar_init(&bar_accfree[1], kEpi);
        mbar_fence_init();
    }
    if (tid < kGnGroups * 2) gs[tid] = 0.0;
    // ---- resident B operands ----
    for (int l = 0; l < nl; ++l) {
        const bool lastl = l == nl - 1;
        const float *Wl = q.W[l] + (lastl ? static_cast<size_t>(c_off) * q.ldw[l] : 0);
        if (l == 0 && gather) {
            const int Cf = q.Cf;
            build_weights(smem + q.off_w[l], Wl, q.ldw[l], q.C[l], Cf, [](int k) { return 3 + k; }, tid, kThreads);
        } else {
            build_weights(smem + q.off_w[l], Wl, q.ldw[l], q.C[l], l == 0 ? K1 : q.C[l - 1], [](int k) { return k; }, tid, kThreads);
        }
    }
    if (gather) {
        const float *W0 = q.W[0] + (nl == 1 ? static_cast<size_t>(c_off) * q.ldw[0] : 0);
        for (int e = tid; e < q.C[0] * 3; e += kThreads) {
            const int c = e / 3, x = e - c * 3;
            reinterpret_cast<float *>(tab_wx)[(c >> 2) * 12 + x * 4 + (c & 3)] = __ldg(W0 + static_cast<size_t>(c) * q.ldw[0] + x);
        }
    } else {
        for (int c = tid; c < K1; c += kThreads)
            tab_in[c] = __ldg(reinterpret_cast<const float2 *>(q.ss_in) + static_cast<size_t>(b) * K1 + c);
    }
    for (int l = 0; l + 1 < nl; ++l) {
        float2 *t = l == 0 ? tab_ss0 : tab_ss1;
        for (int c = tid; c < q.C[l]; c += kThreads)
            t[c] = __ldg(reinterpret_cast<const float2 *>(q.ss[l]) + static_cast<size_t>(b) * q.C[l] + c);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp >= kProdWarp0 && warp < kMmaWarp) {
        // ============================================ producer ============================================
        const int pw = warp & 3, pg = (warp - kProdWarp0) >> 2;      // lane quadrant; chunk group: chunks pg, pg + 2, ...
        const int pt = pw * 32 + lane;                               // position within the tile
        const uint32_t raw = smem_u32(smem + q.off_raw), pitch = q.raw_pitch;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(pw * 32) << 16) + colA1;
        const int Cf = q.Cf;
        auto tile_p0 = [&](int u) { return (static_cast<int>(blockIdx.x) + u * static_cast<int>(gridDim.x)) * kTile; };
        auto load_j = [&](int u) { return (gather && u < n_my) ? __ldg(q.idx + static_cast<size_t>(b) * P + tile_p0(u) + pt) : 0; };
        // Raw-stage copies, per 32-column K chunk (cp.async, a handful of instructions per row): chunk c of the NEXT
        // tile is requested as soon as every producer thread has consumed chunk c of the current one, so each copy has
        // a whole tile period to land (a single whole-tile stage exposed the full L2 latency once per tile).
        const int nchunks = (K1 + 31) >> 5;
        const int my_chunks = (nchunks - pg + 1) >> 1;               // the two groups are independent pipelines
        const int pbar = kProdBar + pg;
        const float *gsrc = q.feat_pm + static_cast<size_t>(b) * q.N * Cf + (lane & 7) * 4;
        const uint32_t gdst = raw + static_cast<uint32_t>(pw * 32 + (lane >> 3)) * pitch + (lane & 7) * 16;
        const float *dsrc = q.y_in + static_cast<size_t>(b) * K1 * P + static_cast<size_t>(pt >> 5) * P + (pt & 31) * 4;
        const uint32_t ddst = raw + static_cast<uint32_t>(pt >> 5) * kDensePitch + (pt & 31) * 16;
        auto issue_chunk = [&](int u, int c, int j) {
            if (gather) {
                if (q.small) return;
                const bool act = (lane & 7) * 4 < K1 - 32 * c;
#pragma unroll
                for (int i = 0; i < 8; ++i) {                      // 4 rows x 128 B per instruction
                    const int ji = __shfl_sync(OGC_FULL_MASK, j, 4 * i + (lane >> 3));
                    if (act) cp_async16_s(gdst + static_cast<uint32_t>(4 * i) * pitch + c * 128, gsrc + static_cast<size_t>(ji) * Cf + 32 * c);
                }
            } else {
                const float *src = dsrc + tile_p0(u) + static_cast<size_t>(32 * c) * P;
#pragma unroll
                for (int i = 0; i < 8; ++i)                        // rows 32c + (pt>>5) + 4i
                    cp_async16_s(ddst + static_cast<uint32_t>(32 * c + 4 * i) * kDensePitch, src + static_cast<size_t>(4 * i) * P);
            }
        };
        auto load_small = [&](int j, float (&f8)[8]) {
#pragma unroll
            for (int c = 0; c < 8; ++c) f8[c] = c < Cf ? __ldg(q.feat_pm + (static_cast<size_t>(b) * q.N + j) * Cf + c) : 0.f;
        };
        int j_cur = load_j(0);
        float f8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (gather && q.small && n_my > 0) load_small(j_cur, f8);
        for (int c = pg; c < nchunks; c += 2) {
            if (n_my > 0) issue_chunk(0, c, j_cur);
            cp_async_commit();
        }
        for (int u = 0; u < n_my; ++u) {
            const int j_next = load_j(u + 1);
            const uint32_t fpar = (u & 1) ^ 1;
            if (warp == kProdWarp0) OGC_DBG(0, u, 0);
            // The operand is handed over per K chunk as well: chunk c of this tile is written as soon as the previous
            // tile's MMAs have consumed chunk c, and the MMAs of this tile start on chunk 0 while chunk 1 is written.
            if (gather && q.small) {
                if (pg != 0) { j_cur = j_next; continue; }
                mbar_wait(&bar_kfree[0], fpar);
                tc::fence_after_sync();
                float hi[8], lo[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) tc::tf32_split(f8[c], hi[c], lo[c]);
                tc::tmem_st8_nowait(trow, hi);
                tc::tmem_st8_nowait(trow + 8, lo);
                tc::tmem_st_wait();
                tc::fence_before_sync();
                mbar_arrive(&bar_kfull[0]);
                if (u + 1 < n_my) load_small(j_next, f8);      // next tile's rows: in flight during this tile's chain
                j_cur = j_next;
                continue;
            }
            for (int c = pg; c < nchunks; c += 2) {
                const int k0 = 32 * c;
                // groups committed after (u, c): the rest of this tile's chunks and the next tile's chunks requested so far
                cp_async_wait(c == pg ? my_chunks - 1 : my_chunks - 2);   // this thread's copies of (u, c) have landed ...
                named_bar_sync(pbar, 128);                     // ... the group's; and its previous chunk has been consumed
                if (warp == kProdWarp0 && c == 0) OGC_DBG(0, u, 1);
                if (c >= 2) {
                    if (u + 1 < n_my) issue_chunk(u + 1, c - 2, j_next);
                    cp_async_commit();
                }
                mbar_wait(&bar_kfree[c], fpar);
                tc::fence_after_sync();
                if (warp == kProdWarp0 && c == 0) OGC_DBG(0, u, 2);
                if (k0 + 32 <= K1) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {                  // 16 columns at a time: 48 live registers
                        float hi[16], lo[16];
                        if (gather) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float4 x = lds_v4(raw + static_cast<uint32_t>(pt) * pitch + (k0 + 16 * h + 4 * i) * 4);
                                tc::tf32_split(x.x, hi[4 * i], lo[4 * i]);
                                tc::tf32_split(x.y, hi[4 * i + 1], lo[4 * i + 1]);
                                tc::tf32_split(x.z, hi[4 * i + 2], lo[4 * i + 2]);
                                tc::tf32_split(x.w, hi[4 * i + 3], lo[4 * i + 3]);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float2 s2 = tab_in[k0 + 16 * h + i];
                                const float y = lds_f32(raw + static_cast<uint32_t>(k0 + 16 * h + i) * kDensePitch + pt * 4);
                                tc::tf32_split(fmaxf(fmaf(s2.x, y, s2.y), 0.f), hi[i], lo[i]);
                            }
                        }
                        tc::tmem_st16_nowait(trow + k0 + 16 * h, hi);
                        tc::tmem_st16_nowait(trow + K1 + k0 + 16 * h, lo);
                    }
                } else {
                    for (int k1 = k0; k1 < K1; k1 += 8) {          // tail in 8-column pieces (gather only: K1 % 8 == 0)
                        float hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float4 x = lds_v4(raw + static_cast<uint32_t>(pt) * pitch + (k1 + 4 * i) * 4);
                            tc::tf32_split(x.x, hi[4 * i], lo[4 * i]);
                            tc::tf32_split(x.y, hi[4 * i + 1], lo[4 * i + 1]);
                            tc::tf32_split(x.z, hi[4 * i + 2], lo[4 * i + 2]);
                            tc::tf32_split(x.w, hi[4 * i + 3], lo[4 * i + 3]);
                        }
                        tc::tmem_st8_nowait(trow + k1, hi);
                        tc::tmem_st8_nowait(trow + K1 + k1, lo);
                    }
                }
                tc::tmem_st_wait();
                tc::fence_before_sync();
                mbar_arrive(&bar_kfull[c]);
                if (warp == kProdWarp0 && c == 0) OGC_DBG(0, u, 3);
            }
            if (warp == kProdWarp0) OGC_DBG(0, u, 4);
            if (my_chunks > 0) {
                named_bar_sync(pbar, 128);                     // the group's last chunk has been consumed
                if (u + 1 < n_my) issue_chunk(u + 1, pg + 2 * (my_chunks - 1), j_next);
                cp_async_commit();
            }
            j_cur = j_next;
        }
    } else if (warp == kMmaWarp) {
        // ============================================ MMA issuer ============================================
        // warp-uniform: all lanes run the loop, one elected lane issues each MMA / commit (see tcgen05.cuh)
        {
            for (int u = 0; u < n_my; ++u) {
                for (int l = 0; l < nl; ++l) {
                    if (l > 0) mbar_wait(&bar_a[l], u & 1);
                    int buf = 0;
                    if (single) {
                        buf = dual ? (u & 1) : 0;
                        mbar_wait(&bar_accfree[buf], dual ? (((u >> 1) & 1) ^ 1) : ((u & 1) ^ 1));
                    } else {
                        buf = (l == 0 && dual) ? 1 : 0;
                        // region 0 was last read by the previous tile's final epilogue
                        if (l == (dual ? 1 : 0)) mbar_wait(&bar_accfree[0], (u & 1) ^ 1);
                    }
                    const uint32_t d = tmem_base + q.col_acc[buf];
                    if (l == 0) OGC_DBG(1, u, 0);
                    if (l == 0) {
                        // layer 1: chunk by chunk as the producer hands the operand over; each chunk is released to the
                        // producer (for the NEXT tile) as soon as its MMAs have completed
                        const int n = q.C[0];
                        const uint32_t idesc = tc::make_idesc_tf32(kTile, n, 0, 0);
                        const uint32_t blk16 = (2u * static_cast<uint32_t>(n) * 128u) >> 4, lo16 = (static_cast<uint32_t>(n) * 128u) >> 4;
                        const uint64_t d0 = tc::make_desc_sw128(smem_u32(smem + q.off_w[0]), 16, 1024);
                        for (int k0 = 0; k0 < K1; k0 += 32) {
                            mbar_wait(&bar_kfull[k0 >> 5], u & 1);
                            tc::fence_after_sync();
                            if (k0 == 0) OGC_DBG(1, u, 1);
                            const int ks = min(4, (K1 - k0) >> 3);
#pragma unroll
                            for (int s = 0; s < 4; ++s) {
                                if (s < ks) {
                                    const uint64_t bh = d0 + (static_cast<uint32_t>(k0 >> 5) * blk16 + static_cast<uint32_t>(s) * 2u), bl = bh + lo16;
                                    const uint32_t ah = tmem_base + colA1 + static_cast<uint32_t>(k0 + s * 8), al = ah + static_cast<uint32_t>(K1);
                                    tc::mma_tf32_ts_elect(d, ah, bh, idesc, (k0 | s) ? 1u : 0u);
                                    tc::mma_tf32_ts_elect(d, ah, bl, idesc, 1u);
                                    tc::mma_tf32_ts_elect(d, al, bh, idesc, 1u);
                                }
                            }
                            tc::mma_commit_elect(&bar_kfree[k0 >> 5]);
                        }
                    } else {
                        tc::fence_after_sync();
                        issue_layer(d, tmem_base + colA23, q.C[l - 1], smem_u32(smem + q.off_w[l]), q.C[l]);
                    }
                    // one layer per launch with two regions: a "full" barrier per region (the issuer may run two tiles
                    // ahead of the epilogue, which would alias the phase parity of a single barrier)
                    tc::mma_commit_elect(&bar_acc[single ? buf : l]);
                    OGC_DBG(1, u, 2 + l);
                    if (q.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && l == nl - 1) {
                        // diagnostics only: completion time of the tile's MMAs (serialises issue and execution)
                        mbar_wait(&bar_acc[single ? buf : l], (single && dual) ? ((u >> 1) & 1) : (u & 1));
                        OGC_DBG(1, u, 6);
                    }
                }
            }
        }
    } else {
        // ============================================ epilogue ============================================
        const int eq = warp & 3, eg = warp >> 2;             // lane quadrant, column group (chunks eg, eg + 2, ...)
        const int et = eq * 32 + lane;                       // position within the tile
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(eq * 32) << 16);
        // GroupNorm group of every 8-channel block of the last layer, 2 bits each (c_total <= 256: 32 blocks)
        unsigned long long gbits = 0;
        {
            const int gsz8 = q.c_total / (kGnGroups * 8);
            for (int k = 0; k < q.c_total / 8; ++k) gbits |= static_cast<unsigned long long>(k / gsz8) << (2 * k);
        }
        float fsum[kGnGroups] = {0.f, 0.f, 0.f, 0.f}, fsq[kGnGroups] = {0.f, 0.f, 0.f, 0.f};
        float nrx = 0.f, nry = 0.f, nrz = 0.f;
        auto load_rel = [&](int t, int j) {
            const float *pj = q.xyz + (static_cast<size_t>(b) * q.N + j) * 3;
            const float *pc = q.new_xyz + (static_cast<size_t>(b) * q.M + t * 2 + (et >> 6)) * 3;
            nrx = __ldg(pj) - __ldg(pc); nry = __ldg(pj + 1) - __ldg(pc + 1); nrz = __ldg(pj + 2) - __ldg(pc + 2);
        };
        if (gather && n_my > 0)
            load_rel(blockIdx.x, __ldg(q.idx + static_cast<size_t>(b) * P + static_cast<size_t>(blockIdx.x) * kTile + et));
        for (int u = 0; u < n_my; ++u) {
            const int t = static_cast<int>(blockIdx.x) + u * static_cast<int>(gridDim.x);
            const int p0 = t * kTile;
            // relative coordinates of this thread's position (3 input channels kept on the CUDA cores): loaded one tile
            // ahead (index at the top of the previous tile, coordinates after its first layer)
            const float rx = nrx, ry = nry, rz = nrz;
            const int jn = (gather && u + 1 < n_my) ? __ldg(q.idx + static_cast<size_t>(b) * P + p0 + static_cast<int>(gridDim.x) * kTile + et) : 0;
            for (int l = 0; l < nl; ++l) {
                const bool lastl = l == nl - 1;
                const int n = q.C[l];
                const int buf = single ? (dual ? (u & 1) : 0) : ((l == 0 && dual) ? 1 : 0);
                mbar_wait(&bar_acc[single ? buf : l], (single && dual) ? ((u >> 1) & 1) : (u & 1));
                tc::fence_after_sync();
                if (warp == 0) OGC_DBG(2, u, 2 * l);
                if (l == 0 && gather && u + 1 < n_my) load_rel(t + static_cast<int>(gridDim.x), jn);   // next tile's, in flight during this one
                const uint32_t colACC = q.col_acc[buf];
                const float2 *tss = l == 0 ? tab_ss0 : tab_ss1;
                bool released = false;
                for (int c0 = eg * 32; c0 < n; c0 += 64) {
                    float v[32];
                    tc::tmem_ld32(trow + colACC + c0, v);
                    if (lastl && c0 + 64 >= n) {            // this warp has drained its share of the accumulator
                        tc::fence_before_sync();
                        mbar_arrive(&bar_accfree[single ? buf : 0]);
                        released = true;
                    }
                    if (l == 0 && gather) {
                        const float4 *wq = tab_wx + (c0 >> 2) * 3;
                        const float2 rx2 = make_float2(rx, rx), ry2 = make_float2(ry, ry), rz2 = make_float2(rz, rz);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {       // 4 channels: 3 broadcast LDS.128 + 6 packed FMAs
                            const float4 wx = wq[3 * j], wy = wq[3 * j + 1], wz = wq[3 * j + 2];
                            float2 a = make_float2(v[4 * j], v[4 * j + 1]), c = make_float2(v[4 * j + 2], v[4 * j + 3]);
                            a = ffma2(make_float2(wx.x, wx.y), rx2, a); c = ffma2(make_float2(wx.z, wx.w), rx2, c);
                            a = ffma2(make_float2(wy.x, wy.y), ry2, a); c = ffma2(make_float2(wy.z, wy.w), ry2, c);
                            a = ffma2(make_float2(wz.x, wz.y), rz2, a); c = ffma2(make_float2(wz.z, wz.w), rz2, c);
                            v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = c.x; v[4 * j + 3] = c.y;
                        }
                    }
                    if (!lastl) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float hi[16], lo[16];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 s4 = *reinterpret_cast<const float4 *>(tss + c0 + 16 * h + 2 * j);
                                tc::tf32_split(fmaxf(fmaf(s4.x, v[16 * h + 2 * j], s4.y), 0.f), hi[2 * j], lo[2 * j]);
                                tc::tf32_split(fmaxf(fmaf(s4.z, v[16 * h + 2 * j + 1], s4.w), 0.f), hi[2 * j + 1], lo[2 * j + 1]);
                            }
                            tc::tmem_st16_nowait(trow + colA23 + c0 + 16 * h, hi);
                            tc::tmem_st16_nowait(trow + colA23 + n + c0 + 16 * h, lo);
                        }
                        continue;
                    }
                    // ---- last computed layer: statistics, optional store, optional pooling ----
                    const int blk0 = (c_off + c0) >> 3;                  // first 8-channel block of the chunk
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 one2 = make_float2(1.f, 1.f);
                        const float2 p0 = make_float2(v[8 * i], v[8 * i + 1]), p1 = make_float2(v[8 * i + 2], v[8 * i + 3]);
                        const float2 p2 = make_float2(v[8 * i + 4], v[8 * i + 5]), p3 = make_float2(v[8 * i + 6], v[8 * i + 7]);
                        const float2 sa = ffma2(p1, one2, p0), sb = ffma2(p3, one2, p2);       // packed fp32 adds
                        const float2 s2 = ffma2(sa, one2, sb);
                        const float2 qa = ffma2(p1, p1, ffma2(p0, p0, make_float2(0.f, 0.f)));
                        const float2 qb = ffma2(p3, p3, ffma2(p2, p2, qa));
                        const float s = s2.x + s2.y, sq = qb.x + qb.y;
                        const int g = static_cast<int>(gbits >> (2 * (blk0 + i))) & 3;   // warp-uniform
#pragma unroll
                        for (int gg = 0; gg < kGnGroups; ++gg) {
                            fsum[gg] += g == gg ? s : 0.f;
                            fsq[gg] += g == gg ? sq : 0.f;
                        }
                    }
                    if (q.y_out) {
                        // 8 independent address registers: a single running pointer serialises the 32 stores on its
                        // read-after-store scoreboard (measured: 21 % of all stall samples on this line)
                        float *yo[4];
                        yo[0] = q.y_out + (static_cast<size_t>(b) * q.c_total + c_off + c0) * P + p0 + et;
#pragma unroll
                        for (int j = 1; j < 4; ++j) yo[j] = yo[j - 1] + P;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            *yo[j & 3] = v[j];
                            if (j < 28) yo[j & 3] += static_cast<size_t>(4) * P;
                        }
                    }
                    if (q.ymax) {
                        // max / min (+ first position) over this warp's 32 positions, 8 channels at a time through a
                        // per-warp [8][33] transpose: lane = (channel, position quarter) scans 8 positions, two shuffle
                        // rounds join the quarters.  (redux.sync per channel serialised on its uniform result register:
                        // ~190 cycles per channel.)
                        float *sc = pool_s + warp * (8 * 33);
                        const int pch = lane >> 2, pq = lane & 3;
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) sc[j * 33 + lane] = v[8 * g8 + j];
                            __syncwarp();
                            float vmx = sc[pch * 33 + pq * 8], vmn = vmx;
                            int imx = pq * 8, imn = imx;
#pragma unroll
                            for (int k = 1; k < 8; ++k) {
                                const float x = sc[pch * 33 + pq * 8 + k];
                                if (x > vmx) { vmx = x; imx = pq * 8 + k; }
                                if (x < vmn) { vmn = x; imn = pq * 8 + k; }
                            }
#pragma unroll
                            for (int o = 1; o <= 2; o <<= 1) {
                                const float ox = __shfl_xor_sync(OGC_FULL_MASK, vmx, o), on = __shfl_xor_sync(OGC_FULL_MASK, vmn, o);
                                const int oix = __shfl_xor_sync(OGC_FULL_MASK, imx, o), oin = __shfl_xor_sync(OGC_FULL_MASK, imn, o);
                                if (ox > vmx || (ox == vmx && oix < imx)) { vmx = ox; imx = oix; }
                                if (on < vmn || (on == vmn && oin < imn)) { vmn = on; imn = oin; }
                            }
                            if (pq == 0) {
                                const size_t o = (static_cast<size_t>(b) * 2 * q.M + static_cast<size_t>(t) * 4 + eq) * q.c_total + c_off + c0 + 8 * g8 + pch;
                                q.ymax[o] = vmx;
                                q.ymin[o] = vmn;
                                q.amax[o] = static_cast<unsigned char>(imx);
                                q.amin[o] = static_cast<unsigned char>(imn);
                            }
                            __syncwarp();
                        }
                    }
                }
                if (warp == 0) OGC_DBG(2, u, 2 * l + 1);
                if (!lastl) {
                    tc::tmem_st_wait();
                    tc::fence_before_sync();
                    mbar_arrive(&bar_a[l + 1]);
                } else if (!released) {                      // a column group without a chunk of this layer
                    tc::fence_before_sync();
                    mbar_arrive(&bar_accfree[single ? buf : 0]);
                }
            }
            if ((u & 3) == 3 || u + 1 == n_my) {
                // fp32 partial sums (<= 4 tiles) -> warp sum -> fp64 in shared memory (fp64 adds are slow on this part,
                // and 16 fp64 accumulator registers per thread do not fit next to the chunk at 17 warps)
#pragma unroll
                for (int g = 0; g < kGnGroups; ++g) {
                    float s = fsum[g], sq = fsq[g];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        s += __shfl_xor_sync(OGC_FULL_MASK, s, o);
                        sq += __shfl_xor_sync(OGC_FULL_MASK, sq, o);
                    }
                    if (lane == 0) { atomicAdd(&gs[2 * g], static_cast<double>(s)); atomicAdd(&gs[2 * g + 1], static_cast<double>(sq)); }
                    fsum[g] = fsq[g] = 0.f;
                }
            }
        }
        named_bar_sync(kEpiBar, kEpi);
        if (tid < kGnGroups * 2 && n_my > 0) atomicAdd(q.sums + static_cast<size_t>(b) * kGnGroups * 2 + tid, gs[tid]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, 512);
}

// out = relu(max_s(scale*y+shift)) from the per-half-centre extremes written by the chained forward:
// in (B,2M,C) point-major -> out (B,c_total,M) channel-major (+ optional point-major twin), sel / ysel (B,C,M).
// Same contract as sa_finish_kernel (mlp.cu); a 32 x 32 (centre x channel) tile is transposed through shared memory.
__global__ void __launch_bounds__(256)
sa_pool_finish_kernel(int C, int M, const float *__restrict__ ymax, const float *__restrict__ ymin,
                      const unsigned char *__restrict__ amax, const unsigned char *__restrict__ amin,
                      const float *__restrict__ ss, float *__restrict__ out, float *__restrict__ out_pm, int c_total,
                      int c_offset, unsigned char *__restrict__ sel, float *__restrict__ ysel) {
    __shared__ float so[32][33], sy[32][33];
    __shared__ unsigned char ssel[32][36];
    const int b = blockIdx.z, m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = c0 + tx;
    float sc = 0.f, sh = 0.f;
    if (c < C) { sc = ss[(static_cast<size_t>(b) * C + c) * 2]; sh = ss[(static_cast<size_t>(b) * C + c) * 2 + 1]; }
    const bool up = sc >= 0.f;
    for (int r = ty; r < 32; r += 8) {
        const int m = m0 + r;
        if (m < M && c < C) {
            const size_t i0 = (static_cast<size_t>(b) * 2 * M + 2 * m) * C + c, i1 = i0 + C;
            const float v0 = up ? ymax[i0] : ymin[i0], v1 = up ? ymax[i1] : ymin[i1];
            const int a0 = up ? amax[i0] : amin[i0], a1 = up ? amax[i1] : amin[i1];
            const bool first = up ? (v0 >= v1) : (v0 <= v1);          // ties: the earlier position wins
            const float yv = first ? v0 : v1;
            const int pos = first ? a0 : 32 + a1;
            const float z = fmaf(sc, yv, sh);
            const float o = fmaxf(z, 0.f);
            so[r][tx] = o;
            sy[r][tx] = yv;
            ssel[r][tx] = z > 0.f ? static_cast<unsigned char>(pos) : static_cast<unsigned char>(255);
            if (out_pm) out_pm[(static_cast<size_t>(b) * M + m) * c_total + c_offset + c] = o;
        }
    }
    __syncthreads();
    const int m = m0 + tx;
    for (int r = ty; r < 32; r += 8) {
        const int cc = c0 + r;
        if (m < M && cc < C) {
            out[(static_cast<size_t>(b) * c_total + c_offset + cc) * M + m] = so[tx][r];
            const size_t i = (static_cast<size_t>(b) * C + cc) * M + m;
            sel[i] = ssel[tx][r];
            ysel[i] = sy[tx][r];
        }
    }
}

}  // namespace chain
}  // namespace ogc

namespace ogc {
namespace chain {

// Shared / tensor memory plan of a launch: channel slices of the last layer (1, 2 or 4) until the resident weights
// fit next to the raw stage; false when the shape cannot run.  Fills the offsets, C[nl-1] (slice width) and c_total.
static bool fwd_plan(FwdParams &q, const int *widths, int &nslice, size_t &smem) {
    const int nl = q.nl;
    const bool gather = q.gather != 0;
    q.c_total = widths[nl - 1];
    q.raw_pitch = gather ? static_cast<uint32_t>(q.Cf) * 4u + 16u : kTile * 4u + 16u;
    const uint32_t raw_bytes = q.small ? 0u : (gather ? kTile * q.raw_pitch : static_cast<uint32_t>(q.K1) * q.raw_pitch);
    const uint32_t tab_bytes = 3u * kMaxC * 8u + kMaxC * 16u + kEpiWarps * 8u * 33u * 4u;
    const size_t budget = static_cast<size_t>(kMaxSmemPerCta) - 2048;
    for (nslice = 1;; nslice *= 2) {
        if (nslice > 4 || (widths[nl - 1] / nslice) % 32 != 0) return false;
        size_t wsum = 0;
        for (int l = 0; l < nl; ++l)
            wsum += w_tile_bytes(l == nl - 1 ? widths[l] / nslice : widths[l], l == 0 ? q.K1 : widths[l - 1]);
        if (wsum + raw_bytes + tab_bytes + 1024 <= budget) break;
    }
    const int sw = widths[nl - 1] / nslice, gw = widths[nl - 1] / kGnGroups;
    if (sw % gw != 0 && gw % sw != 0) return false;          // a 8-channel sub-chunk never straddles two groups anyway
    for (int l = 0; l < nl; ++l) q.C[l] = widths[l];
    q.C[nl - 1] = sw;
    uint32_t off = 0;
    for (int l = 0; l < nl; ++l) {
        q.off_w[l] = off;
        off += w_tile_bytes(q.C[l], l == 0 ? q.K1 : q.C[l - 1]);
    }
    q.off_raw = off;
    off += (raw_bytes + 127u) & ~127u;
    q.off_tab = off;
    off += tab_bytes;
    smem = static_cast<size_t>(off) + 1024;
    // tensor-memory columns: A1 (hi | lo), A2 / A3 (hi | lo), accumulator region 0 (+ region 1 when it fits)
    int cols = align_up(2 * q.K1, 32);
    q.col_a23 = static_cast<uint32_t>(cols);
    if (nl > 1) cols += 2 * (nl > 2 ? (widths[0] > widths[1] ? widths[0] : widths[1]) : widths[0]);
    int wa = 0;                                   // region 0: layers 2.. (chained) or the only layer
    for (int l = nl > 1 ? 1 : 0; l < nl; ++l) wa = q.C[l] > wa ? q.C[l] : wa;
    const int wb = q.C[0];                        // region 1: layer 1 (chained) / odd tiles (one layer per launch)
    q.dual = cols + wa + wb <= 512;
    q.col_acc[0] = static_cast<uint32_t>(cols);
    q.col_acc[1] = static_cast<uint32_t>(q.dual ? cols + wa : cols);
    return cols + (q.dual ? wa + wb : (wa > wb ? wa : wb)) <= 512;
}

static int fwd_shape(FwdParams &q, int m, int nsample, int cf, int gather, int nl, const int *widths) {
    if (nsample != kNS || (m & 1)) return OGC_ERR_UNSUPPORTED;
    q.Cf = cf; q.nl = nl; q.gather = gather != 0;
    q.small = gather && cf <= 8;
    if (gather) {
        if (!q.small && cf % 8 != 0) return OGC_ERR_UNSUPPORTED;
        q.K1 = q.small ? 8 : cf;
    } else {
        if (cf % 32 != 0) return OGC_ERR_UNSUPPORTED;
        q.K1 = cf;
    }
    if (q.K1 > kMaxC) return OGC_ERR_UNSUPPORTED;
    for (int l = 0; l < nl; ++l)
        if (widths[l] % 32 != 0 || widths[l] < 32 || widths[l] > kMaxC) return OGC_ERR_UNSUPPORTED;
    return OGC_OK;
}

}  // namespace chain
}  // namespace ogc

// 1 when ogc_sa_chain_fwd can run `nl` chained layers of these widths in one launch (weights of all of them resident
// in one SM's shared memory, operands + accumulator within the 512 tensor-memory columns), else 0.
// Diagnostics: when set, the next launches record a per-tile timeline of CTA (0,0,0) (3 roles x 64 tiles x 8 events).
static long long *g_chain_dbg = nullptr;
static int g_chain_dbg_count = 0, g_chain_dbg_sel = -1;    // OGC_CHAIN_DBG_SEL = k: only the k-th launch records
extern "C" int ogc_sa_chain_debug(long long *buf) {
    g_chain_dbg = buf;
    g_chain_dbg_count = 0;
    const char *e = getenv("OGC_CHAIN_DBG_SEL");
    g_chain_dbg_sel = e ? atoi(e) : -1;
    return OGC_OK;
}

extern "C" int ogc_sa_chain_fits(int m, int nsample, int cf, int gather, int nl, const int *widths) {
    using namespace ogc::chain;
    if (nl < 1 || nl > 3 || !widths || cf <= 0) return 0;
    FwdParams q{};
    if (fwd_shape(q, m, nsample, cf, gather, nl, widths) != OGC_OK) return 0;
    int nslice = 0;
    size_t smem = 0;
    return fwd_plan(q, widths, nslice, smem) ? 1 : 0;
}

// Chained forward of `nl` (1..3) SharedMLP layers of one set-abstraction grouper; see the header of this file.
//   gather != 0: input = [xyz[idx] - new_xyz (3) | feat_pm[idx] (cf)], w1 is (c1, 3 + cf) row-major;
//   gather == 0: input = relu(ss_in * y_in + ss_in') with y_in (b, cf, m*64) channel-major, w1 is (c1, cf).
// widths[l] = output channels of computed layer l (multiples of 32); ss1 / ss2 = GroupNorm (scale, shift) (b,c,2) of
// computed layers 1 / 2 (needed when nl > 1 / nl > 2).  Outputs, all for the LAST computed layer: sums (b,4,2)
// (accumulated: zero it first), y_out (b,c,m*64) or NULL, and -- when ymax_h is given -- the extremes over each half
// centre (b, 2m, c) for ogc_sa_pool_finish.  nsample must be 64, m even.  Returns OGC_ERR_UNSUPPORTED for shapes that
// do not fit tensor / shared memory (callers fall back to the per-layer kernels).
extern "C" int ogc_sa_chain_fwd(int b, int n, int m, int nsample, int cf, int gather, int nl, const int *widths,
                                const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                                const float *y_in, const float *ss_in, const float *w1, const float *w2,
                                const float *w3, const float *ss1, const float *ss2, double *sums, float *y_out,
                                float *ymax_h, float *ymin_h, unsigned char *amax_h, unsigned char *amin_h,
                                void *stream) {
    using namespace ogc;
    using namespace ogc::chain;
    if (b < 0 || m <= 0 || cf <= 0 || nl < 1 || nl > 3 || !widths || !w1 || !sums) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    FwdParams q{};
    int rc = fwd_shape(q, m, nsample, cf, gather, nl, widths);
    if (rc != OGC_OK) return rc;
    q.B = b; q.N = n; q.M = m;
    if (gather && (!xyz || !new_xyz || !feat_pm || !idx)) return OGC_ERR_INVALID_ARG;
    if (!gather && (!y_in || !ss_in)) return OGC_ERR_INVALID_ARG;
    const float *W[3] = {w1, w2, w3};
    const float *SS[2] = {ss1, ss2};
    for (int l = 0; l < nl; ++l) {
        if (!W[l] || (l + 1 < nl && !SS[l])) return OGC_ERR_INVALID_ARG;
        q.W[l] = W[l];
        q.ldw[l] = l == 0 ? (gather ? cf + 3 : cf) : widths[l - 1];
    }
    if (ymax_h && (!ymin_h || !amax_h || !amin_h)) return OGC_ERR_INVALID_ARG;
    q.xyz = xyz; q.new_xyz = new_xyz; q.feat_pm = feat_pm; q.idx = idx; q.y_in = y_in; q.ss_in = ss_in;
    q.ss[0] = ss1; q.ss[1] = ss2; q.sums = sums; q.y_out = y_out;
    q.ymax = ymax_h; q.ymin = ymin_h; q.amax = amax_h; q.amin = amin_h;
    q.dbg = (g_chain_dbg_sel < 0 || g_chain_dbg_count == g_chain_dbg_sel) ? g_chain_dbg : nullptr;
    ++g_chain_dbg_count;
    int nslice = 0;
    size_t smem = 0;
    if (!fwd_plan(q, widths, nslice, smem)) return OGC_ERR_UNSUPPORTED;
    int per_sample = kNumSMs / (b * nslice);
    per_sample = per_sample > m / 2 ? m / 2 : (per_sample < 1 ? 1 : per_sample);
    dim3 grid(per_sample, b, nslice);
    cudaError_t e = cudaFuncSetAttribute(sa_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    sa_chain_fwd_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(q);
    OGC_RETURN_LAUNCH_STATUS();
}

// Pooling epilogue of the chained forward: combines the two half-centre extremes, applies GroupNorm + ReLU of the
// last layer and emits what ogc_sa_finish emits (out (b,c_total,m) at channel offset c_offset, optional point-major
// twin, sel / ysel (b,c,m) for the backward pass).
extern "C" int ogc_sa_pool_finish(int b, int c, int m, const float *ymax_h, const float *ymin_h,
                                  const unsigned char *amax_h, const unsigned char *amin_h, const float *scale_shift,
                                  float *out, float *out_pm, int c_total, int c_offset, unsigned char *sel,
                                  float *ysel, void *stream) {
    using namespace ogc;
    if (b < 0 || c <= 0 || m <= 0 || c_offset < 0 || c_offset + c > c_total) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!ymax_h || !ymin_h || !amax_h || !amin_h || !scale_shift || !out || !sel || !ysel) return OGC_ERR_INVALID_ARG;
    if (b > 65535) return OGC_ERR_UNSUPPORTED;
    dim3 grid((m + 31) / 32, (c + 31) / 32, b);
    chain::sa_pool_finish_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        c, m, ymax_h, ymin_h, amax_h, amin_h, scale_shift, out, out_pm, c_total, c_offset, sel, ysel);
    OGC_RETURN_LAUNCH_STATUS();
}
