// Backward of the fused set-abstraction MLP on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split):
// tensor-core twins of mlp_dx_kernel / mlp_dw_kernel (mlp_bwd.cu) for the layers where the contraction dominates.
// Same warp-specialised skeleton as mlp_tc.cu (warps 0-3 epilogue, 4-7 loader, warp 8 MMA issuer + TMEM owner).
//
//   dX:  da[ci][p] = sum_co W[co][ci] dY[co][p]     M = 128 input channels (lanes), N = 64 positions, K = co
//        A = W^T rows ci, K-major (built once per CTA from W), B = dY rows p, K-major (rebuilt per tile from
//        dz / y / coef).  Epilogue thread = input channel: ReLU mask from y_prev, store dz_prev, per-channel
//        dgamma / dbeta sums in registers (no shuffles); layer 1: coalesced red.add of 32 consecutive feature
//        channels per point instead.  K > 128 is split over two launches (partial sums through dz_prev).
//   dW:  dW[co][ci] = sum_p dY[co][p] a[ci][p]      M = 128 output channels, N = C_in (<= 160), K = positions
//        both operands K-major with rows = channels (a row = 32 consecutive positions = 128 contiguous bytes of the
//        channel-major tensors: vectorised loads and stores, no transposition), 2-stage operand pipeline, ONE TMEM
//        accumulator per CTA accumulated over all its position tiles, read out once and red.add'ed into dW.
#include "mlp_dy.cuh"
#include "tcgen05.cuh"

namespace ogc {

constexpr int kTbLoaderWarps = 8;
constexpr int kTbLoaders = kTbLoaderWarps * 32;
constexpr int kTbMmaWarp = 4 + kTbLoaderWarps;
constexpr int kTbThreads = (kTbMmaWarp + 1) * 32;
constexpr int kTbDxThreads = kTbThreads + 128;   // dX: a second epilogue warpgroup (warps 13-16) takes the odd tiles
constexpr int kTbNT = 64;     // positions per tile
constexpr int kTbM = 128;

// ------------------------------------------------------------------------------------------------ dX
struct MlpDxTcParams {
    DySrc dy;                     // layer l; its channels [k0, k0+kn) are this launch's K range
    int k0, kn;
    int cin_full, row_off, rows;  // W (dy.C, cin_full); output rows = W columns [row_off, row_off+rows), rows <= 128
    const float *W;
    int add_partial, final;       // read a partial sum from dz_prev first / apply the epilogue (else store raw)
    const float *y_prev, *ss_prev, *mean_rstd_prev, *gamma_prev;
    float *dz_prev;
    double *ab_prev;
    float *dgamma_prev, *dbeta_prev;
    const int *idx;
    float *dfeat_pm;
    int N, dfeat_stride, dfeat_off;
    int n_raw, n_op, raw_stage_bytes;
};

constexpr int kTbMaxK = 128;
constexpr int kTbWCol = 2 * kTbNT;   // first TMEM column of (W^T)_hi
constexpr int kTbRowPad = 16;
constexpr int kTbXfBar = 2;

// Same skeleton as mlp_fwd_tc_kernel (mlp_tc.cu): W^T (hi, lo) stationary in TENSOR MEMORY as the A operand, raw
// y_l / dz_l rows of tile u+D streamed into a shared-memory ring with cp.async, dY built from them (GroupNorm
// backward folded into 4 coefficients per channel) as the K-major B operand, accumulator double buffered.
// The epilogue (y_prev reads, ReLU mask, per-channel sums, dz_prev stores: ~2 KB of global traffic per thread and tile at
// 16 bytes per row and instruction) is the slowest stage of this kernel (source-level profile: transformer and MMA
// warps wait for free accumulators), so TWO epilogue warpgroups alternate tiles: group g owns accumulator buffer g.
// (Scatter variant only: with 544 threads the register budget drops to 96 per thread, which the dense variant's
// epilogue -- 64 prefetched y_prev values next to the accumulator half -- does not fit without spilling; measured
// 0.437 -> 0.303 ms and 0.280 -> 0.260 ms for the two scatter launches, a net loss for the dense ones.)
template <bool SCATTER>
__global__ void __launch_bounds__(SCATTER ? kTbDxThreads : kTbThreads, 1)
mlp_dx_tc_kernel(MlpDxTcParams q) {
    constexpr int kEpGroups = SCATTER ? 2 : 1;
    constexpr int kThreads = SCATTER ? kTbDxThreads : kTbThreads;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ double gs[kGnGroups][2];
    __shared__ __align__(16) float4 coef_s[kTbMaxK];          // k1, k2, k3r, mean of this launch's K channels
    __shared__ __align__(16) float2 selgo_s[2][kTbMaxK];      // last layer: (winning position, pooled gradient) per channel of a tile

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int kn = q.kn, KB = (kn + 31) / 32, Kp = KB * 32;
    const int P = q.dy.P;
    const int ntiles = P / kTbNT;
    const int n_my = ntiles > static_cast<int>(blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_op = q.n_op, n_raw = q.n_raw, D = n_raw - 1;
    const bool synth = q.dy.dz == nullptr;       // dz from (sel, go) instead of a stored tensor

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_bytes = static_cast<uint32_t>(KB) * kTbNT * 128u;
    uint8_t *raw_base = smem + static_cast<size_t>(n_op) * 2 * a_bytes;
    auto op_hi = [&](int ob) { return smem + static_cast<size_t>(ob) * 2 * a_bytes; };
    auto raw_stage = [&](int u) { return raw_base + static_cast<size_t>(u % n_raw) * q.raw_stage_bytes; };
    const uint32_t pitch = kTbNT * 4u + kTbRowPad;
    const uint32_t dz_off = static_cast<uint32_t>(kn) * pitch;   // dz rows follow the y rows

    if (warp == kTbMmaWarp) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_full[i], kTbLoaders); mbar_init(&bar_empty[i], 1);
            mbar_init(&bar_tfull[i], 1); mbar_init(&bar_tempty[i], 128);
        }
        mbar_fence_init();
    }
    if (tid < kGnGroups * 2) (&gs[0][0])[tid] = 0.0;
    for (int c = tid; c < kn; c += kThreads)
        coef_s[c] = __ldg(reinterpret_cast<const float4 *>(q.dy.coef) + static_cast<size_t>(b) * q.dy.C + q.k0 + c);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    const int lt = tid - 128;
    const uint32_t raw_u32 = smem_u32(raw_base);
    auto issue = [&](int u) {
        const int t = blockIdx.x + u * gridDim.x;
        const uint32_t dst = raw_u32 + static_cast<uint32_t>(u % n_raw) * q.raw_stage_bytes + (lt & 15) * 16;
        const size_t base = (static_cast<size_t>(b) * q.dy.C + q.k0) * P + static_cast<size_t>(t) * kTbNT + (lt & 15) * 4;
        // 16 lanes copy one channel's 256 B; rows advance by 16 per step
        for (int c = lt >> 4; c < kn; c += kTbLoaders / 16) {
            cp_async16_s(dst + c * pitch, q.dy.y + base + static_cast<size_t>(c) * P);
            if (!synth) cp_async16_s(dst + dz_off + c * pitch, q.dy.dz + base + static_cast<size_t>(c) * P);
        }
    };
    // last layer: (sel, go) of tile u for channel lt (one centre per tile: m = tile index)
    auto load_selgo = [&](int u) {
        float2 r = make_float2(255.f, 0.f);
        if (synth && lt < kn && u < n_my) {
            const int m = blockIdx.x + u * gridDim.x;
            r.x = static_cast<float>(__ldg(q.dy.sel + (static_cast<size_t>(b) * q.dy.C + q.k0 + lt) * q.dy.M + m));
            r.y = __ldg(q.dy.go + (static_cast<size_t>(b) * q.dy.go_ctotal + q.dy.go_coff + q.k0 + lt) * q.dy.M + m);
        }
        return r;
    };

    if (warp < 4) {
        // ---- W^T -> tensor memory: thread = row r (input channel row_off + r), column k = output channel k0 + k ----
        const int r = tid;
        for (int kb = 0; kb < KB; ++kb) {
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int k = kb * 32 + j;
                const float v = (r < q.rows && k < kn) ? __ldg(q.W + static_cast<size_t>(q.k0 + k) * q.cin_full + q.row_off + r) : 0.f;
                hi[j] = tc::tf32_hi(v);
                lo[j] = tc::tf32_hi(v - hi[j]);
            }
            const uint32_t ta = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + kTbWCol + kb * 32;
            tc::tmem_st32(ta, hi);
            tc::tmem_st32(ta + Kp, lo);
        }
    } else if (warp < kTbMmaWarp) {
        for (int d = 0; d < D; ++d) {
            if (d < n_my) issue(d);
            cp_async_commit();
        }
        if (synth && lt < kn) selgo_s[0][lt] = load_selgo(0);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    if (warp >= 4 && warp < kTbMmaWarp) {
        // ================================ transformers: dY tile, rows = positions, K = channels ================
        float2 sg_next = load_selgo(1);
        for (int u = 0; u < n_my; ++u) {
            const int ob = u % n_op;
            cp_async_wait(D - 1);
            named_bar_sync(kTbXfBar, kTbLoaders);          // tile u's raw rows visible; transform u-1 finished everywhere
            if (u + D < n_my) issue(u + D);
            cp_async_commit();
            if (synth) {                                   // slot (u+1)&1 held tile u-1: free now, read after the next barrier
                if (lt < kn) selgo_s[(u + 1) & 1][lt] = sg_next;
                sg_next = load_selgo(u + 2);
            }
            mbar_wait(&bar_empty[ob], ((u / n_op) & 1) ^ 1);
            const uint32_t st = raw_u32 + static_cast<uint32_t>(u % n_raw) * q.raw_stage_bytes;
            const uint32_t a_hi = smem_u32(op_hi(ob)), a_lo = a_hi + a_bytes;
            // thread = (position, channel quad mod 4): fixed position, the channel quad advances by 4 (16 channels) per
            // step; every address is a per-thread constant plus a multiple of the loop counter
            const int p = lt & (kTbNT - 1), c0 = (lt >> 6) * 4;
            const uint32_t rd = st + c0 * pitch + p * 4;
            const uint32_t cfa = smem_u32(coef_s) + c0 * 16, sga = smem_u32(selgo_s[u & 1]) + c0 * 8;
            const uint32_t wr0 = p * 128u + ((((c0 >> 2)) ^ (p & 7)) * 16u), wr1 = p * 128u + ((((c0 >> 2) + 4) ^ (p & 7)) * 16u);
#pragma unroll 2
            for (int i = 0; i < KB * 2; ++i) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (c0 + 16 * i < kn) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float4 cf = lds_v4(cfa + i * 256 + e * 16);
                        const float y = lds_f32(rd + i * (16 * pitch) + e * pitch);
                        float dz;
                        if (synth) {
                            const float2 s2 = lds_v2(sga + i * 128 + e * 8);
                            dz = (static_cast<int>(s2.x) == p) ? s2.y : 0.f;
                        } else {
                            dz = lds_f32(rd + dz_off + i * (16 * pitch) + e * pitch);
                        }
                        v[e] = fmaf(cf.x, dz, -cf.y) - (y - cf.w) * cf.z;
                    }
                }
                float4 hi, lo;
                tc::tf32_split4(v, hi, lo);
                const uint32_t off = static_cast<uint32_t>(i >> 1) * (kTbNT * 128u) + ((i & 1) ? wr1 : wr0);
                sts_v4(a_hi + off, hi);
                sts_v4(a_lo + off, lo);
            }
            tc::fence_proxy_async();
            mbar_arrive(&bar_full[ob]);
        }
    } else if (warp == kTbMmaWarp) {
        // warp-uniform issue (see mlp_tc.cu): all lanes run the loop, one elected lane issues
        {
            const uint32_t idesc = tc::make_idesc_tf32(kTbM, kTbNT, 0, 0);
            for (int u = 0; u < n_my; ++u) {
                const int ob = u % n_op, tb = u & 1;
                mbar_wait(&bar_full[ob], (u / n_op) & 1);
                mbar_wait(&bar_tempty[tb], ((u >> 1) & 1) ^ 1);
                tc::fence_after_sync();
                const uint32_t d = tmem_base + static_cast<uint32_t>(tb * kTbNT);
                const uint64_t bh0 = tc::make_desc_sw128(smem_u32(op_hi(ob)), 16, 1024);
                const uint64_t bl0 = bh0 + (a_bytes >> 4);
                for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        const uint32_t off16 = static_cast<uint32_t>(kb) * ((kTbNT * 128u) >> 4) + static_cast<uint32_t>(s4) * 2u;
                        const uint32_t wh = tmem_base + kTbWCol + static_cast<uint32_t>((kb * 4 + s4) * 8), wl = wh + Kp;
                        tc::mma_tf32_ts_elect(d, wh, bh0 + off16, idesc, (kb | s4) ? 1u : 0u);
                        tc::mma_tf32_ts_elect(d, wh, bl0 + off16, idesc, 1);
                        tc::mma_tf32_ts_elect(d, wl, bh0 + off16, idesc, 1);
                    }
                }
                tc::mma_commit_elect(&bar_empty[ob]);
                tc::mma_commit_elect(&bar_tfull[tb]);
            }
        }
    } else {
        // ================================ epilogue: thread = input channel r ============================
        const int eg = (kEpGroups == 2 && warp > kTbMmaWarp) ? 1 : 0;   // epilogue group: warps 0-3 even tiles, warps 13-16 odd tiles
        const int r = (warp & 3) * 32 + lane;              // TMEM lane quarter of a warp = warp % 4
        const bool valid = r < q.rows;
        const bool mask = !SCATTER && q.final;
        float sc = 0.f, sh = 0.f, mu = 0.f, rs = 0.f;
        if (mask && valid) {
            const int g = r / (q.rows / kGnGroups);
            sc = __ldg(q.ss_prev + (static_cast<size_t>(b) * q.rows + r) * 2);
            sh = __ldg(q.ss_prev + (static_cast<size_t>(b) * q.rows + r) * 2 + 1);
            mu = __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2);
            rs = __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2 + 1);
        }
        double dsum = 0.0, dsumy = 0.0;
        for (int u = eg; u < n_my; u += kEpGroups) {
            const int buf = u & 1;
            const int p0 = (blockIdx.x + u * gridDim.x) * kTbNT;
            const uint32_t ta = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(buf * kTbNT);
            // the tile is finished in two halves of 32 positions; the y_prev row reads (ReLU mask) of the first half are
            // issued BEFORE waiting for the accumulator, those of the second half fly while the first is processed
            const float4 *yp = reinterpret_cast<const float4 *>(q.y_prev + (static_cast<size_t>(b) * q.rows + (valid ? r : 0)) * P + p0);
            float4 *dp = reinterpret_cast<float4 *>(q.dz_prev + (static_cast<size_t>(b) * q.rows + (valid ? r : 0)) * P + p0);
            float4 ya[8], yb[8];
            int j_lo = 0, j_hi = 0;                      // scatter: lane l holds the neighbour index of positions l, l + 32
            if (SCATTER) {
                j_lo = __ldg(q.idx + static_cast<size_t>(b) * P + p0 + lane);
                j_hi = __ldg(q.idx + static_cast<size_t>(b) * P + p0 + 32 + lane);
            } else if (mask && valid) {
#pragma unroll
                for (int j = 0; j < 8; ++j) ya[j] = __ldg(yp + j);
            }
            mbar_wait(&bar_tfull[buf], (u >> 1) & 1);
            tc::fence_after_sync();
            if (mask && valid) {
#pragma unroll
                for (int j = 0; j < 8; ++j) yb[j] = __ldg(yp + 8 + j);
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float v[32];
                tc::tmem_ld32(ta + hf * 32, v);
                if (hf == 1) {
                    tc::fence_before_sync();
                    mbar_arrive(&bar_tempty[buf]);
                }
                if (SCATTER) {
                    // 32 lanes = 32 consecutive feature channels of the same point: one coalesced red per position
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int pt = __shfl_sync(OGC_FULL_MASK, hf ? j_hi : j_lo, j);
                        if (valid) atomicAdd(q.dfeat_pm + (static_cast<size_t>(b) * q.N + pt) * q.dfeat_stride + q.dfeat_off + r, v[j]);
                    }
                    continue;
                }
                if (!valid) continue;
                if (q.add_partial) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 o = dp[hf * 8 + j];
                        v[4 * j] += o.x; v[4 * j + 1] += o.y; v[4 * j + 2] += o.z; v[4 * j + 3] += o.w;
                    }
                }
                if (q.final) {
                    float s = 0.f, sy = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 y4 = hf ? yb[j] : ya[j];
                        const float yy[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float g = fmaf(sc, yy[e], sh) > 0.f ? v[4 * j + e] : 0.f;
                            v[4 * j + e] = g;
                            s += g;
                            sy = fmaf(g, (yy[e] - mu) * rs, sy);
                        }
                    }
                    dsum += static_cast<double>(s);
                    dsumy += static_cast<double>(sy);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) dp[hf * 8 + j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
        if (mask && valid) {
            atomicAdd(q.dbeta_prev + r, static_cast<float>(dsum));
            atomicAdd(q.dgamma_prev + r, static_cast<float>(dsumy));
            const int g = r / (q.rows / kGnGroups);
            const double gm = static_cast<double>(__ldg(q.gamma_prev + r));
            atomicAdd(&gs[g][0], gm * dsum);
            atomicAdd(&gs[g][1], gm * dsumy);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (!SCATTER && q.final && tid < kGnGroups * 2)        // both epilogue groups have added their group sums
        atomicAdd(q.ab_prev + static_cast<size_t>(b) * kGnGroups * 2 + tid, (&gs[0][0])[tid]);
    if (warp == kTbMmaWarp) tc::tmem_dealloc(tmem_base, 512);
}

static inline bool tc_dx_plan(int kn, bool synth, int &n_raw, int &n_op, int &stage, size_t &smem) {
    const int KB = (kn + 31) / 32;
    const size_t op = static_cast<size_t>(KB) * kTbNT * 128 * 2;
    stage = kn * (kTbNT * 4 + kTbRowPad) * (synth ? 1 : 2);
    stage = (stage + 127) & ~127;
    const size_t budget = static_cast<size_t>(kMaxSmemPerCta) - 8 * 1024 - 1024;
    for (n_op = 2; n_op >= 1; --n_op) {
        if (op * n_op + 2 * static_cast<size_t>(stage) > budget) continue;
        n_raw = static_cast<int>((budget - op * n_op) / stage);
        n_raw = n_raw > 4 ? 4 : n_raw;
        smem = op * n_op + static_cast<size_t>(n_raw) * stage + 1024;
        return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------ dW
struct MlpDwTcParams {
    DySrc dy;                     // layer l: rows of dW = dy.C
    int Cin, B;
    const float *y_prev, *ss_prev;
    const float *xyz, *new_xyz, *feat_pm;
    const int *idx;
    int N, Cf;
    float *dW;                    // (Cout, Cin) accumulated with red.add
    int n_raw, raw_stage_bytes;
};

constexpr int kDwTile = 32;       // positions per pipeline stage = one 128-byte row of every operand
constexpr int kDwPitch = kDwTile * 4 + 16;   // raw row pitch (16-byte reads of consecutive rows at a 4-bank skew)
constexpr int kDwAccCols = 192;   // accumulator: up to 160 columns; dY operand buffers follow
constexpr int kDwMaxCin = 160;

// K = positions.  Per stage (32 positions of one sample):
//   raw ring (cp.async, D stages ahead): y_l and dz_l rows of this M block's channels, and the layer's input -- the
//       previous layer's pre-norm rows, or the gathered feature rows + xyz of the stage's 32 neighbours
//   A = dY (128 channels x 32 positions, hi and lo): thread = channel builds its row from the raw rows and its four
//       GroupNorm-backward coefficients and writes it straight to TENSOR MEMORY (tcgen05.st) -- no shared memory
//   B = a  (C_in rows x 32 positions, K-major 128 B rows, hi and lo) in shared memory, double buffered
//   ONE accumulator (128 x C_in) per CTA in TMEM over all its stages, read out once and red.add'ed into dW.
// a-tile row order: dense: ci; gather: [feat, xyz].
template <bool GATHER>
__global__ void __launch_bounds__(kTbThreads, 1)
mlp_dw_tc_kernel(MlpDwTcParams q) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float2 ss_s[2][kDwMaxCin];      // GroupNorm (scale, shift) of the input channels for the stage's sample
    __shared__ int idx_s[2][kDwTile];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mb = blockIdx.y;
    const int P = q.dy.P, Cout = q.dy.C;
    const int tiles_per_sample = P / kDwTile;
    const int total = q.B * tiles_per_sample;
    const int n_my = total > static_cast<int>(blockIdx.x) ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nrows_a = GATHER ? q.Cf + 3 : q.Cin;              // valid rows of the a tile
    const int n_mma = ((nrows_a + 15) / 16) * 16;                // MMA N (rows of B), <= 160
    const int a_rows_alloc = ((n_mma + 7) / 8) * 8;
    const int m_rows = min(kTbM, Cout - mb * kTbM);              // valid channels of this M block
    const int n_raw = q.n_raw, D = n_raw - 1;
    const bool synth = q.dy.dz == nullptr;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_bytes = (static_cast<uint32_t>(a_rows_alloc) * 128u + 1023u) & ~1023u;
    auto a_hi = [&](int ob) { return smem + static_cast<size_t>(ob) * 2 * a_bytes; };
    uint8_t *raw_base = smem + 4 * static_cast<size_t>(a_bytes);
    auto raw_stage = [&](int u) { return raw_base + static_cast<size_t>(u % n_raw) * q.raw_stage_bytes; };
    const uint32_t dz_off = static_cast<uint32_t>(m_rows) * kDwPitch;
    const uint32_t in_off = dz_off * (synth ? 1u : 2u);                      // input rows follow y (and dz)
    const uint32_t g_pitch = static_cast<uint32_t>(q.Cf) * 4u + 16u;         // gather: one neighbour's features
    const uint32_t xyz_off = in_off + kDwTile * g_pitch;

    if (warp == kTbMmaWarp) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&bar_full[s], kTbLoaders); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        mbar_fence_init();
    }
    for (uint32_t e = tid * 16u; e < 4 * a_bytes; e += kTbThreads * 16u)     // padding rows stay zero for the CTA's lifetime
        *reinterpret_cast<float4 *>(smem + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    const int lt = tid - 128;
    auto stage_of = [&](int u, int &b, int &p0) {
        const int w = blockIdx.x + u * gridDim.x;
        b = w / tiles_per_sample;
        p0 = (w - b * tiles_per_sample) * kDwTile;
    };
    const uint32_t raw_u32 = smem_u32(raw_base);
    auto issue = [&](int u) {
        int b, p0;
        stage_of(u, b, p0);
        const uint32_t st = raw_u32 + static_cast<uint32_t>(u % n_raw) * q.raw_stage_bytes;
        const size_t ybase = (static_cast<size_t>(b) * Cout + mb * kTbM) * P + p0 + (lt & 7) * 4;
        // 8 lanes copy one channel's 128 B; rows advance by 32 per step
        for (int c = lt >> 3; c < m_rows; c += kTbLoaders / 8) {
            cp_async16_s(st + c * kDwPitch + (lt & 7) * 16, q.dy.y + ybase + static_cast<size_t>(c) * P);
            if (!synth) cp_async16_s(st + dz_off + c * kDwPitch + (lt & 7) * 16, q.dy.dz + ybase + static_cast<size_t>(c) * P);
        }
        if (GATHER) {
            const int *js = idx_s[u & 1];
            if (lane < (q.Cf >> 2)) {
                for (int p = warp - 4; p < kDwTile; p += kTbLoaderWarps)
                    cp_async16_s(st + in_off + p * g_pitch + lane * 16, q.feat_pm + (static_cast<size_t>(b) * q.N + js[p]) * q.Cf + lane * 4);
            }
            if (lt < kDwTile * 3) {
                const int p = lt / 3, c = lt - p * 3;
                cp_async4_s(st + xyz_off + lt * 4, q.xyz + (static_cast<size_t>(b) * q.N + js[p]) * 3 + c);
            } else if (lt < kDwTile * 3 + 3) {
                const int c = lt - kDwTile * 3;
                cp_async4_s(st + xyz_off + lt * 4, q.new_xyz + (static_cast<size_t>(b) * q.dy.M + p0 / q.dy.S) * 3 + c);
            }
        } else {
            const size_t abase = static_cast<size_t>(b) * q.Cin * P + p0 + (lt & 7) * 4;
            for (int c = lt >> 3; c < q.Cin; c += kTbLoaders / 8)
                cp_async16_s(st + in_off + c * kDwPitch + (lt & 7) * 16, q.y_prev + abase + static_cast<size_t>(c) * P);
        }
    };
    auto load_idx = [&](int u) {
        int b, p0;
        stage_of(u, b, p0);
        return (GATHER && lt < kDwTile && u < n_my) ? __ldg(q.idx + static_cast<size_t>(b) * P + p0 + lt) : 0;
    };
    auto load_ss = [&](int u) {      // thread lt < Cin: (scale, shift) of input channel lt for stage u's sample
        int b, p0;
        stage_of(u, b, p0);
        return (!GATHER && lt < q.Cin && u < n_my) ? __ldg(reinterpret_cast<const float2 *>(q.ss_prev) + static_cast<size_t>(b) * q.Cin + lt)
                                                  : make_float2(0.f, 0.f);
    };
    // this thread's dY row: channel co (TMEM lane), 16 of the stage's 32 positions
    const int lq = warp & 3;                                     // TMEM lane quarter this warp may access
    const int chalf = (warp - 4) >> 2;                           // 0: positions 0-15, 1: positions 16-31
    const int row = lq * 32 + lane, co = mb * kTbM + row;
    auto load_coef = [&](int u) {
        int b, p0;
        stage_of(u, b, p0);
        return (co < Cout && u < n_my) ? __ldg(reinterpret_cast<const float4 *>(q.dy.coef) + static_cast<size_t>(b) * Cout + co)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto load_selgo = [&](int u) {
        float2 r = make_float2(255.f, 0.f);
        if (synth && co < Cout && u < n_my) {
            int b, p0;
            stage_of(u, b, p0);
            const int m = p0 / q.dy.S;
            r.x = static_cast<float>(__ldg(q.dy.sel + (static_cast<size_t>(b) * Cout + co) * q.dy.M + m));
            r.y = __ldg(q.dy.go + (static_cast<size_t>(b) * q.dy.go_ctotal + q.dy.go_coff + co) * q.dy.M + m);
        }
        return r;
    };

    if (warp >= 4 && warp < kTbMmaWarp) {
        // ================================ transformers ================================
        for (int d = 0; d < D; ++d) {
            if (GATHER) {
                const int j = load_idx(d);
                if (lt < kDwTile) idx_s[d & 1][lt] = j;
                named_bar_sync(kTbXfBar, kTbLoaders);
            }
            if (d < n_my) issue(d);
            cp_async_commit();
        }
        if (GATHER) {
            const int j = load_idx(D);
            named_bar_sync(kTbXfBar, kTbLoaders);
            if (lt < kDwTile) idx_s[D & 1][lt] = j;
        } else if (lt < q.Cin) {
            ss_s[0][lt] = load_ss(0);
        }
        float4 cf = load_coef(0);
        float2 sg = load_selgo(0);
        for (int u = 0; u < n_my; ++u) {
            const int ob = u & 1;
            cp_async_wait(D - 1);
            named_bar_sync(kTbXfBar, kTbLoaders);          // stage u landed everywhere; transform u-1 finished everywhere
            if (u + D < n_my) issue(u + D);
            cp_async_commit();
            const int jn = load_idx(u + D + 1);
            const float2 ssn = load_ss(u + 1);
            const float4 cfn = load_coef(u + 1);
            const float2 sgn = load_selgo(u + 1);
            mbar_wait(&bar_empty[ob], ((u >> 1) & 1) ^ 1);  // the MMAs that read operand buffers `ob` are done
            const uint32_t st = raw_u32 + static_cast<uint32_t>(u % n_raw) * q.raw_stage_bytes;
            int b, p0;
            stage_of(u, b, p0);
            // ---- dY row -> tensor memory ----
            {
                float hi[16], lo[16];
                if (row < m_rows) {
                    const uint32_t yr = st + row * kDwPitch + chalf * 64, zr = yr + dz_off;
                    const int s0 = p0 % q.dy.S + chalf * 16;        // position of column 0 inside its centre's group
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 y4 = lds_v4(yr + i * 16);
                        float4 z4;
                        if (synth) {
                            const int sl = static_cast<int>(sg.x) - s0 - 4 * i;
                            z4 = make_float4(sl == 0 ? sg.y : 0.f, sl == 1 ? sg.y : 0.f, sl == 2 ? sg.y : 0.f, sl == 3 ? sg.y : 0.f);
                        } else {
                            z4 = lds_v4(zr + i * 16);
                        }
                        const float yy[4] = {y4.x, y4.y, y4.z, y4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            tc::tf32_split(fmaf(cf.x, zz[e], -cf.y) - (yy[e] - cf.w) * cf.z, hi[4 * i + e], lo[4 * i + e]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) hi[i] = lo[i] = 0.f;
                }
                const uint32_t ta = tmem_base + (static_cast<uint32_t>(lq * 32) << 16) + kDwAccCols + ob * 64 + chalf * 16;
                tc::tmem_st16(ta, hi);
                tc::tmem_st16(ta + 32, lo);
            }
            // ---- a tile: row = input channel, 128 B = the stage's 32 positions ----
            const uint32_t ah = smem_u32(a_hi(ob)), al = ah + a_bytes;
            if (GATHER) {
                // lane = channel (mod 32), warp = position quad: 4 scalar reads down the gathered rows, one 16-byte
                // chunk of the channel's operand row
                const int pq = warp - 4;
                for (int c = lane; c < q.Cf; c += 32) {
                    const uint32_t rd = st + in_off + (pq * 4) * g_pitch + c * 4;
                    const float v[4] = {lds_f32(rd), lds_f32(rd + g_pitch), lds_f32(rd + 2 * g_pitch), lds_f32(rd + 3 * g_pitch)};
                    float4 hi, lo;
                    tc::tf32_split4(v, hi, lo);
                    const uint32_t off = static_cast<uint32_t>(c) * 128u + static_cast<uint32_t>((pq ^ (c & 7)) * 16);
                    sts_v4(ah + off, hi);
                    sts_v4(al + off, lo);
                }
                if (lt < kDwTile * 3) {
                    const int c = lt >> 5, p = lt & 31;
                    const float v = lds_f32(st + xyz_off + (p * 3 + c) * 4) - lds_f32(st + xyz_off + (kDwTile * 3 + c) * 4);
                    float hi, lo;
                    tc::tf32_split(v, hi, lo);
                    const uint32_t off = tc::sw128_offset(q.Cf + c, p);
                    sts_f32(ah + off, hi);
                    sts_f32(al + off, lo);
                }
            } else {
                // thread = (row mod 32, position quad): rows advance by 32 per step
                const int pq = lt & 7;
                const uint32_t ssa = smem_u32(ss_s[u & 1]);
                for (int r = lt >> 3; r < q.Cin; r += kTbLoaders / 8) {
                    const float4 x = lds_v4(st + in_off + r * kDwPitch + pq * 16);
                    const float2 s2 = lds_v2(ssa + r * 8);
                    const float v[4] = {fmaxf(fmaf(s2.x, x.x, s2.y), 0.f), fmaxf(fmaf(s2.x, x.y, s2.y), 0.f),
                                        fmaxf(fmaf(s2.x, x.z, s2.y), 0.f), fmaxf(fmaf(s2.x, x.w, s2.y), 0.f)};
                    float4 hi, lo;
                    tc::tf32_split4(v, hi, lo);
                    const uint32_t off = static_cast<uint32_t>(r) * 128u + static_cast<uint32_t>((pq ^ (r & 7)) * 16);
                    sts_v4(ah + off, hi);
                    sts_v4(al + off, lo);
                }
            }
            tc::fence_proxy_async();
            tc::fence_before_sync();
            mbar_arrive(&bar_full[ob]);
            if (GATHER) { if (lt < kDwTile) idx_s[(u + D + 1) & 1][lt] = jn; }
            else if (lt < q.Cin) ss_s[(u + 1) & 1][lt] = ssn;     // read after the next barrier
            cf = cfn;
            sg = sgn;
        }
    } else if (warp == kTbMmaWarp) {
        {
            const uint32_t idesc = tc::make_idesc_tf32(kTbM, n_mma, 0, 0);
            uint32_t acc = 0;
            for (int u = 0; u < n_my; ++u) {
                const int ob = u & 1;
                mbar_wait(&bar_full[ob], (u >> 1) & 1);
                tc::fence_after_sync();
                const uint64_t ahd0 = tc::make_desc_sw128(smem_u32(a_hi(ob)), 16, 1024);
                const uint64_t ald0 = ahd0 + (a_bytes >> 4);
                const uint32_t dh0 = tmem_base + kDwAccCols + ob * 64, dl0 = dh0 + 32;
#pragma unroll
                for (int s = 0; s < kDwTile / 8; ++s) {     // 32 positions = 4 K-steps: 32 B of every a row, 8 dY columns
                    tc::mma_tf32_ts_elect(tmem_base, dh0 + s * 8, ahd0 + static_cast<uint32_t>(s * 2), idesc, acc);
                    tc::mma_tf32_ts_elect(tmem_base, dh0 + s * 8, ald0 + static_cast<uint32_t>(s * 2), idesc, 1);
                    tc::mma_tf32_ts_elect(tmem_base, dl0 + s * 8, ahd0 + static_cast<uint32_t>(s * 2), idesc, 1);
                    acc = 1;
                }
                tc::mma_commit_elect(&bar_empty[ob]);
            }
            tc::mma_commit_elect(&bar_done);
        }
    } else {
        // epilogue: wait for every MMA of this CTA, then lane = output channel adds its row of dW
        mbar_wait(&bar_done, 0);
        tc::fence_after_sync();
        const int eco = mb * kTbM + tid;
        if (n_my > 0) {
            for (int c0 = 0; c0 < n_mma; c0 += 32) {
                float h[32];
                tc::tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c0), h);
                if (eco >= Cout) continue;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int c = c0 + j;
                    if (c >= nrows_a) continue;
                    const int ci = GATHER ? (c < q.Cf ? 3 + c : c - q.Cf) : c;
                    atomicAdd(q.dW + static_cast<size_t>(eco) * q.Cin + ci, h[j]);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kTbMmaWarp) tc::tmem_dealloc(tmem_base, 512);
}

static inline bool tc_dw_plan(int cout, int cin, bool gather, bool synth, int &n_raw, int &stage, size_t &smem) {
    const int nrows_a = cin;                              // gather: Cf + 3 == cin
    const int n_mma = ((nrows_a + 15) / 16) * 16, a_rows = ((n_mma + 7) / 8) * 8;
    const size_t a_bytes = (static_cast<size_t>(a_rows) * 128 + 1023) & ~static_cast<size_t>(1023);
    const int m_rows = cout < kTbM ? cout : kTbM;
    stage = m_rows * kDwPitch * (synth ? 1 : 2);
    stage += gather ? kDwTile * ((cin - 3) * 4 + 16) + 512 : cin * kDwPitch;
    stage = (stage + 127) & ~127;
    const size_t budget = static_cast<size_t>(kMaxSmemPerCta) - 8 * 1024 - 1024;
    if (4 * a_bytes + 2 * static_cast<size_t>(stage) > budget) return false;
    n_raw = static_cast<int>((budget - 4 * a_bytes) / stage);
    n_raw = n_raw > 4 ? 4 : n_raw;
    smem = 4 * a_bytes + static_cast<size_t>(n_raw) * stage + 1024;
    return true;
}

}  // namespace ogc

// Tensor-core variant of ogc_sa_mlp_layer_dx (same argument meaning).  nsample == 64, rows <= 128, cout a multiple
// of 4; cout > 128 runs as two launches whose partial sums travel through dz_prev.  Scatter mode needs cout <= 128.
extern "C" int ogc_sa_mlp_layer_dx_tc(int b, int n, int m, int nsample, int cout, int cin_full, int row_off, int rows,
                                      const float *dz, const float *go, int go_ctotal, int go_coff,
                                      const unsigned char *sel, const float *y, const float *coef, const float *w,
                                      const float *y_prev, const float *ss_prev, const float *mean_rstd_prev,
                                      const float *gamma_prev, float *dz_prev, double *ab_prev, float *dgamma_prev,
                                      float *dbeta_prev, const int *idx, float *dfeat_pm, int dfeat_stride,
                                      int dfeat_off, void *stream) {
    using namespace ogc;
    if (b < 0 || m <= 0 || nsample <= 0 || cout <= 0 || rows <= 0 || row_off < 0 || row_off + rows > cin_full || !w)
        return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    const bool scatter = dfeat_pm != nullptr;
    if (nsample != kTbNT || rows > kTbM || b > 65535 || cout > 256 || cout < 32 || (scatter && cout > 128))
        return OGC_ERR_UNSUPPORTED;
    MlpDxTcParams q;
    int rc = fill_dy(q.dy, cout, m, nsample, dz, go, go_ctotal, go_coff, sel, y, coef);
    if (rc != OGC_OK) return rc;
    if (scatter) {
        if (!idx) return OGC_ERR_INVALID_ARG;
    } else {
        if (!y_prev || !ss_prev || !mean_rstd_prev || !gamma_prev || !dz_prev || !ab_prev || !dgamma_prev || !dbeta_prev)
            return OGC_ERR_INVALID_ARG;
        if (rows % 16 != 0) return OGC_ERR_UNSUPPORTED;
    }
    q.cin_full = cin_full; q.row_off = row_off; q.rows = rows; q.W = w;
    q.y_prev = y_prev; q.ss_prev = ss_prev; q.mean_rstd_prev = mean_rstd_prev; q.gamma_prev = gamma_prev;
    q.dz_prev = dz_prev; q.ab_prev = ab_prev; q.dgamma_prev = dgamma_prev; q.dbeta_prev = dbeta_prev;
    q.idx = idx; q.dfeat_pm = dfeat_pm; q.N = n; q.dfeat_stride = dfeat_stride; q.dfeat_off = dfeat_off;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int per_sample = kNumSMs / b;
    per_sample = per_sample > m ? m : (per_sample < 1 ? 1 : per_sample);
    dim3 grid(per_sample, b);
    for (int k0 = 0; k0 < cout; k0 += kTbM) {
        q.k0 = k0;
        q.kn = cout - k0 < kTbM ? cout - k0 : kTbM;
        q.add_partial = k0 > 0;
        q.final = k0 + kTbM >= cout;
        size_t smem = 0;
        if (!tc_dx_plan(q.kn, dz == nullptr, q.n_raw, q.n_op, q.raw_stage_bytes, smem)) return OGC_ERR_UNSUPPORTED;
        cudaError_t e;
        if (scatter) {
            e = cudaFuncSetAttribute(mlp_dx_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return static_cast<int>(e);
            mlp_dx_tc_kernel<true><<<grid, kTbDxThreads, smem, st>>>(q);
        } else {
            e = cudaFuncSetAttribute(mlp_dx_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return static_cast<int>(e);
            mlp_dx_tc_kernel<false><<<grid, kTbThreads, smem, st>>>(q);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    return OGC_OK;
}

// Tensor-core variant of ogc_sa_mlp_layer_dw (same argument meaning).  nsample == 64, cin (+3 gathered) <= 160.
extern "C" int ogc_sa_mlp_layer_dw_tc(int b, int n, int m, int nsample, int cout, int cin, int gather, const float *dz,
                                      const float *go, int go_ctotal, int go_coff, const unsigned char *sel,
                                      const float *y, const float *coef, const float *y_prev, const float *ss_prev,
                                      const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                                      float *dw, void *stream) {
    using namespace ogc;
    if (b < 0 || m <= 0 || nsample <= 0 || cout <= 0 || cin <= 0 || !dw) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (nsample != kTbNT || cin > 160 || cin < 16 || cout < 32 || cout > 256) return OGC_ERR_UNSUPPORTED;
    MlpDwTcParams q;
    int rc = fill_dy(q.dy, cout, m, nsample, dz, go, go_ctotal, go_coff, sel, y, coef);
    if (rc != OGC_OK) return rc;
    if (gather && (!xyz || !new_xyz || !idx || cin < 3 || (cin > 3 && !feat_pm))) return OGC_ERR_INVALID_ARG;
    if (!gather && (!y_prev || !ss_prev)) return OGC_ERR_INVALID_ARG;
    q.Cin = cin; q.B = b; q.y_prev = y_prev; q.ss_prev = ss_prev; q.xyz = xyz; q.new_xyz = new_xyz;
    q.feat_pm = feat_pm; q.idx = idx; q.N = n; q.Cf = cin - 3; q.dW = dw;
    size_t smem = 0;
    if (gather && (cin - 3) % 4 != 0) return OGC_ERR_UNSUPPORTED;
    if (!tc_dw_plan(cout, cin, gather != 0, dz == nullptr, q.n_raw, q.raw_stage_bytes, smem)) return OGC_ERR_UNSUPPORTED;
    const int mblocks = (cout + kTbM - 1) / kTbM;
    const int total = b * m * (nsample / kDwTile);
    int gx = kNumSMs / mblocks;
    gx = gx > total ? total : (gx < 1 ? 1 : gx);
    dim3 grid(gx, mblocks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (gather) {
        e = cudaFuncSetAttribute(mlp_dw_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        mlp_dw_tc_kernel<true><<<grid, kTbThreads, smem, st>>>(q);
    } else {
        e = cudaFuncSetAttribute(mlp_dw_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        mlp_dw_tc_kernel<false><<<grid, kTbThreads, smem, st>>>(q);
    }
    OGC_RETURN_LAUNCH_STATUS();
}
