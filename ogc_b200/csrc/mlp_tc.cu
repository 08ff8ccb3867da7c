// Fused set-abstraction MLP layer on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 split).
//
// Same contract as mlp_fwd_kernel in mlp.cu (one SharedMLP layer: y = W a, GroupNorm statistics, optional
// max/min over nsample), for the layers where the contraction dominates (C_out >= 64, K a multiple of 4
// after taking the three xyz input channels of a gathered layer out of the GEMM).
//
//   D[co][p] = sum_k W[co][k] a[k][p]      M = 128 output channels (TMEM lanes), N = 64 positions (one centre,
//                                          TMEM columns), K = input channels, fp32 accumulators in TMEM
//   fp32-grade accuracy from tf32 tensor cores: x = hi + lo (hi = rna_tf32(x), lo = rna_tf32(x - hi));
//   D += W_hi a_hi + W_hi a_lo + W_lo a_hi      (error ~3x an fp32 GEMM: tests/test_gpu_tcgen05.py)
//   issued as TWO instructions per k-step: the hi and lo operand rows of a K block are stacked (64 + 64 rows), so
//   W_hi x [a_hi ; a_lo] is one N = 128 MMA into 128 accumulator columns and W_lo x a_hi one N = 64 MMA into the
//   first 64; the epilogue adds the two column halves.  (The source-level profile showed every warp waiting for the
//   tensor pipe at ~140 cycles per tiny M128 N64 K8 instruction: fewer, larger instructions.)
//
// One persistent CTA per SM, warp-specialised, every stage asynchronous to the next:
//   W (hi and lo)  lives in TENSOR MEMORY for the CTA's lifetime (2 x K columns next to the accumulators) and is
//              the MMA's A operand from there: no shared memory, and no shared-memory bandwidth per MMA
//   warps 4-11 "transformers": (1) cp.async the RAW input of tile u+D (rows of the previous layer's pre-norm
//              output, or the gathered feature rows + xyz of the tile's 64 neighbours) into a ring of raw stages,
//              D tiles ahead -- the HBM/L2 latency is covered by copies in flight, not by registers;
//              (2) turn the raw stage of tile u into the B operand: GroupNorm+ReLU of the previous layer on the
//              fly, hi/lo split, 128B-swizzled K-major rows, double buffered against the MMAs
//   warp 12    one thread issues the MMAs (A from TMEM, B from shared memory) and commits to mbarriers
//   warps 0-3  epilogue: tcgen05.ld their 32 TMEM lanes (= 32 channels) x 64 columns; thread = channel, so
//              the GroupNorm sums, the max/min/arg over the 64 samples and the xyz contribution of a gathered
//              layer are plain per-thread loops (no shuffles); stores y channel-major (16 x 16 B per thread);
//              the accumulator is double buffered so this overlaps the next tile's MMAs
#include <cstdlib>

#include "mlp_common.cuh"
#include "tcgen05.cuh"

namespace ogc {

constexpr int kTcXfWarps = 8;                 // transformer warps
constexpr int kTcXf = kTcXfWarps * 32;
constexpr int kTcMmaWarp = 4 + kTcXfWarps;
constexpr int kTcThreads = (kTcMmaWarp + 1) * 32;
constexpr int kTcNT = 64;          // positions per tile == nsample
constexpr int kTcM = 128;          // output channels per CTA (one M block)
constexpr int kTcMaxK = 128;       // contraction length (TMEM: 128 accumulator + 2 * 128 weight columns)
constexpr int kTcAccCols = 2 * kTcNT;   // accumulator columns per buffer: [W_hi a_hi + W_lo a_hi | W_hi a_lo]
constexpr int kTcWCol = 2 * kTcAccCols; // first TMEM column of W_hi
constexpr int kTcRowPad = 16;      // raw rows are padded by 16 B: 16-byte row reads at a 4-bank skew
constexpr int kTcXfBar = 2;        // named barrier of the transformer warps
constexpr int kTcStagePitch = kTcNT + 4;   // floats per row of the epilogue staging tile
constexpr int kTcStageBytes = kTcM * kTcStagePitch * 4;

struct MlpTcParams {
    MlpFwdParams f;
    const float *W;       // (Cout, Cin) row-major (NOT transposed)
    int k_off, K;         // tensor-core K range: W columns [k_off, k_off+K); gather: k_off = 3, K = Cf
    int n_raw, n_op;      // raw ring stages (2..4), operand buffers (1..2)
    int raw_stage_bytes;
};

template <bool GATHER, bool LAST>
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_fwd_tc_kernel(MlpTcParams q) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ double gs[kGnGroups][2];
    __shared__ float rel[2][kTcNT][4];   // centred xyz of the tile's positions (gather), double buffered
    __shared__ __align__(16) float2 ss_s[kTcMaxK];     // GroupNorm (scale, shift) of the input channels (dense)
    __shared__ int idx_s[2][kTcNT];      // neighbour indices of the tile whose copies are issued next (gather)

    const MlpFwdParams &f = q.f;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, mb = blockIdx.z;
    const int K = q.K, KB = (K + 31) / 32, Kp = KB * 32;
    const int Cout = f.Cout, Cin = f.Cin, P = f.P;
    const int ntiles = P / kTcNT;
    const int n_my = ntiles > static_cast<int>(blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_op = q.n_op, n_raw = q.n_raw, D = n_raw - 1;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_bytes = static_cast<uint32_t>(KB) * kTcNT * 128u;      // one of hi / lo
    uint8_t *raw_base = smem + static_cast<size_t>(n_op) * 2 * a_bytes;
    // epilogue staging: the accumulator tile is (channel = thread) x 64 positions; storing it row by row from the
    // registers puts 16 bytes into 32 different rows per instruction (ncu: 2-2.5x sector inflation, l1tex / lts the
    // busiest units).  Through this buffer every store instruction writes two full 256-byte row segments.
    float *stage_y = reinterpret_cast<float *>(raw_base + static_cast<size_t>(n_raw) * q.raw_stage_bytes);
    auto op_hi = [&](int ob) { return smem + static_cast<size_t>(ob) * 2 * a_bytes; };
    auto raw_stage = [&](int u) { return raw_base + static_cast<size_t>(u % n_raw) * q.raw_stage_bytes; };
    const uint32_t g_pitch = static_cast<uint32_t>(K) * 4u + kTcRowPad;     // gather: row = one neighbour's K features
    const uint32_t d_pitch = kTcNT * 4u + kTcRowPad;                        // dense: row = one channel's 64 positions
    const uint32_t xyz_off = kTcNT * g_pitch;                               // gather: 64 x 3 xyz, then the centre

    if (warp == kTcMmaWarp) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_full[i], kTcXf); mbar_init(&bar_empty[i], 1);
            mbar_init(&bar_tfull[i], 1); mbar_init(&bar_tempty[i], 128);
        }
        mbar_fence_init();
    }
    if (tid < kGnGroups * 2) (&gs[0][0])[tid] = 0.0;
    if (!GATHER)
        for (int c = tid; c < K; c += kTcThreads)
            ss_s[c] = __ldg(reinterpret_cast<const float2 *>(f.ss_prev) + static_cast<size_t>(b) * Cin + c);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    // ---- raw-stage copies of this CTA's u-th tile (transformer threads; no wait) ----
    const int lt = tid - 128;
    const uint32_t raw_u32 = smem_u32(raw_base);
    auto issue = [&](int u) {
        const int t = blockIdx.x + u * gridDim.x;
        const uint32_t st = raw_u32 + static_cast<uint32_t>(u % n_raw) * q.raw_stage_bytes;
        if (GATHER) {
            const int *js = idx_s[u & 1];
            // lane = 16-byte chunk of a feature row (<= 32 chunks), warp = neighbour: coalesced row reads
            if (lane < (K >> 2)) {
                for (int p = warp - 4; p < kTcNT; p += kTcXfWarps)
                    cp_async16_s(st + p * g_pitch + lane * 16, f.feat_pm + (static_cast<size_t>(b) * f.N + js[p]) * f.Cf + lane * 4);
            }
            if (lt < kTcNT * 3) {
                const int p = lt / 3, c = lt - p * 3;
                cp_async4_s(st + xyz_off + lt * 4, f.xyz + (static_cast<size_t>(b) * f.N + js[p]) * 3 + c);
            } else if (lt < kTcNT * 3 + 3) {
                const int c = lt - kTcNT * 3;
                cp_async4_s(st + xyz_off + lt * 4, f.new_xyz + (static_cast<size_t>(b) * f.M + t) * 3 + c);
            }
        } else {
            // thread = (row mod 16, chunk): 16 lanes copy one channel's 256 B, rows advance by 16 per step
            const float *src = f.y_prev + static_cast<size_t>(b) * Cin * P + static_cast<size_t>(t) * kTcNT + (lt & 15) * 4;
            const uint32_t dst = st + (lt & 15) * 16;
            for (int c = lt >> 4; c < K; c += kTcXf / 16)
                cp_async16_s(dst + c * d_pitch, src + static_cast<size_t>(c) * P);
        }
    };
    auto load_idx = [&](int u) {     // neighbour indices of tile u (threads lt < 64), -1 past the end
        return (lt < kTcNT && u < n_my)
                   ? __ldg(f.idx + static_cast<size_t>(b) * P + static_cast<size_t>(blockIdx.x + u * gridDim.x) * kTcNT + lt) : 0;
    };

    if (warp < 4) {
        // ---- W -> tensor memory: thread = output channel = TMEM lane; columns [kTcWCol, +Kp) hi, then lo ----
        const int co = mb * kTcM + tid;
        const float *wrow = q.W + static_cast<size_t>(co < Cout ? co : 0) * Cin + q.k_off;
        for (int kb = 0; kb < KB; ++kb) {
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int k = kb * 32 + j;
                const float v = (co < Cout && k < K) ? __ldg(wrow + k) : 0.f;
                hi[j] = tc::tf32_hi(v);
                lo[j] = tc::tf32_hi(v - hi[j]);
            }
            const uint32_t ta = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + kTcWCol + kb * 32;
            tc::tmem_st32(ta, hi);
            tc::tmem_st32(ta + Kp, lo);
        }
    } else if (warp < kTcMmaWarp) {
        // ---- prologue: the first D tiles' copies ----
        for (int d = 0; d < D; ++d) {
            if (GATHER) {
                const int j = load_idx(d);
                if (lt < kTcNT) idx_s[d & 1][lt] = j;
                named_bar_sync(kTcXfBar, kTcXf);
            }
            if (d < n_my) issue(d);
            cp_async_commit();
        }
        if (GATHER) {
            const int j = load_idx(D);
            named_bar_sync(kTcXfBar, kTcXf);             // issue(D-2) has read the slot
            if (lt < kTcNT) idx_s[D & 1][lt] = j;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    if (warp >= 4 && warp < kTcMmaWarp) {
        // ================================ transformers ================================
        for (int u = 0; u < n_my; ++u) {
            const int ob = u % n_op, tb = u & 1;
            cp_async_wait(D - 1);                          // this thread's copies of tile u have landed
            named_bar_sync(kTcXfBar, kTcXf);               // ... everyone's; and transform u-1 is finished everywhere
            if (u + D < n_my) issue(u + D);                // into the stage tile u-1 just vacated
            cp_async_commit();
            int jn = 0;
            if (GATHER) jn = load_idx(u + D + 1);
            mbar_wait(&bar_empty[ob], ((u / n_op) & 1) ^ 1);            // MMAs of the tile that used this buffer are done
            const uint32_t st = raw_u32 + static_cast<uint32_t>(u % n_raw) * q.raw_stage_bytes;
            const uint32_t a_hi = smem_u32(op_hi(ob)), a_lo = a_hi + kTcNT * 128u;   // K block = [64 hi rows | 64 lo rows]
            if (GATHER) {
                // rel[tb] was last read by the epilogue of tile u-2, which arrives on bar_tempty AFTER that read
                mbar_wait(&bar_tempty[tb], ((u >> 1) & 1) ^ 1);
                if (lt < kTcNT) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        rel[tb][lt][c] = lds_f32(st + xyz_off + (lt * 3 + c) * 4) - lds_f32(st + xyz_off + (kTcNT * 3 + c) * 4);
                }
                // lane = channel quad (a 16-byte chunk of both the raw row and the operand row), warp = position mod 8:
                // every address is a per-thread constant plus a multiple of the loop counter
                const int cq = lane, pw = warp - 4;
                const uint32_t rd = st + pw * g_pitch + cq * 16;
                const uint32_t wr = static_cast<uint32_t>(cq >> 3) * (2u * kTcNT * 128u) + pw * 128u + (((cq & 7) ^ (pw & 7)) * 16u);
                if (cq < KB * 8) {
#pragma unroll 4
                    for (int i = 0; i < kTcNT / kTcXfWarps; ++i) {
                        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (cq * 4 < K) x = lds_v4(rd + i * (kTcXfWarps * g_pitch));
                        const float v[4] = {x.x, x.y, x.z, x.w};
                        float4 hi, lo;
                        tc::tf32_split4(v, hi, lo);
                        sts_v4(a_hi + wr + i * (kTcXfWarps * 128u), hi);
                        sts_v4(a_lo + wr + i * (kTcXfWarps * 128u), lo);
                    }
                }
            } else {
                // raw rows are channels: thread = (position, channel quad mod 4); 4 conflict-free scalar reads (lanes =
                // consecutive positions) -> GroupNorm + ReLU -> one 16-byte chunk of the K-major operand row; the
                // thread's position is fixed, its channel quad advances by 4 (16 channels) per step
                const int p = lt & (kTcNT - 1), c0 = (lt >> 6) * 4;
                const uint32_t rd = st + c0 * d_pitch + p * 4;
                const uint32_t ssa = smem_u32(ss_s) + c0 * 8;
                const uint32_t wr0 = p * 128u + ((((c0 >> 2)) ^ (p & 7)) * 16u), wr1 = p * 128u + ((((c0 >> 2) + 4) ^ (p & 7)) * 16u);
#pragma unroll 2
                for (int i = 0; i < KB * 2; ++i) {
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (c0 + 16 * i < K) {
                        const float4 s01 = lds_v4(ssa + i * 128), s23 = lds_v4(ssa + i * 128 + 16);
                        v[0] = fmaxf(fmaf(s01.x, lds_f32(rd + i * (16 * d_pitch)), s01.y), 0.f);
                        v[1] = fmaxf(fmaf(s01.z, lds_f32(rd + i * (16 * d_pitch) + d_pitch), s01.w), 0.f);
                        v[2] = fmaxf(fmaf(s23.x, lds_f32(rd + i * (16 * d_pitch) + 2 * d_pitch), s23.y), 0.f);
                        v[3] = fmaxf(fmaf(s23.z, lds_f32(rd + i * (16 * d_pitch) + 3 * d_pitch), s23.w), 0.f);
                    }
                    float4 hi, lo;
                    tc::tf32_split4(v, hi, lo);
                    const uint32_t off = static_cast<uint32_t>(i >> 1) * (2u * kTcNT * 128u) + ((i & 1) ? wr1 : wr0);
                    sts_v4(a_hi + off, hi);
                    sts_v4(a_lo + off, lo);
                }
            }
            tc::fence_proxy_async();
            mbar_arrive(&bar_full[ob]);
            if (GATHER && lt < kTcNT) idx_s[(u + D + 1) & 1][lt] = jn;   // read by issue(u+D+1) after the next barrier
        }
    } else if (warp == kTcMmaWarp) {
        // ================================ MMA issuer ================================
        // warp-uniform: all lanes run the loop, one elected lane issues each MMA / commit.  (Inside `if (lane == 0)`
        // every operand went through vector registers + R2UR, ~12 dependent instructions per MMA: THAT, not the tensor
        // pipe, was the ~140 cycles per instruction seen in the round-1 source-level profile.)
        {
            const uint32_t idesc = tc::make_idesc_tf32(kTcM, kTcNT, 0, 0), idesc2 = tc::make_idesc_tf32(kTcM, 2 * kTcNT, 0, 0);
            for (int u = 0; u < n_my; ++u) {
                const int ob = u % n_op, tb = u & 1;
                mbar_wait(&bar_full[ob], (u / n_op) & 1);
                mbar_wait(&bar_tempty[tb], ((u >> 1) & 1) ^ 1);
                tc::fence_after_sync();
                const uint32_t d = tmem_base + static_cast<uint32_t>(tb * kTcAccCols);
                const uint64_t bd0 = tc::make_desc_sw128(smem_u32(op_hi(ob)), 16, 1024);     // rows 0-63 hi, 64-127 lo
                for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        const uint64_t bd = bd0 + (static_cast<uint32_t>(kb) * ((2u * kTcNT * 128u) >> 4) + static_cast<uint32_t>(s4) * 2u);
                        const uint32_t wh = tmem_base + kTcWCol + static_cast<uint32_t>((kb * 4 + s4) * 8), wl = wh + Kp;
                        tc::mma_tf32_ts_elect(d, wh, bd, idesc2, (kb | s4) ? 1u : 0u);   // cols [0,64) (+)= W_hi a_hi, [64,128) (+)= W_hi a_lo
                        tc::mma_tf32_ts_elect(d, wl, bd, idesc, 1);                      // cols [0,64) += W_lo a_hi
                    }
                }
                tc::mma_commit_elect(&bar_empty[ob]);    // operand buffer may be overwritten
                tc::mma_commit_elect(&bar_tfull[tb]);    // accumulator ready
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
        for (int u = 0; u < n_my; ++u) {
            const int buf = u & 1;
            const int p0 = (blockIdx.x + u * gridDim.x) * kTcNT;
            mbar_wait(&bar_tfull[buf], (u >> 1) & 1);
            tc::fence_after_sync();
            float v[kTcNT];
            {
                // four loads in flight, one wait: [0,32) + [64,96) and [32,64) + [96,128) are the two halves' partial sums
                uint32_t r0[32], r1[32], r2[32], r3[32];
                const uint32_t ta = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(buf * kTcAccCols);
                tc::tmem_ld32_issue(ta, r0);
                tc::tmem_ld32_issue(ta + kTcNT, r1);
                tc::tmem_ld32_issue(ta + 32, r2);
                tc::tmem_ld32_issue(ta + kTcNT + 32, r3);
                tc::tmem_ld_wait();
                tc::tmem_ld_fence(r0); tc::tmem_ld_fence(r1); tc::tmem_ld_fence(r2); tc::tmem_ld_fence(r3);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
                    v[32 + j] = __uint_as_float(r2[j]) + __uint_as_float(r3[j]);
                }
            }
            if (GATHER) {
                // the three xyz input channels stay on the CUDA cores
#pragma unroll
                for (int j = 0; j < kTcNT; ++j)
                    v[j] = fmaf(wx[2], rel[buf][j][2], fmaf(wx[1], rel[buf][j][1], fmaf(wx[0], rel[buf][j][0], v[j])));
            }
            tc::fence_before_sync();
            mbar_arrive(&bar_tempty[buf]);          // TMEM buffer (and rel[buf]) free for tile u+2
            if (valid) {
                float s = 0.f, sq = 0.f;
#pragma unroll
                for (int j = 0; j < kTcNT; ++j) { s += v[j]; sq = fmaf(v[j], v[j], sq); }
                ds += static_cast<double>(s);
                dq += static_cast<double>(sq);
            }
            if (f.y) {
                // registers -> staging (row = channel, 16-byte stores, conflict-free at a 68-float pitch) -> coalesced rows
                float *srow = stage_y + tid * kTcStagePitch;
#pragma unroll
                for (int j = 0; j < kTcNT / 4; ++j)
                    *reinterpret_cast<float4 *>(srow + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                named_bar_sync(1, 128);
                const int hl = lane & 15, hr = lane >> 4;          // half-warp = one row, lane = 16-byte chunk
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int r = warp * 32 + i * 2 + hr, cr = mb * kTcM + r;
                    if (cr < Cout) {
                        const float4 t = *reinterpret_cast<const float4 *>(stage_y + r * kTcStagePitch + hl * 4);
                        *reinterpret_cast<float4 *>(f.y + (static_cast<size_t>(b) * Cout + cr) * P + p0 + hl * 4) = t;
                    }
                }
                named_bar_sync(1, 128);                            // staging free for the next tile
            }
            if (valid) {
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
    if (warp == kTcMmaWarp) tc::tmem_dealloc(tmem_base, 512);
}

// raw ring / operand buffer counts for a contraction length K (0 = does not fit)
static inline bool tc_fwd_plan(int K, bool gather, int &n_raw, int &n_op, int &stage, size_t &smem) {
    const int KB = (K + 31) / 32;
    const size_t op = static_cast<size_t>(KB) * kTcNT * 128 * 2;
    stage = gather ? kTcNT * (K * 4 + kTcRowPad) + 1024 : K * (kTcNT * 4 + kTcRowPad);
    stage = (stage + 127) & ~127;
    const size_t budget = static_cast<size_t>(kMaxSmemPerCta) - 6 * 1024 - 1024 - kTcStageBytes;
    for (n_op = 2; n_op >= 1; --n_op) {
        if (op * n_op + 2 * static_cast<size_t>(stage) > budget) continue;
        n_raw = static_cast<int>((budget - op * n_op) / stage);
        n_raw = n_raw > 4 ? 4 : n_raw;
        smem = op * n_op + static_cast<size_t>(n_raw) * stage + kTcStageBytes + 1024;
        return true;
    }
    return false;
}

}  // namespace ogc

// Tensor-core variant of ogc_sa_mlp_layer_fwd (same meaning of every argument; `w` is W (cout,cin), not W^T).
// Supported: nsample == 64, cout a multiple of 16 with cout/4 groups aligned, K = (gather ? cin-3 : cin) a
// multiple of 4 and <= 128.  Returns OGC_ERR_UNSUPPORTED otherwise (callers fall back to the SIMT kernel).
extern "C" int ogc_sa_mlp_layer_fwd_tc(int b, int n, int m, int nsample, int cin, int cout, int gather, int last,
                                       const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx,
                                       const float *y_prev, const float *ss_prev, const float *w, float *y,
                                       double *sums, float *ymax, float *ymin, unsigned char *amax,
                                       unsigned char *amin, void *stream) {
    using namespace ogc;
    if (b < 0 || m <= 0 || nsample <= 0 || cin <= 0 || cout <= 0 || !w || !sums) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    const int K = gather ? cin - 3 : cin;
    if (nsample != kTcNT || cout % 16 != 0 || cout > 256 || K < 8 || K > kTcMaxK || K % 4 != 0 || b > 65535)
        return OGC_ERR_UNSUPPORTED;
    if (last && (!ymax || !ymin || !amax || !amin)) return OGC_ERR_INVALID_ARG;
    if (gather && (!xyz || !new_xyz || !idx || !feat_pm)) return OGC_ERR_INVALID_ARG;
    if (!gather && (!y_prev || !ss_prev)) return OGC_ERR_INVALID_ARG;
    MlpTcParams q;
    q.f.Cin = cin; q.f.Cout = cout; q.f.P = m * nsample; q.f.S = nsample; q.f.M = m; q.f.N = n; q.f.Cf = cin - 3;
    q.f.xyz = xyz; q.f.new_xyz = new_xyz; q.f.feat_pm = feat_pm; q.f.idx = idx; q.f.y_prev = y_prev; q.f.ss_prev = ss_prev;
    q.f.Wt = nullptr; q.f.y = y; q.f.sums = sums; q.f.ymax = ymax; q.f.ymin = ymin; q.f.amax = amax; q.f.amin = amin;
    q.W = w; q.k_off = gather ? 3 : 0; q.K = K;
    size_t smem = 0;
    if (!tc_fwd_plan(K, gather != 0, q.n_raw, q.n_op, q.raw_stage_bytes, smem)) return OGC_ERR_UNSUPPORTED;
    if (const char *e = getenv("OGC_TC_NRAW")) { const int v = atoi(e); if (v >= 2 && v <= q.n_raw) q.n_raw = v; }
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
