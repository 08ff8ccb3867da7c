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

namespace ogc {
namespace chain {

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
    uint32_t off_w, off_tab, off_scr;
    uint32_t col_acc[2];
    int ring;                      // A-operand ring slots (32-column chunks, hi + lo = 64 tensor-memory columns each)
};

template <bool SYNTH, bool SCATTER>
__global__ void __launch_bounds__(kThreads, 1)
sa_dx_kernel(DxParams q) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_acc[2], bar_accfree[2], bar_kfull[8], bar_kfree[8];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int C = q.dy.C, P = q.dy.P, M = q.dy.M, rows = q.rows;
    const int ntiles = M / 2;
    const int n_my = ntiles > static_cast<int>(blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nchunks = C >> 5, ring = q.ring;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float4 *tab_cf = reinterpret_cast<float4 *>(smem + q.off_tab);              // [C]: k1, k2, k3r, mean of layer l
    float2 *tab_ss = reinterpret_cast<float2 *>(tab_cf + kMaxC);                // [rows]: scale, shift of layer l-1
    float *csum = reinterpret_cast<float *>(tab_ss + 128);                      // [rows][2] per-CTA channel sums
    float *scr = reinterpret_cast<float *>(smem + q.off_scr);                   // [8 epilogue warps][32][33]

    if (warp == kMmaWarp) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc[i], 1); mbar_init(&bar_accfree[i], kEpi); }
        for (int c = 0; c < 8; ++c) { mbar_init(&bar_kfull[c], 128); mbar_init(&bar_kfree[c], 1); }
        mbar_fence_init();
    }
    // resident B operand: rows k < rows, K index = co: element (k, co) = W[co][row_off + k]
    {
        const int kp = align_up(C, 32);
        const uint32_t blk = 2u * static_cast<uint32_t>(rows) * 128u;
        for (int e = tid; e < rows * kp; e += kThreads) {
            const int co = e / rows, k = e - co * rows;                        // consecutive threads: consecutive W columns
            const float v = co < C ? __ldg(q.W + static_cast<size_t>(co) * q.cin_full + q.row_off + k) : 0.f;
            const float hi = tc::tf32_hi(v), lo = tc::tf32_hi(v - hi);
            const uint32_t off = static_cast<uint32_t>(co >> 5) * blk + tc::sw128_offset(k, co & 31);
            *reinterpret_cast<float *>(smem + q.off_w + off) = hi;
            *reinterpret_cast<float *>(smem + q.off_w + off + static_cast<uint32_t>(rows) * 128u) = lo;
        }
    }
    for (int c = tid; c < C; c += kThreads)
        tab_cf[c] = __ldg(reinterpret_cast<const float4 *>(q.dy.coef) + static_cast<size_t>(b) * C + c);
    if (!SCATTER) {
        for (int c = tid; c < rows; c += kThreads) {
            tab_ss[c] = __ldg(reinterpret_cast<const float2 *>(q.ss_prev) + static_cast<size_t>(b) * rows + c);
            csum[2 * c] = csum[2 * c + 1] = 0.f;
        }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t colA = 0;                                   // ring slot s: hi at 64 s, lo at 64 s + 32
    auto tile_of = [&](int u) { return static_cast<int>(blockIdx.x) + u * static_cast<int>(gridDim.x); };

    if (warp >= kProdWarp0 && warp < kMmaWarp) {
        // ============================================ producer: dY chunks ============================================
        const int pw = warp & 3, pg = (warp - kProdWarp0) >> 2;
        const int pt = pw * 32 + lane;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(pw * 32) << 16) + colA;
        for (int u = 0; u < n_my; ++u) {
            const int t = tile_of(u);
            const size_t pos = static_cast<size_t>(t) * kTile + pt;
            const int m = t * 2 + (pt >> 6), s_own = pt & 63;
            const float *yb = q.dy.y + static_cast<size_t>(b) * C * P + pos;
            const float *zb = SYNTH ? nullptr : q.dy.dz + static_cast<size_t>(b) * C * P + pos;
            for (int c = pg; c < nchunks; c += 2) {
                const int chunk_seq = u * nchunks + c;             // position in the CTA's chunk sequence
                const int slot = chunk_seq % ring;
                const uint32_t use = static_cast<uint32_t>(chunk_seq / ring);      // how many times the slot was used before
                // loads of the whole chunk first (64 independent requests in flight), then the slot hand-shake
                float yv[32], zv[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) yv[j] = __ldg(yb + static_cast<size_t>(32 * c + j) * P);
                if (SYNTH) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const size_t o = (static_cast<size_t>(b) * C + 32 * c + j) * M + m;
                        const int sl = __ldg(q.dy.sel + o);
                        zv[j] = sl == s_own ? __ldg(q.dy.go + (static_cast<size_t>(b) * q.dy.go_ctotal + q.dy.go_coff + 32 * c + j) * M + m) : 0.f;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) zv[j] = __ldg(zb + static_cast<size_t>(32 * c + j) * P);
                }
                mbar_wait(&bar_kfree[slot], (use & 1) ^ 1);
                tc::fence_after_sync();
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    float hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 cf = tab_cf[32 * c + 8 * h + j];
                        const float v = fmaf(cf.x, zv[8 * h + j], -cf.y) - (yv[8 * h + j] - cf.w) * cf.z;
                        tc::tf32_split(v, hi[j], lo[j]);
                    }
                    tc::tmem_st8_nowait(trow + slot * 64 + 8 * h, hi);
                    tc::tmem_st8_nowait(trow + slot * 64 + 32 + 8 * h, lo);
                }
                tc::tmem_st_wait();
                tc::fence_before_sync();
                mbar_arrive(&bar_kfull[slot]);
            }
        }
    } else if (warp == kMmaWarp) {
        // ============================================ MMA issuer (warp-uniform) ============================================
        const uint32_t idesc = tc::make_idesc_tf32(kTile, rows, 0, 0);
        const uint32_t blk16 = (2u * static_cast<uint32_t>(rows) * 128u) >> 4, lo16 = (static_cast<uint32_t>(rows) * 128u) >> 4;
        const uint64_t d0 = tc::make_desc_sw128(smem_u32(smem + q.off_w), 16, 1024);
        int chunk_seq = 0;
        for (int u = 0; u < n_my; ++u) {
            const int buf = u & 1;
            mbar_wait(&bar_accfree[buf], ((u >> 1) & 1) ^ 1);
            const uint32_t d = tmem_base + q.col_acc[buf];
            for (int c = 0; c < nchunks; ++c, ++chunk_seq) {
                const int slot = chunk_seq % ring;
                const uint32_t use = static_cast<uint32_t>(chunk_seq / ring);
                mbar_wait(&bar_kfull[slot], use & 1);
                tc::fence_after_sync();
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const uint64_t bh = d0 + (static_cast<uint32_t>(c) * blk16 + static_cast<uint32_t>(s) * 2u), bl = bh + lo16;
                    const uint32_t ah = tmem_base + colA + static_cast<uint32_t>(slot * 64 + s * 8), al = ah + 32u;
                    tc::mma_tf32_ts_elect(d, ah, bh, idesc, (c | s) ? 1u : 0u);
                    tc::mma_tf32_ts_elect(d, ah, bl, idesc, 1u);
                    tc::mma_tf32_ts_elect(d, al, bh, idesc, 1u);
                }
                tc::mma_commit_elect(&bar_kfree[slot]);
            }
            tc::mma_commit_elect(&bar_acc[buf]);
        }
    } else {
        // ============================================ epilogue ============================================
        const int eq = warp & 3, eg = warp >> 2;
        const int et = eq * 32 + lane;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(eq * 32) << 16);
        float *sc = scr + warp * (32 * 33);
        const int gsz = rows / kGnGroups;
        float mu[kGnGroups] = {0.f, 0.f, 0.f, 0.f}, rs[kGnGroups] = {0.f, 0.f, 0.f, 0.f};
        if (!SCATTER) {
#pragma unroll
            for (int g = 0; g < kGnGroups; ++g) {
                mu[g] = __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2);
                rs[g] = __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2 + 1);
            }
        }
        const int nchunk_mine = rows > eg * 32 ? (rows - eg * 32 + 63) / 64 : 0;
        for (int u = 0; u < n_my; ++u) {
            const int t = tile_of(u), buf = u & 1;
            const size_t pos = static_cast<size_t>(t) * kTile + et;
            if (SCATTER) {
                const int jpt = __ldg(q.idx + static_cast<size_t>(b) * P + pos);
                mbar_wait(&bar_acc[buf], (u >> 1) & 1);
                tc::fence_after_sync();
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
                continue;
            }
            // dense: 16-column pieces; the ReLU-mask inputs of the NEXT piece are requested before this one is processed
            const int npieces = 2 * nchunk_mine;
            auto piece_col = [&](int i) { return eg * 32 + (i >> 1) * 64 + (i & 1) * 16; };
            const float *ypb = q.y_prev + static_cast<size_t>(b) * rows * P + pos;
            float ypn[16];
            if (npieces > 0) {
                const int c0 = piece_col(0);
#pragma unroll
                for (int j = 0; j < 16; ++j) ypn[j] = __ldg(ypb + static_cast<size_t>(c0 + j) * P);
            }
            mbar_wait(&bar_acc[buf], (u >> 1) & 1);
            tc::fence_after_sync();
            for (int i = 0; i < npieces; ++i) {
                const int c0 = piece_col(i);
                float yp[16], v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) yp[j] = ypn[j];
                if (i + 1 < npieces) {
                    const int c1 = piece_col(i + 1);
#pragma unroll
                    for (int j = 0; j < 16; ++j) ypn[j] = __ldg(ypb + static_cast<size_t>(c1 + j) * P);
                }
                tc::tmem_ld16(trow + q.col_acc[buf] + c0, v);
                if (i == npieces - 1) {
                    tc::fence_before_sync();
                    mbar_arrive(&bar_accfree[buf]);
                }
                const int g = c0 / gsz;                    // a 16-channel piece straddles groups only when rows < 64
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 s2 = tab_ss[c0 + j];
                    const int gj = gsz >= 16 ? g : (c0 + j) / gsz;
                    float m_ = mu[0], r_ = rs[0];
#pragma unroll
                    for (int gg = 1; gg < kGnGroups; ++gg) { m_ = gj == gg ? mu[gg] : m_; r_ = gj == gg ? rs[gg] : r_; }
                    const float gz = fmaf(s2.x, yp[j], s2.y) > 0.f ? v[j] : 0.f;
                    v[j] = gz;
                    yp[j] = gz * ((yp[j] - m_) * r_);          // dz_prev * xhat_prev
                }
                {
                    float *dzp = q.dz_prev + (static_cast<size_t>(b) * rows + c0) * P + pos;
                    float *yo[4] = {dzp, dzp + P, dzp + 2 * static_cast<size_t>(P), dzp + 3 * static_cast<size_t>(P)};
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        *yo[j & 3] = v[j];
                        if (j < 12) yo[j & 3] += static_cast<size_t>(4) * P;
                    }
                }
                // per-channel sums over this warp's 32 positions: transpose through shared memory; lane = (channel, half)
                const int ch = lane & 15, hf = lane >> 4;
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) sc[j * 33 + lane] = v[j];
                __syncwarp();
#pragma unroll
                for (int p = 0; p < 16; ++p) s0 += sc[ch * 33 + hf * 16 + p];
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 16; ++j) sc[j * 33 + lane] = yp[j];
                __syncwarp();
#pragma unroll
                for (int p = 0; p < 16; ++p) s1 += sc[ch * 33 + hf * 16 + p];
                __syncwarp();
                s0 += __shfl_xor_sync(OGC_FULL_MASK, s0, 16);
                s1 += __shfl_xor_sync(OGC_FULL_MASK, s1, 16);
                if (hf == 0) {
                    atomicAdd(&csum[2 * (c0 + ch)], s0);
                    atomicAdd(&csum[2 * (c0 + ch) + 1], s1);
                }
            }
            if (npieces == 0) {
                tc::fence_before_sync();
                mbar_arrive(&bar_accfree[buf]);
            }
        }
        if (!SCATTER) {
            named_bar_sync(kEpiBar, kEpi);
            for (int c = tid; c < rows * 2; c += kEpi)
                if (n_my > 0) atomicAdd(q.chan_sums + static_cast<size_t>(b) * rows * 2 + c, csum[c]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, 512);
}

// chan_sums (B,rows,2) -> dgamma / dbeta (+=), ab (B,4,2) (+=): the inputs of ogc_gn_bwd_coef for layer l-1
__global__ void dx_finalize_kernel(int B, int rows, const float *__restrict__ chan_sums, const float *__restrict__ gamma,
                                   double *__restrict__ ab, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= rows) return;
    const int g = c / (rows / kGnGroups);
    const double gm = static_cast<double>(gamma[c]);
    float sb = 0.f, sg = 0.f;
    for (int b = 0; b < B; ++b) {
        const float s0 = chan_sums[(static_cast<size_t>(b) * rows + c) * 2], s1 = chan_sums[(static_cast<size_t>(b) * rows + c) * 2 + 1];
        sb += s0; sg += s1;
        atomicAdd(ab + (b * kGnGroups + g) * 2, gm * s0);
        atomicAdd(ab + (b * kGnGroups + g) * 2 + 1, gm * s1);
    }
    atomicAdd(dbeta + c, sb);
    atomicAdd(dgamma + c, sg);
}

}  // namespace chain
}  // namespace ogc

// Drop-in replacement of ogc_sa_mlp_layer_dx_tc (same arguments, same outputs) in the positions-on-M orientation, plus
// `chan_sums`: a caller-zeroed (b, rows, 2) fp32 workspace (dense mode).  nsample == 64, m even, cout a multiple of 32
// (<= 256), rows a multiple of 16 (<= 128; dense: a multiple of 32).  OGC_ERR_UNSUPPORTED otherwise.
extern "C" int ogc_sa_chain_dx(int b, int n, int m, int nsample, int cout, int cin_full, int row_off, int rows,
                               const float *dz, const float *go, int go_ctotal, int go_coff, const unsigned char *sel,
                               const float *y, const float *coef, const float *w, const float *y_prev,
                               const float *ss_prev, const float *mean_rstd_prev, const float *gamma_prev,
                               float *dz_prev, double *ab_prev, float *dgamma_prev, float *dbeta_prev, const int *idx,
                               float *dfeat_pm, int dfeat_stride, int dfeat_off, float *chan_sums, void *stream) {
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
    // shared memory: W^T tile, tables, per-warp transpose scratch
    const uint32_t w_bytes = w_tile_bytes(rows, cout);
    q.off_w = 0;
    q.off_tab = w_bytes;
    const uint32_t tab_bytes = kMaxC * 16u + 128u * 8u + 128u * 2u * 4u;
    q.off_scr = (q.off_tab + tab_bytes + 127u) & ~127u;
    const size_t smem = static_cast<size_t>(q.off_scr) + kEpiWarps * 32u * 33u * 4u + 1024;
    if (smem > static_cast<size_t>(kMaxSmemPerCta) - 2048) return OGC_ERR_UNSUPPORTED;
    // tensor memory: A ring (64 columns per 32-channel chunk) + two accumulators of `rows` columns
    const int acc = align_up(rows, 32);
    int ring = (512 - 2 * acc) / 64;
    const int nchunks = cout / 32;
    ring = ring > nchunks ? nchunks : ring;
    if (ring < 2) return OGC_ERR_UNSUPPORTED;
    if (ring > 8) ring = 8;
    q.ring = ring;
    q.col_acc[0] = static_cast<uint32_t>(ring * 64);
    q.col_acc[1] = q.col_acc[0] + static_cast<uint32_t>(acc);
    int per_sample = kNumSMs / b;
    per_sample = per_sample > m / 2 ? m / 2 : (per_sample < 1 ? 1 : per_sample);
    dim3 grid(per_sample, b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool synth = dz == nullptr;
#define OGC_DX_LAUNCH(S, SC)                                                                                          \
    do {                                                                                                              \
        cudaError_t e = cudaFuncSetAttribute(sa_dx_kernel<S, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                             static_cast<int>(smem));                                                 \
        if (e != cudaSuccess) return static_cast<int>(e);                                                             \
        sa_dx_kernel<S, SC><<<grid, kThreads, smem, st>>>(q);                                                         \
    } while (0)
    if (synth && scatter) OGC_DX_LAUNCH(true, true);
    else if (synth) OGC_DX_LAUNCH(true, false);
    else if (scatter) OGC_DX_LAUNCH(false, true);
    else OGC_DX_LAUNCH(false, false);
#undef OGC_DX_LAUNCH
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    if (!scatter) {
        dx_finalize_kernel<<<(rows + 127) / 128, 128, 0, st>>>(b, rows, chan_sums, gamma_prev, ab_prev, dgamma_prev, dbeta_prev);
    }
    OGC_RETURN_LAUNCH_STATUS();
}
