// Weight gradient of a dense SharedMLP layer of the set-abstraction block, operands staged by tensor-map TMA.
//
//   dW[co][k] = sum_{b,p} dY[b,p,co] a[b,p,k]       dY = k1 dz - k2 - (y - mean) k3r   (mlp_dy.cuh)
//                                                    a  = relu(scale y_prev + shift)     (GroupNorm + ReLU of layer l-1)
//   M = C_out rows (128 per MMA, two blocks for 256), N = C_in, K = positions.
//
// Both operands are K-major with the channel as the row -- exactly the layout of the stored (B,C,P) tensors.  A TMA
// tile [channels][32 positions] with the 128-byte swizzle lands in shared memory in the canonical tcgen05 operand
// layout, so operand preparation is purely ELEMENTWISE and in place:
//   y tile  <- dY                      (the "hi" operand: the tensor core truncates an fp32 operand to TF32, probed in
//                                       scratch/tf32_round_probe.py / tests/test_gpu_tcgen05.py)
//   dz tile <- dY - trunc_tf32(dY)     (the "lo" operand, exact)
//   y_prev tile <- a, and a second tile <- a - trunc_tf32(a)
// A thread owns (channel row, 4 consecutive positions): per-channel coefficients, the last layer's arg-max slot and
// pooled gradient are per-row constants -- no shuffles, no transposes.  3 MMAs per 8 positions (hi hi, hi lo, lo hi)
// accumulate into tensor memory for the CTA's whole life; one atomic pass at the end.
// Replaces mlp_dw_tc_kernel (mlp_tc_bwd.cu) for the dense layers; utils/nn_util.py:151-168 autograd (conv weight).
#include "mlp_dy.cuh"
#include "tcgen05.cuh"
#include "tma.cuh"
#include <cstring>

namespace ogc {
namespace dwt {

constexpr int kSplitWarps = 8, kSplit = kSplitWarps * 32;
constexpr int kMmaWarp = kSplitWarps, kLoadWarp = kSplitWarps + 1;
constexpr int kThreads = (kLoadWarp + 1) * 32;
constexpr int kWin = 32;                 // positions per stage = one 128-byte swizzled row
constexpr int kMaxStages = 6;

struct Params {
    int C, Cin, P, M, S, synth;
    const float *go;                     // (B, go_ctotal, M)
    const unsigned char *sel;            // (B, C, M)
    int go_ctotal, go_coff;
    const float *coef;                   // (B, C, 4): k1, k2, k3r, mean
    const float *ss_prev;                // (B, Cin, 2)
    float *dw;                           // (C, Cin) accumulated atomically
    int stages, wins;                    // ring depth; 32-position windows per stage (1, 2 or 4: small layers amortise the
                                         // per-stage hand-shake over more bytes)
    uint32_t stage_bytes, win_bytes, off_v, off_l, off_a, off_al, off_tab;
};

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <bool SYNTH>
__global__ void __launch_bounds__(kThreads, 1)
sa_dw_tma_kernel(const __grid_constant__ Params q, const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_dz,
                 const __grid_constant__ CUtensorMap tm_yp) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kMaxStages], bar_ready[kMaxStages], bar_free[kMaxStages], bar_done;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int C = q.C, Cin = q.Cin, M = q.M;
    const int ntiles = q.P / 128;
    const int n_my = ntiles > static_cast<int>(blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int W = q.wins, nst = n_my * 4 / W, NS = q.stages;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float4 *tab_cf = reinterpret_cast<float4 *>(smem + q.off_tab);              // [C]
    float2 *tab_ss = reinterpret_cast<float2 *>(tab_cf + C);                    // [Cin]
    float2 *tab_sg = reinterpret_cast<float2 *>(tab_ss + Cin);                  // synth: [2 tiles][2 centres][C]: (slot, pooled gradient)

    if (warp == kMmaWarp) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_ready[i], kSplit); mbar_init(&bar_free[i], 1); }
        mbar_init(&bar_done, 1);
        mbar_fence_init();
    }
    for (int c = tid; c < C; c += kThreads)
        tab_cf[c] = __ldg(reinterpret_cast<const float4 *>(q.coef) + static_cast<size_t>(b) * C + c);
    for (int c = tid; c < Cin; c += kThreads)
        tab_ss[c] = __ldg(reinterpret_cast<const float2 *>(q.ss_prev) + static_cast<size_t>(b) * Cin + c);
    auto tile_of = [&](int u) { return static_cast<int>(blockIdx.x) + u * static_cast<int>(gridDim.x); };
    auto load_sg = [&](int u, int e) {           // e in [0, 2 C): centre e / C, channel e % C of the CTA's u-th tile
        const int cen = e / C, c = e - cen * C;
        const int m = 2 * tile_of(u) + cen;
        // the slot stays an integer bit pattern until it is stored into the table: a conversion right behind the load
        // would wait for it here, a whole tile before the value is needed
        return make_float2(__int_as_float(static_cast<int>(__ldg(q.sel + (static_cast<size_t>(b) * C + c) * M + m))),
                           __ldg(q.go + (static_cast<size_t>(b) * q.go_ctotal + q.go_coff + c) * M + m));
    };
    auto sg_entry = [](float2 raw) { return make_float2(static_cast<float>(__float_as_int(raw.x)), raw.y); };
    if (SYNTH && n_my > 0)
        for (int e = tid; e < 2 * C; e += kThreads) tab_sg[e] = sg_entry(load_sg(0, e));
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == kLoadWarp) {
        // ============================================ TMA loader (warp-uniform loop, lane 0 issues) ============================================
        if (lane == 0) {
            tma::prefetch_map(&tm_y);
            tma::prefetch_map(&tm_yp);
            if (!SYNTH) tma::prefetch_map(&tm_dz);
        }
        const uint32_t bytes = static_cast<uint32_t>((SYNTH ? 1 : 2) * C + Cin) * 128u * static_cast<uint32_t>(W);
        for (int s = 0; s < nst; ++s) {
            const int st = s % NS;
            mbar_wait(&bar_free[st], ((s / NS) & 1) ^ 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&bar_full[st], bytes);
                for (int j = 0; j < W; ++j) {
                    uint8_t *base = smem + static_cast<size_t>(st) * q.stage_bytes + static_cast<size_t>(j) * q.win_bytes;
                    const int wq = s * W + j;
                    const int p0 = tile_of(wq >> 2) * 128 + (wq & 3) * kWin;
                    tma::load_2d(base + q.off_v, &tm_y, p0, b * C, &bar_full[st]);
                    if (!SYNTH) tma::load_2d(base + q.off_l, &tm_dz, p0, b * C, &bar_full[st]);
                    tma::load_2d(base + q.off_a, &tm_yp, p0, b * Cin, &bar_full[st]);
                }
            }
            __syncwarp();
        }
    } else if (warp == kMmaWarp) {
        // ============================================ MMA issuer ============================================
        // two MMAs per 8 positions instead of three: dY x [a ; a_lo] (the two tiles are adjacent: one B operand of 2 C_in rows)
        // puts hi*hi and hi*lo into adjacent column ranges, dY_lo x a accumulates into the first; the epilogue adds the ranges
        const uint32_t idesc1 = tc::make_idesc_tf32(128, 2 * Cin, 0, 0), idesc2 = tc::make_idesc_tf32(128, Cin, 0, 0);
        const int mblocks = C > 128 ? 2 : 1;
        for (int s = 0; s < nst; ++s) {
            const int st = s % NS;
            mbar_wait(&bar_ready[st], (s / NS) & 1);
            tc::fence_after_sync();
            for (int j = 0; j < W; ++j) {
                const uint32_t base = smem_u32(smem + static_cast<size_t>(st) * q.stage_bytes + static_cast<size_t>(j) * q.win_bytes);
                const uint64_t dv = tc::make_desc_sw128(base + q.off_v, 16, 1024), dl = tc::make_desc_sw128(base + q.off_l, 16, 1024);
                    const uint64_t da = tc::make_desc_sw128(base + q.off_a, 16, 1024);          // a tile, a_lo tile right behind it
                for (int mb = 0; mb < mblocks; ++mb) {
                    const uint32_t d = tmem_base + static_cast<uint32_t>(mb * 2 * Cin);
                    const uint64_t moff = static_cast<uint64_t>(mb) * ((128u * 128u) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t acc = (s | j | k) ? 1u : 0u;
                        tc::mma_tf32_ss_elect(d, dv + moff + 2u * k, da + 2u * k, idesc1, acc);
                        tc::mma_tf32_ss_elect(d, dl + moff + 2u * k, da + 2u * k, idesc2, 1u);
                    }
                }
            }
            tc::mma_commit_elect(&bar_free[st]);
        }
        tc::mma_commit_elect(&bar_done);
    } else {
        // ============================================ splitters: elementwise operand preparation, in place ============================================
        const int q4 = tid & 7, r0 = tid >> 3;          // float4 chunk of the 128-byte row, first row (rows r0 + 32 i)
        float2 sg_next[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        for (int s = 0; s < nst; ++s) {
            const int st = s % NS, u = (s * W) >> 2;
            const bool tile_first = ((s * W) & 3) == 0, tile_last = ((s * W + W - 1) & 3) == 3;
            if (SYNTH && tile_first && u + 1 < n_my) {  // next tile's (slot, gradient) pairs: requested now, stored at the tile's end
                if (tid < 2 * C) sg_next[0] = load_sg(u + 1, tid);
                if (tid + kSplit < 2 * C) sg_next[1] = load_sg(u + 1, tid + kSplit);
            }
            mbar_wait(&bar_full[st], (s / NS) & 1);
            for (int j = 0; j < W; ++j) {
                uint8_t *base = smem + static_cast<size_t>(st) * q.stage_bytes + static_cast<size_t>(j) * q.win_bytes;
                const int g = (s * W + j) & 3;
                const float2 *sg = tab_sg + ((u & 1) * 2 + (g >> 1)) * C;       // this window's centre
                const int s0 = (g & 1) * kWin + 4 * q4;                          // slot of this thread's first position
                for (int r = r0; r < C; r += 32) {
                    const uint32_t off = static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(q4 ^ (r & 7)) << 4);
                    float4 *pv = reinterpret_cast<float4 *>(base + q.off_v + off), *pl = reinterpret_cast<float4 *>(base + q.off_l + off);
                    const float4 y4 = *pv;
                    const float4 cf = tab_cf[r];
                    float z[4];
                    if (SYNTH) {
                        const float2 e = sg[r];
                        const int sl = static_cast<int>(e.x) - s0;
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
                for (int r = r0; r < Cin; r += 32) {
                    const uint32_t off = static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(q4 ^ (r & 7)) << 4);
                    float4 *pa = reinterpret_cast<float4 *>(base + q.off_a + off), *pal = reinterpret_cast<float4 *>(base + q.off_al + off);
                    const float4 y4 = *pa;
                    const float2 ss = tab_ss[r];
                    const float a0 = fmaxf(fmaf(ss.x, y4.x, ss.y), 0.f), a1 = fmaxf(fmaf(ss.x, y4.y, ss.y), 0.f);
                    const float a2 = fmaxf(fmaf(ss.x, y4.z, ss.y), 0.f), a3 = fmaxf(fmaf(ss.x, y4.w, ss.y), 0.f);
                    *pa = make_float4(a0, a1, a2, a3);
                    *pal = make_float4(a0 - trunc_tf32(a0), a1 - trunc_tf32(a1), a2 - trunc_tf32(a2), a3 - trunc_tf32(a3));
                }
            }
            tc::fence_proxy_async();                    // generic writes -> the tensor core's shared-memory reads
            mbar_arrive(&bar_ready[st]);
            if (SYNTH && tile_last && u + 1 < n_my) {   // every splitter has finished this tile's windows once it passes the barrier
                asm volatile("bar.sync 1, %0;" ::"r"(kSplit) : "memory");
                float2 *dst = tab_sg + (((u + 1) & 1) * 2) * C;
                if (tid < 2 * C) dst[tid] = sg_entry(sg_next[0]);
                if (tid + kSplit < 2 * C) dst[tid + kSplit] = sg_entry(sg_next[1]);
                asm volatile("bar.sync 1, %0;" ::"r"(kSplit) : "memory");
            }
        }
        // ---- epilogue: warps 0-3 own the 128 accumulator lanes ----
        if (warp < 4 && n_my > 0) {
            mbar_wait(&bar_done, 0);
            tc::fence_after_sync();
            const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
            for (int mb = 0; mb < (C > 128 ? 2 : 1); ++mb) {
                const int co = mb * 128 + warp * 32 + lane;
                for (int c0 = 0; c0 < Cin; c0 += 32) {
                    float v[32], w[32];
                    tc::tmem_ld32(trow + static_cast<uint32_t>(mb * 2 * Cin + c0), v);             // hi*hi + lo*hi
                    tc::tmem_ld32(trow + static_cast<uint32_t>(mb * 2 * Cin + Cin + c0), w);       // hi*lo
                    if (co < C) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(q.dw + static_cast<size_t>(co) * Cin + c0 + j, v[j] + w[j]);
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace dwt
}  // namespace ogc

// Same arguments and result as ogc_sa_mlp_narrow_dw / the dense mode of ogc_sa_mlp_layer_dw_tc (dw accumulated into a
// caller-zeroed (cout, cin) buffer).  nsample == 64, m even, cout in {32, 64, 128, 256}, cin in {32, 64, 96, 128};
// OGC_ERR_UNSUPPORTED otherwise.
extern "C" int ogc_sa_dw_tma(int b, int m, int nsample, int cout, int cin, const float *dz, const float *go, int go_ctotal,
                             int go_coff, const unsigned char *sel, const float *y, const float *coef, const float *y_prev,
                             const float *ss_prev, float *dw, void *stream) {
    using namespace ogc;
    using namespace ogc::dwt;
    if (b < 0 || m <= 0 || cout <= 0 || cin <= 0 || !y || !coef || !y_prev || !ss_prev || !dw) return OGC_ERR_INVALID_ARG;
    if (!dz && (!go || !sel)) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (nsample != 64 || (m & 1) || b > 65535 || cout % 32 != 0 || cout > 256 || cin % 32 != 0 || cin > 128) return OGC_ERR_UNSUPPORTED;
    const bool synth = dz == nullptr;
    Params q{};
    q.C = cout; q.Cin = cin; q.P = m * nsample; q.M = m; q.S = nsample; q.synth = synth;
    q.go = go; q.sel = sel; q.go_ctotal = go_ctotal; q.go_coff = go_coff; q.coef = coef; q.ss_prev = ss_prev; q.dw = dw;
    // a stage: dY tile | lo tile | a tile | a-lo tile, each [rows][128 B], 1024-byte aligned; the 128-row MMA may read past a
    // narrow dY tile into the tiles behind it (unused accumulator lanes)
    const uint32_t vt = static_cast<uint32_t>(cout) * 128u, at = static_cast<uint32_t>(cin) * 128u;
    q.off_v = 0; q.off_l = vt; q.off_a = 2 * vt; q.off_al = 2 * vt + at;
    q.win_bytes = 2 * vt + 2 * at;
    if (q.win_bytes < 128u * 128u + vt) q.win_bytes = 128u * 128u + vt;          // lo tile + 128 rows readable
    q.win_bytes = (q.win_bytes + 1023u) & ~1023u;
    q.wins = q.win_bytes <= 16u * 1024u ? 4 : (q.win_bytes <= 32u * 1024u ? 2 : 1);
    q.stage_bytes = q.win_bytes * static_cast<uint32_t>(q.wins);
    const uint32_t tab_bytes = static_cast<uint32_t>(cout) * 16u + static_cast<uint32_t>(cin) * 8u + (synth ? 4u * cout * 8u : 0u);
    const long long budget = static_cast<long long>(kMaxSmemPerCta) - 2048 - tab_bytes;
    int stages = static_cast<int>(budget / q.stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return OGC_ERR_UNSUPPORTED;
    q.stages = stages;
    q.off_tab = static_cast<uint32_t>(stages) * q.stage_bytes;
    const size_t smem = static_cast<size_t>(q.off_tab) + tab_bytes + 1024;
    CUtensorMap tm_y, tm_dz, tm_yp;
    memset(&tm_dz, 0, sizeof(tm_dz));
    const uint64_t p64 = static_cast<uint64_t>(q.P);
    bool ok = tma::make_2d_f32(&tm_y, y, p64, static_cast<uint64_t>(b) * cout, kWin, cout, 1);
    if (ok && !synth) ok = tma::make_2d_f32(&tm_dz, dz, p64, static_cast<uint64_t>(b) * cout, kWin, cout, 1);
    if (ok) ok = tma::make_2d_f32(&tm_yp, y_prev, p64, static_cast<uint64_t>(b) * cin, kWin, cin, 1);
    if (!ok) return OGC_ERR_UNSUPPORTED;
    int per_sample = kNumSMs / b;
    const int ntiles = q.P / 128;
    per_sample = per_sample > ntiles ? ntiles : (per_sample < 1 ? 1 : per_sample);
    dim3 grid(per_sample, b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (synth) {
        e = cudaFuncSetAttribute(sa_dw_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        sa_dw_tma_kernel<true><<<grid, kThreads, smem, st>>>(q, tm_y, tm_dz, tm_yp);
    } else {
        e = cudaFuncSetAttribute(sa_dw_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        sa_dw_tma_kernel<false><<<grid, kThreads, smem, st>>>(q, tm_y, tm_dz, tm_yp);
    }
    OGC_RETURN_LAUNCH_STATUS();
}
