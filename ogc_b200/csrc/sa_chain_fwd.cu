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
//   producer warps 4-7: cp.async the tile's raw input into a shared-memory stage one tile ahead (gathered
//        feature rows, or rows of the stored previous layer), then thread = position turns its row into the layer-1
//        A operand in TENSOR MEMORY (GroupNorm+ReLU of the stored layer, hi/lo split)
//   MMA warp 8: one thread issues the layer's k-steps (A from tensor memory, B = resident weights), commits to mbarriers
//   epilogue warps 0-3: thread = position reads its accumulator row, adds the three relative-coordinate input
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
};

__global__ void __launch_bounds__(kThreads, 1)
sa_chain_fwd_kernel(FwdParams q) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_a[3], bar_acc[3], bar_a1free, bar_accfree;
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
    float4 *tab_wx = reinterpret_cast<float4 *>(tab_ss1 + kMaxC);            // [C0] xyz columns of W1

    // tensor-memory columns
    const uint32_t colA1 = 0, colA23 = static_cast<uint32_t>(align_up(2 * K1, 32));
    const uint32_t a23 = nl > 1 ? 2u * static_cast<uint32_t>(max(q.C[0], nl > 2 ? q.C[1] : 0)) : 0u;
    const uint32_t colACC = colA23 + a23;

    if (warp == kMmaWarp) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int l = 0; l < 3; ++l) { mbar_init(&bar_a[l], 128); mbar_init(&bar_acc[l], 1); }
        mbar_init(&bar_a1free, 1);
        mbar_init(&bar_accfree, 128);
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
        for (int c = tid; c < q.C[0]; c += kThreads)
            tab_wx[c] = make_float4(__ldg(W0 + static_cast<size_t>(c) * q.ldw[0]), __ldg(W0 + static_cast<size_t>(c) * q.ldw[0] + 1),
                                    __ldg(W0 + static_cast<size_t>(c) * q.ldw[0] + 2), 0.f);
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

    if (warp >= 4 && warp < kMmaWarp) {
        // ============================================ producer ============================================
        const int pt = tid - 128, pw = warp - 4;
        const uint32_t raw = smem_u32(smem + q.off_raw), pitch = q.raw_pitch;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(pw * 32) << 16) + colA1;
        const int Cf = q.Cf;
        auto tile_p0 = [&](int u) { return (static_cast<int>(blockIdx.x) + u * static_cast<int>(gridDim.x)) * kTile; };
        auto load_j = [&](int u) { return (gather && u < n_my) ? __ldg(q.idx + static_cast<size_t>(b) * P + tile_p0(u) + pt) : 0; };
        auto issue = [&](int u, int j) {
            if (gather) {
                if (q.small) return;
                for (int i = 0; i < 32; ++i) {
                    const int ji = __shfl_sync(OGC_FULL_MASK, j, i);
                    const float *src = q.feat_pm + (static_cast<size_t>(b) * q.N + ji) * Cf;
                    const uint32_t dst = raw + static_cast<uint32_t>(pw * 32 + i) * pitch;
                    for (int ch = lane; ch < (Cf >> 2); ch += 32) cp_async16_s(dst + ch * 16, src + ch * 4);
                }
            } else {
                const float *src = q.y_in + static_cast<size_t>(b) * K1 * P + tile_p0(u) + (pt & 31) * 4;
                for (int c = pt >> 5; c < K1; c += 4)
                    cp_async16_s(raw + static_cast<uint32_t>(c) * pitch + (pt & 31) * 16, src + static_cast<size_t>(c) * P);
            }
        };
        int j_cur = load_j(0);
        if (n_my > 0) issue(0, j_cur);
        cp_async_commit();
        for (int u = 0; u < n_my; ++u) {
            const int j_next = load_j(u + 1);
            float f8[8];
            if (gather && q.small) {
#pragma unroll
                for (int c = 0; c < 8; ++c) f8[c] = c < Cf ? __ldg(q.feat_pm + (static_cast<size_t>(b) * q.N + j_cur) * Cf + c) : 0.f;
            }
            cp_async_wait(0);
            named_bar_sync(kProdBar, 128);                 // the tile's raw rows have landed for every producer thread
            mbar_wait(&bar_a1free, (u & 1) ^ 1);           // layer-1 MMAs of the previous tile have read the operand
            tc::fence_after_sync();
            if (gather && q.small) {
                float hi[8], lo[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) tc::tf32_split(f8[c], hi[c], lo[c]);
                tc::tmem_st8_nowait(trow, hi);
                tc::tmem_st8_nowait(trow + 8, lo);
            } else {
                for (int k0 = 0; k0 < K1; k0 += 32) {
                    if (k0 + 32 <= K1) {
                        float hi[32], lo[32];
                        if (gather) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 x = lds_v4(raw + static_cast<uint32_t>(pt) * pitch + (k0 + 4 * i) * 4);
                                tc::tf32_split(x.x, hi[4 * i], lo[4 * i]);
                                tc::tf32_split(x.y, hi[4 * i + 1], lo[4 * i + 1]);
                                tc::tf32_split(x.z, hi[4 * i + 2], lo[4 * i + 2]);
                                tc::tf32_split(x.w, hi[4 * i + 3], lo[4 * i + 3]);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const float2 s2 = tab_in[k0 + i];
                                const float y = lds_f32(raw + static_cast<uint32_t>(k0 + i) * pitch + pt * 4);
                                tc::tf32_split(fmaxf(fmaf(s2.x, y, s2.y), 0.f), hi[i], lo[i]);
                            }
                        }
                        tc::tmem_st32_nowait(trow + k0, hi);
                        tc::tmem_st32_nowait(trow + K1 + k0, lo);
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
                }
            }
            tc::tmem_st_wait();
            tc::fence_before_sync();
            mbar_arrive(&bar_a[0]);
            named_bar_sync(kProdBar, 128);                 // every producer thread is done with the raw stage
            if (u + 1 < n_my) issue(u + 1, j_next);
            cp_async_commit();
            j_cur = j_next;
        }
    } else if (warp == kMmaWarp) {
        // ============================================ MMA issuer ============================================
        if (lane == 0) {
            const uint32_t d = tmem_base + colACC;
            for (int u = 0; u < n_my; ++u) {
                for (int l = 0; l < nl; ++l) {
                    mbar_wait(&bar_a[l], u & 1);
                    if (l == 0) mbar_wait(&bar_accfree, (u & 1) ^ 1);
                    tc::fence_after_sync();
                    issue_layer(d, tmem_base + (l == 0 ? colA1 : colA23), l == 0 ? K1 : q.C[l - 1],
                                smem_u32(smem + q.off_w[l]), q.C[l]);
                    tc::mma_commit(&bar_acc[l]);
                    if (l == 0) tc::mma_commit(&bar_a1free);
                }
            }
        }
    } else {
        // ============================================ epilogue ============================================
        const int et = tid, ew = warp;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
        const int gsz = q.c_total / kGnGroups;
        double dsum[kGnGroups] = {0.0, 0.0, 0.0, 0.0}, dsq[kGnGroups] = {0.0, 0.0, 0.0, 0.0};
        for (int u = 0; u < n_my; ++u) {
            const int t = static_cast<int>(blockIdx.x) + u * static_cast<int>(gridDim.x);
            const int p0 = t * kTile;
            float rx = 0.f, ry = 0.f, rz = 0.f;
            if (gather) {
                // relative coordinates of this thread's position (3 input channels kept on the CUDA cores)
                const int j = __ldg(q.idx + static_cast<size_t>(b) * P + p0 + et);
                const float *pj = q.xyz + (static_cast<size_t>(b) * q.N + j) * 3;
                const float *pc = q.new_xyz + (static_cast<size_t>(b) * q.M + t * 2 + (et >> 6)) * 3;
                rx = __ldg(pj) - __ldg(pc); ry = __ldg(pj + 1) - __ldg(pc + 1); rz = __ldg(pj + 2) - __ldg(pc + 2);
            }
            float fsum[kGnGroups] = {0.f, 0.f, 0.f, 0.f}, fsq[kGnGroups] = {0.f, 0.f, 0.f, 0.f};
            for (int l = 0; l < nl; ++l) {
                const bool lastl = l == nl - 1;
                const int n = q.C[l];
                mbar_wait(&bar_acc[l], u & 1);
                tc::fence_after_sync();
                const float2 *tss = l == 0 ? tab_ss0 : tab_ss1;
                for (int c0 = 0; c0 < n; c0 += 32) {
                    float v[32];
                    tc::tmem_ld32(trow + colACC + c0, v);
                    if (lastl && c0 + 32 >= n) {            // accumulator drained: the next tile's layer 1 may start
                        tc::fence_before_sync();
                        mbar_arrive(&bar_accfree);
                    }
                    if (l == 0 && gather) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float4 w = tab_wx[c0 + j];
                            v[j] = fmaf(w.z, rz, fmaf(w.y, ry, fmaf(w.x, rx, v[j])));
                        }
                    }
                    if (!lastl) {
                        float hi[32], lo[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float2 s2 = tss[c0 + j];
                            tc::tf32_split(fmaxf(fmaf(s2.x, v[j], s2.y), 0.f), hi[j], lo[j]);
                        }
                        tc::tmem_st32_nowait(trow + colA23 + c0, hi);
                        tc::tmem_st32_nowait(trow + colA23 + n + c0, lo);
                        continue;
                    }
                    // ---- last computed layer: statistics, optional store, optional pooling ----
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float s = 0.f, sq = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) { s += v[8 * i + j]; sq = fmaf(v[8 * i + j], v[8 * i + j], sq); }
                        const int g = (c_off + c0 + 8 * i) / gsz;
#pragma unroll
                        for (int gg = 0; gg < kGnGroups; ++gg)
                            if (g == gg) { fsum[gg] += s; fsq[gg] += sq; }
                    }
                    if (q.y_out) {
                        float *yo = q.y_out + (static_cast<size_t>(b) * q.c_total + c_off + c0) * P + p0 + et;
#pragma unroll
                        for (int j = 0; j < 32; ++j) yo[static_cast<size_t>(j) * P] = v[j];
                    }
                    if (q.ymax) {
                        uint32_t kmx = 0, kmn = 0;
                        int imx = 0, imn = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const uint32_t key = f2key(v[j]);
                            const uint32_t mx = __reduce_max_sync(OGC_FULL_MASK, key), mn = __reduce_min_sync(OGC_FULL_MASK, key);
                            const uint32_t bmx = __ballot_sync(OGC_FULL_MASK, key == mx), bmn = __ballot_sync(OGC_FULL_MASK, key == mn);
                            if (lane == j) { kmx = mx; kmn = mn; imx = __ffs(bmx) - 1; imn = __ffs(bmn) - 1; }
                        }
                        const size_t o = (static_cast<size_t>(b) * 2 * q.M + static_cast<size_t>(t) * 4 + ew) * q.c_total + c_off + c0 + lane;
                        q.ymax[o] = key2f(kmx);
                        q.ymin[o] = key2f(kmn);
                        q.amax[o] = static_cast<unsigned char>(imx);
                        q.amin[o] = static_cast<unsigned char>(imn);
                    }
                }
                if (!lastl) {
                    tc::tmem_st_wait();
                    tc::fence_before_sync();
                    mbar_arrive(&bar_a[l + 1]);
                }
            }
#pragma unroll
            for (int g = 0; g < kGnGroups; ++g) { dsum[g] += static_cast<double>(fsum[g]); dsq[g] += static_cast<double>(fsq[g]); }
        }
#pragma unroll
        for (int g = 0; g < kGnGroups; ++g) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                dsum[g] += __shfl_xor_sync(OGC_FULL_MASK, dsum[g], o);
                dsq[g] += __shfl_xor_sync(OGC_FULL_MASK, dsq[g], o);
            }
            if (lane == 0) { atomicAdd(&gs[2 * g], dsum[g]); atomicAdd(&gs[2 * g + 1], dsq[g]); }
        }
        named_bar_sync(kEpiBar, 128);
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
    const uint32_t tab_bytes = 3u * kMaxC * 8u + kMaxC * 16u;
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
    // tensor-memory columns: A1 (hi | lo), A2 / A3 (hi | lo), one accumulator
    int cols = align_up(2 * q.K1, 32);
    if (nl > 1) cols += 2 * (nl > 2 ? (widths[0] > widths[1] ? widths[0] : widths[1]) : widths[0]);
    int accw = 0;
    for (int l = 0; l < nl; ++l) accw = q.C[l] > accw ? q.C[l] : accw;
    return cols + accw <= 512;
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
