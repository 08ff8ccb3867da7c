// Narrow SharedMLP layers of the first set-abstraction level (32 -> 32 and 32 -> 64 channels over
// B * M * 64 = 2.1 M positions at KITTI-SF sizes): forward, input gradient and weight gradient.
//
// These layers move ~1 GB each and need only 2-4 KFLOP per position, so they are bound by how many bytes a
// CTA keeps in flight, not by the contraction.  The tiled kernels (mlp.cu / mlp_tc*.cu: shared-memory operand
// tiles, barriers and -- on the tensor-core path -- a 64-position tile per MMA round trip) reach 1.2-2.3 TB/s
// here (profiles/r01_ncu_full_sa_mlp_tc_kernels.csv: DRAM bytes == algorithmic bytes, 20 % warp occupancy,
// issue slots idle).  This file maps the work the other way round:
//
//   forward / dX : a WARP owns the 64 samples of one centre, a LANE owns two adjacent positions and keeps ALL
//                  their channels in registers; every global access is a coalesced 256-byte row segment
//                  (channel-major tensors), 64 of them in flight per warp before the first use; the weights are
//                  broadcast from shared memory (one LDS.128 per 8 FMAs); no barrier inside the main loop.
//                  GroupNorm+ReLU of the previous layer is applied in registers, the GroupNorm sums fall out of
//                  the accumulators, the max / min over the 64 samples is two redux.sync + two ballots per
//                  channel, and the per-channel backward sums use a 31-shuffle transpose-reduction.
//   dW           : C_out x 32 outputs reduced over millions of positions: dY and a tiles are staged channel-major
//                  in shared memory, lane = input channel, each warp a slice of the tile's positions, the whole
//                  C_out column of partial sums in registers across all tiles of the CTA.
//
// fp32 FMA throughout (bit-level fp32 conv semantics).  Same operand / result contract as the generic kernels,
// which the tests compare them against.
#include <type_traits>

#include "mlp_common.cuh"
#include "mlp_dy.cuh"

namespace ogc {

constexpr int kNwThreads = 256;
constexpr int kNwWarps = kNwThreads / 32;
constexpr int kNwUnit = 64;   // positions per warp step == nsample

// Weights of the layer whose dX is being computed, (COUT,32) row-major, copied device-to-device in stream order
// before the launch.  Every lane of a warp needs the same weight at the same time: from shared memory that is a
// broadcast LDS.128, which still occupies the SM's load/store path that the dX kernel also needs for its three global
// streams; through the constant bank (LDC) the weights bypass it.  Measured: dX 1.18 -> 1.03 ms per step with the
// constant bank, forward 0.98 -> 1.18 ms (its LSU is otherwise idle), so only dX uses it.
__constant__ float4 cNwW[64 * 32 / 4];

// order-preserving map float -> uint32 (for redux.sync max / min)
__device__ __forceinline__ uint32_t f2ord(float x) {
    const uint32_t b = __float_as_uint(x);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}


// ------------------------------------------------------------------------------------------ forward
struct NarrowFwdParams {
    int P, M;
    const float *y_prev, *ss_prev;   // (B,CIN,P), (B,CIN,2)
    const float *W;                  // (COUT,CIN) row-major
    const float *gamma;              // (COUT) GroupNorm weight of THIS layer (LAST: decides max vs min per channel)
    float *y;                        // (B,COUT,P)
    double *sums;                    // (B,4,2)
    float *ymax, *ymin;              // (B,COUT,M) when LAST
    unsigned char *amax, *amin;
};

template <int CIN, int COUT, bool LAST>
__global__ void __launch_bounds__(kNwThreads, 2)
narrow_fwd_kernel(NarrowFwdParams q) {
    constexpr int CPG = COUT / kGnGroups;
    __shared__ __align__(16) float Ws[COUT * CIN];
    __shared__ float2 ss_s[CIN];
    __shared__ uint32_t flip_s[COUT];   // LAST: 0xffffffff for channels whose pooled value is the MINIMUM (gamma < 0)
    __shared__ double gs[kGnGroups][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, P = q.P;
    for (int e = tid; e < COUT * CIN; e += kNwThreads) Ws[e] = __ldg(q.W + e);
    for (int c = tid; c < CIN; c += kNwThreads) ss_s[c] = __ldg(reinterpret_cast<const float2 *>(q.ss_prev) + static_cast<size_t>(b) * CIN + c);
    if (LAST)
        for (int c = tid; c < COUT; c += kNwThreads) flip_s[c] = __ldg(q.gamma + c) < 0.f ? 0xffffffffu : 0u;
    if (tid < kGnGroups * 2) (&gs[0][0])[tid] = 0.0;
    __syncthreads();
    const float4 *Ws4 = reinterpret_cast<const float4 *>(Ws);

    float gsum0 = 0.f, gsum1 = 0.f, gsum2 = 0.f, gsum3 = 0.f, gsq0 = 0.f, gsq1 = 0.f, gsq2 = 0.f, gsq3 = 0.f;

    const int nunits = P / kNwUnit;
    for (int u = blockIdx.x * kNwWarps + warp; u < nunits; u += gridDim.x * kNwWarps) {
        const size_t p0 = static_cast<size_t>(u) * kNwUnit + 2 * lane;
        const float *src = q.y_prev + static_cast<size_t>(b) * CIN * P + p0;
        float a0[CIN], a1[CIN];
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(src + static_cast<size_t>(c) * P));
            a0[c] = v.x; a1[c] = v.y;
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
            const float2 s = ss_s[c];
            a0[c] = fmaxf(fmaf(s.x, a0[c], s.y), 0.f);
            a1[c] = fmaxf(fmaf(s.x, a1[c], s.y), 0.f);
        }
        float *dst = q.y + static_cast<size_t>(b) * COUT * P + p0;
        float pool_v = 0.f;        // LAST: lane c collects the pooled value / slot of channel (32 k + c)
        int pool_p = 0;
#pragma unroll 1
        for (int co0 = 0; co0 < COUT; co0 += 4) {
            // packed FMAs over pairs of input channels: (even-c partial, odd-c partial) per output, summed at the end
            float2 e0[4], e1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) e0[j] = e1[j] = make_float2(0.f, 0.f);
#pragma unroll
            for (int c4 = 0; c4 < CIN / 4; ++c4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w = Ws4[(co0 + j) * (CIN / 4) + c4];
                    e0[j] = ffma2(make_float2(w.x, w.y), make_float2(a0[c4 * 4 + 0], a0[c4 * 4 + 1]), e0[j]);
                    e1[j] = ffma2(make_float2(w.x, w.y), make_float2(a1[c4 * 4 + 0], a1[c4 * 4 + 1]), e1[j]);
                    e0[j] = ffma2(make_float2(w.z, w.w), make_float2(a0[c4 * 4 + 2], a0[c4 * 4 + 3]), e0[j]);
                    e1[j] = ffma2(make_float2(w.z, w.w), make_float2(a1[c4 * 4 + 2], a1[c4 * 4 + 3]), e1[j]);
                }
            }
            float r0[4], r1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { r0[j] = e0[j].x + e0[j].y; r1[j] = e1[j].x + e1[j].y; }
            float s = 0.f, sq = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                *reinterpret_cast<float2 *>(dst + static_cast<size_t>(co0 + j) * P) = make_float2(r0[j], r1[j]);
                s += r0[j] + r1[j];
                sq += r0[j] * r0[j] + r1[j] * r1[j];
                if (LAST) {
                    // pooled extreme of the 64 samples and its FIRST slot: one redux + two ballots per channel
                    const uint32_t flip = flip_s[co0 + j];
                    const uint32_t k0 = f2ord(r0[j]) ^ flip, k1 = f2ord(r1[j]) ^ flip;
                    const uint32_t kbest = __reduce_max_sync(OGC_FULL_MASK, max(k0, k1));
                    const unsigned e0 = __ballot_sync(OGC_FULL_MASK, k0 == kbest), e1 = __ballot_sync(OGC_FULL_MASK, k1 == kbest);
                    const int q0 = e0 ? 2 * (__ffs(e0) - 1) : 64, q1 = e1 ? 2 * (__ffs(e1) - 1) + 1 : 64;   // even slots: r0, odd: r1
                    if (lane == ((co0 + j) & 31)) { pool_v = ord2f(kbest ^ flip); pool_p = min(q0, q1); }
                }
            }
            const int g = co0 / CPG;       // the 4 channels of a step share their group (CPG % 4 == 0)
            gsum0 += g == 0 ? s : 0.f; gsq0 += g == 0 ? sq : 0.f;
            gsum1 += g == 1 ? s : 0.f; gsq1 += g == 1 ? sq : 0.f;
            gsum2 += g == 2 ? s : 0.f; gsq2 += g == 2 ? sq : 0.f;
            gsum3 += g == 3 ? s : 0.f; gsq3 += g == 3 ? sq : 0.f;
            if (LAST && ((co0 + 4) & 31) == 0) {
                // both arrays get the value sa_finish will pick (it chooses by the sign of gamma * rstd)
                const size_t o = (static_cast<size_t>(b) * COUT + (co0 + 4 - 32) + lane) * q.M + u;
                q.ymax[o] = pool_v; q.ymin[o] = pool_v;
                q.amax[o] = static_cast<unsigned char>(pool_p); q.amin[o] = static_cast<unsigned char>(pool_p);
            }
        }
    }
    float gsum[kGnGroups] = {gsum0, gsum1, gsum2, gsum3}, gsq[kGnGroups] = {gsq0, gsq1, gsq2, gsq3};
#pragma unroll
    for (int g = 0; g < kGnGroups; ++g) {
        float s = gsum[g], sq = gsq[g];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(OGC_FULL_MASK, s, o);
            sq += __shfl_xor_sync(OGC_FULL_MASK, sq, o);
        }
        if (lane == 0) {
            atomicAdd(&gs[g][0], static_cast<double>(s));
            atomicAdd(&gs[g][1], static_cast<double>(sq));
        }
    }
    __syncthreads();
    if (tid < kGnGroups * 2) atomicAdd(q.sums + static_cast<size_t>(b) * kGnGroups * 2 + tid, (&gs[0][0])[tid]);
}

// ------------------------------------------------------------------------------------------ dX
struct NarrowDxParams {
    int P, M;
    const float *dz;                  // (B,COUT,P), or NULL: synthesised from go / sel (last layer)
    const float *go;                  // (B,go_ctotal,M)
    const unsigned char *sel;         // (B,COUT,M)
    int go_ctotal, go_coff;
    const float *y, *coef;            // (B,COUT,P), (B,COUT,4) = k1, k2, k3r, mean
    const float *W;                   // (COUT,CPREV) row-major
    const float *y_prev, *ss_prev, *mean_rstd_prev, *gamma_prev;
    float *dz_prev;                   // (B,CPREV,P)
    double *ab_prev;                  // (B,4,2)
    float *dgamma_prev, *dbeta_prev;
};

// One pipeline stage of the dX kernel: 4 channels x (y, dz) x 2 positions, or 8 channels of y_prev (epilogue)
struct NwStage {
    float2 v[8];
};

template <int COUT, bool SYNTH>
__global__ void __launch_bounds__(kNwThreads, 2)
narrow_dx_kernel(NarrowDxParams q) {
    constexpr int CPREV = 32, CH = 4;
    __shared__ __align__(16) float4 coef_s[COUT];
    __shared__ float4 prev_s[CPREV];            // scale, shift, mean, rstd of the previous layer's channels
    __shared__ float rowacc[CPREV][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, P = q.P;
    for (int c = tid; c < COUT; c += kNwThreads) coef_s[c] = __ldg(reinterpret_cast<const float4 *>(q.coef) + static_cast<size_t>(b) * COUT + c);
    for (int c = tid; c < CPREV; c += kNwThreads) {
        const int g = c / (CPREV / kGnGroups);
        prev_s[c] = make_float4(__ldg(q.ss_prev + (static_cast<size_t>(b) * CPREV + c) * 2), __ldg(q.ss_prev + (static_cast<size_t>(b) * CPREV + c) * 2 + 1),
                                __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2), __ldg(q.mean_rstd_prev + (b * kGnGroups + g) * 2 + 1));
        rowacc[c][0] = rowacc[c][1] = 0.f;
    }
    __syncthreads();
    float tot_s = 0.f, tot_sy = 0.f;            // lane = channel of layer l-1

    const int nunits = P / kNwUnit;
    const int u_first = blockIdx.x * kNwWarps + warp, u_step = gridDim.x * kNwWarps;

    // stage loaders: every load of a stage is issued before any is consumed; the NEXT stage is always in flight
    // while the current one is being multiplied (software pipeline across the channel loop, the epilogue and units)
    auto load_main = [&](NwStage &st, int u, int cb) {
        const size_t p0 = static_cast<size_t>(u) * kNwUnit + 2 * lane;
        const float *yp = q.y + (static_cast<size_t>(b) * COUT + cb) * P + p0;
#pragma unroll
        for (int j = 0; j < CH; ++j) st.v[j] = __ldg(reinterpret_cast<const float2 *>(yp + static_cast<size_t>(j) * P));
        if (SYNTH) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const size_t o = (static_cast<size_t>(b) * COUT + cb + j) * q.M + u;
                const int sl = __ldg(q.sel + o);
                const float g = __ldg(q.go + (static_cast<size_t>(b) * q.go_ctotal + q.go_coff + cb + j) * q.M + u);
                st.v[CH + j] = make_float2(sl == 2 * lane ? g : 0.f, sl == 2 * lane + 1 ? g : 0.f);
            }
        } else {
            const float *zp = q.dz + (static_cast<size_t>(b) * COUT + cb) * P + p0;
#pragma unroll
            for (int j = 0; j < CH; ++j) st.v[CH + j] = __ldg(reinterpret_cast<const float2 *>(zp + static_cast<size_t>(j) * P));
        }
    };
    auto load_prev = [&](NwStage &st, int u, int cb) {
        const float *pp = q.y_prev + (static_cast<size_t>(b) * CPREV + cb) * P + static_cast<size_t>(u) * kNwUnit + 2 * lane;
#pragma unroll
        for (int j = 0; j < 8; ++j) st.v[j] = __ldg(reinterpret_cast<const float2 *>(pp + static_cast<size_t>(j) * P));
    };

    float2 acc0[CPREV / 2], acc1[CPREV / 2];     // (channel 2k, channel 2k+1) of the two positions
    auto mul_main = [&](const NwStage &st, int cb) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const float4 cf = coef_s[cb + j];
            const float d0 = fmaf(cf.x, st.v[CH + j].x, -cf.y) - (st.v[j].x - cf.w) * cf.z;
            const float d1 = fmaf(cf.x, st.v[CH + j].y, -cf.y) - (st.v[j].y - cf.w) * cf.z;
            const float2 dd0 = make_float2(d0, d0), dd1 = make_float2(d1, d1);
#pragma unroll
            for (int c4 = 0; c4 < CPREV / 4; ++c4) {
                const float4 w = cNwW[(cb + j) * (CPREV / 4) + c4];
                acc0[c4 * 2 + 0] = ffma2(dd0, make_float2(w.x, w.y), acc0[c4 * 2 + 0]);
                acc1[c4 * 2 + 0] = ffma2(dd1, make_float2(w.x, w.y), acc1[c4 * 2 + 0]);
                acc0[c4 * 2 + 1] = ffma2(dd0, make_float2(w.z, w.w), acc0[c4 * 2 + 1]);
                acc1[c4 * 2 + 1] = ffma2(dd1, make_float2(w.z, w.w), acc1[c4 * 2 + 1]);
            }
        }
    };

    NwStage sa, sb;
    if (u_first < nunits) load_main(sa, u_first, 0);
    for (int u = u_first; u < nunits; u += u_step) {
#pragma unroll
        for (int i = 0; i < CPREV / 2; ++i) acc0[i] = acc1[i] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int cb = 0; cb < COUT; cb += 2 * CH) {
            load_main(sb, u, cb + CH);
            mul_main(sa, cb);
            if (cb + 2 * CH < COUT) load_main(sa, u, cb + 2 * CH);
            else load_prev(sa, u, 0);
            mul_main(sb, cb + CH);
        }
        // ---- epilogue: ReLU mask of layer l-1, store dz_{l-1}, per-channel GroupNorm-backward sums ----
        float *dp = q.dz_prev + static_cast<size_t>(b) * CPREV * P + static_cast<size_t>(u) * kNwUnit + 2 * lane;
        float fs[CPREV], fsy[CPREV];
        auto finish8 = [&](const NwStage &st, auto cbc) {
            constexpr int cb = decltype(cbc)::value;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 pr = prev_s[cb + j];
                const float x0 = (j & 1) ? acc0[(cb + j) >> 1].y : acc0[(cb + j) >> 1].x;
                const float x1 = (j & 1) ? acc1[(cb + j) >> 1].y : acc1[(cb + j) >> 1].x;
                const float g0 = fmaf(pr.x, st.v[j].x, pr.y) > 0.f ? x0 : 0.f;
                const float g1 = fmaf(pr.x, st.v[j].y, pr.y) > 0.f ? x1 : 0.f;
                *reinterpret_cast<float2 *>(dp + static_cast<size_t>(cb + j) * P) = make_float2(g0, g1);
                fs[cb + j] = g0 + g1;
                fsy[cb + j] = g0 * ((st.v[j].x - pr.z) * pr.w) + g1 * ((st.v[j].y - pr.z) * pr.w);
            }
        };
        load_prev(sb, u, 8);
        finish8(sa, std::integral_constant<int, 0>{});
        load_prev(sa, u, 16);
        finish8(sb, std::integral_constant<int, 8>{});
        load_prev(sb, u, 24);
        finish8(sa, std::integral_constant<int, 16>{});
        if (u + u_step < nunits) load_main(sa, u + u_step, 0);      // next unit's first stage
        finish8(sb, std::integral_constant<int, 24>{});
        tot_s += warp_transpose_sum32(fs, lane);
        tot_sy += warp_transpose_sum32(fsy, lane);
    }
    atomicAdd(&rowacc[lane][0], tot_s);
    atomicAdd(&rowacc[lane][1], tot_sy);
    __syncthreads();
    if (tid < CPREV) {
        const float a = rowacc[tid][0], c = rowacc[tid][1];
        atomicAdd(q.dbeta_prev + tid, a);
        atomicAdd(q.dgamma_prev + tid, c);
        const int g = tid / (CPREV / kGnGroups);
        const double gm = static_cast<double>(__ldg(q.gamma_prev + tid));
        atomicAdd(q.ab_prev + (b * kGnGroups + g) * 2, gm * a);
        atomicAdd(q.ab_prev + (b * kGnGroups + g) * 2 + 1, gm * c);
    }
}

// ------------------------------------------------------------------------------------------ dW
struct NarrowDwParams {
    DySrc dy;                        // layer l (rows of dW = COUT)
    int B;
    const float *y_prev, *ss_prev;   // (B,32,P), (B,32,2)
    float *dW;                       // (COUT,32), accumulated atomically
};

// One CTA handles 32 output channels [co_off, co_off+32) of layer l (C_out = 64: two CTAs streams via blockIdx.y).
// lane = (position half h, input-channel pair {cq, cq+16}): per LDS.128 of the dY tile (2 positions x 2 output
// channels) a lane issues 4 FFMA2 -- twice the arithmetic per shared-memory byte of a lane-per-input-channel mapping,
// which matters because a warp-wide LDS.128 occupies the load/store path for 4 cycles even when it broadcasts.
__global__ void __launch_bounds__(kNwThreads, 2)
narrow_dw_kernel(NarrowDwParams q) {
    constexpr int CIN = 32, CO = 32, TP = 128, LD = TP + 4, LDD = 2 * TP + 8, Q4 = TP / 4;
    extern __shared__ __align__(16) float smem[];
    float *Ds = smem;                      // [CO/2][LDD]: (dY[2k][p], dY[2k+1][p]) interleaved per position
    float *As = smem + (CO / 2) * LDD;     // [CIN][LD]   a_{l-1}
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = lane >> 4, cq = lane & 15;
    const int co_off = blockIdx.y * CO;
    const int P = q.dy.P;
    const int tiles_per_sample = P / TP;
    const int total = q.B * tiles_per_sample;
    float2 acc0[CO / 2], acc1[CO / 2];     // (dW[2k][ci], dW[2k+1][ci]) for ci = cq and ci = cq + 16
#pragma unroll
    for (int i = 0; i < CO / 2; ++i) acc0[i] = acc1[i] = make_float2(0.f, 0.f);

    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int b = w / tiles_per_sample, p_base = (w - b * tiles_per_sample) * TP;
        __syncthreads();
        // a tile: thread -> (channel, quad); 4 independent 16-byte loads in flight
        {
            float4 v[CIN * Q4 / kNwThreads];
#pragma unroll
            for (int i = 0; i < CIN * Q4 / kNwThreads; ++i) {
                const int e = tid + i * kNwThreads, c = e / Q4, p = (e - c * Q4) * 4;
                v[i] = __ldg(reinterpret_cast<const float4 *>(q.y_prev + (static_cast<size_t>(b) * CIN + c) * P + p_base + p));
            }
#pragma unroll
            for (int i = 0; i < CIN * Q4 / kNwThreads; ++i) {
                const int e = tid + i * kNwThreads, c = e / Q4, p = (e - c * Q4) * 4;
                const float s = __ldg(q.ss_prev + (static_cast<size_t>(b) * CIN + c) * 2), hh = __ldg(q.ss_prev + (static_cast<size_t>(b) * CIN + c) * 2 + 1);
                *reinterpret_cast<float4 *>(As + c * LD + p) = make_float4(fmaxf(fmaf(s, v[i].x, hh), 0.f), fmaxf(fmaf(s, v[i].y, hh), 0.f),
                                                                          fmaxf(fmaf(s, v[i].z, hh), 0.f), fmaxf(fmaf(s, v[i].w, hh), 0.f));
            }
        }
        // dY tile: thread -> (channel PAIR, quad): both channels' quads are loaded, interleaved, stored as 2 x 16 B
        {
            DyRaw raw[2][2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int e = tid + k * kNwThreads, cp = e / Q4, p = (e - cp * Q4) * 4;
                dy_quad_load(q.dy, b, co_off + 2 * cp, p_base + p, raw[k][0]);
                dy_quad_load(q.dy, b, co_off + 2 * cp + 1, p_base + p, raw[k][1]);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int e = tid + k * kNwThreads, cp = e / Q4, p = (e - cp * Q4) * 4;
                const float4 d0 = dy_quad_finish(q.dy, p_base + p, raw[k][0]), d1 = dy_quad_finish(q.dy, p_base + p, raw[k][1]);
                float4 *dst = reinterpret_cast<float4 *>(Ds + cp * LDD + 2 * p);
                dst[0] = make_float4(d0.x, d1.x, d0.y, d1.y);
                dst[1] = make_float4(d0.z, d1.z, d0.w, d1.w);
            }
        }
        __syncthreads();
        // warp = 16 positions of the tile, half-warp h = 8 of them
#pragma unroll
        for (int pq = 0; pq < 2; ++pq) {
            const int p = warp * (TP / kNwWarps) + h * 8 + pq * 4;
            const float4 a0 = *reinterpret_cast<const float4 *>(As + cq * LD + p);
            const float4 a1 = *reinterpret_cast<const float4 *>(As + (cq + 16) * LD + p);
#pragma unroll
            for (int cp = 0; cp < CO / 2; ++cp) {
                const float4 d01 = *reinterpret_cast<const float4 *>(Ds + cp * LDD + 2 * p);
                const float4 d23 = *reinterpret_cast<const float4 *>(Ds + cp * LDD + 2 * p + 4);
                const float2 e0 = make_float2(d01.x, d01.y), e1 = make_float2(d01.z, d01.w), e2 = make_float2(d23.x, d23.y), e3 = make_float2(d23.z, d23.w);
                acc0[cp] = ffma2(e0, make_float2(a0.x, a0.x), acc0[cp]); acc1[cp] = ffma2(e0, make_float2(a1.x, a1.x), acc1[cp]);
                acc0[cp] = ffma2(e1, make_float2(a0.y, a0.y), acc0[cp]); acc1[cp] = ffma2(e1, make_float2(a1.y, a1.y), acc1[cp]);
                acc0[cp] = ffma2(e2, make_float2(a0.z, a0.z), acc0[cp]); acc1[cp] = ffma2(e2, make_float2(a1.z, a1.z), acc1[cp]);
                acc0[cp] = ffma2(e3, make_float2(a0.w, a0.w), acc0[cp]); acc1[cp] = ffma2(e3, make_float2(a1.w, a1.w), acc1[cp]);
            }
        }
    }
    // reduce the 16 half-warps' partial sums through shared memory, then one atomic per output
    __syncthreads();
    float *red = smem;               // [2 * kNwWarps][CO][33]
    const int hw = warp * 2 + h;
#pragma unroll
    for (int cp = 0; cp < CO / 2; ++cp) {
        red[(hw * CO + 2 * cp) * 33 + cq] = acc0[cp].x;      red[(hw * CO + 2 * cp + 1) * 33 + cq] = acc0[cp].y;
        red[(hw * CO + 2 * cp) * 33 + cq + 16] = acc1[cp].x; red[(hw * CO + 2 * cp + 1) * 33 + cq + 16] = acc1[cp].y;
    }
    __syncthreads();
    for (int e = tid; e < CO * CIN; e += kNwThreads) {
        const int co = e / CIN, ci = e - co * CIN;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 2 * kNwWarps; ++w) s += red[(w * CO + co) * 33 + ci];
        atomicAdd(q.dW + static_cast<size_t>(co_off + co) * CIN + ci, s);
    }
}

static int narrow_grid_x(int B, int nunits) {
    int per_sample = (kNumSMs * 2) / B;   // floor: the whole grid must be resident at 2 CTAs/SM (a 297th CTA would run as a second wave)
    const int need = (nunits + kNwWarps - 1) / kNwWarps;
    if (per_sample > need) per_sample = need;
    return per_sample < 1 ? 1 : per_sample;
}

}  // namespace ogc

extern "C" int ogc_sa_mlp_narrow_fwd(int b, int m, int nsample, int cin, int cout, int last, const float *y_prev,
                                     const float *ss_prev, const float *w, const float *gamma, float *y, double *sums,
                                     float *ymax, float *ymin, unsigned char *amax, unsigned char *amin, void *stream) {
    using namespace ogc;
    if (b < 0 || m <= 0 || nsample <= 0 || cin <= 0 || cout <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!y_prev || !ss_prev || !w || !y || !sums) return OGC_ERR_INVALID_ARG;
    if (last && (!ymax || !ymin || !amax || !amin || !gamma)) return OGC_ERR_INVALID_ARG;
    if (nsample != 64 || cin != 32 || (cout != 32 && cout != 64) || b > 65535) return OGC_ERR_UNSUPPORTED;
    NarrowFwdParams q;
    q.P = m * nsample; q.M = m; q.y_prev = y_prev; q.ss_prev = ss_prev; q.W = w; q.gamma = gamma; q.y = y; q.sums = sums;
    q.ymax = ymax; q.ymin = ymin; q.amax = amax; q.amin = amin;
    dim3 grid(narrow_grid_x(b, m), b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cout == 32) {
        if (last) narrow_fwd_kernel<32, 32, true><<<grid, kNwThreads, 0, st>>>(q);
        else narrow_fwd_kernel<32, 32, false><<<grid, kNwThreads, 0, st>>>(q);
    } else {
        if (last) narrow_fwd_kernel<32, 64, true><<<grid, kNwThreads, 0, st>>>(q);
        else narrow_fwd_kernel<32, 64, false><<<grid, kNwThreads, 0, st>>>(q);
    }
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_sa_mlp_narrow_dx(int b, int m, int nsample, int cout, int cprev, const float *dz, const float *go,
                                    int go_ctotal, int go_coff, const unsigned char *sel, const float *y,
                                    const float *coef, const float *w, const float *y_prev, const float *ss_prev,
                                    const float *mean_rstd_prev, const float *gamma_prev, float *dz_prev,
                                    double *ab_prev, float *dgamma_prev, float *dbeta_prev, void *stream) {
    using namespace ogc;
    if (b < 0 || m <= 0 || nsample <= 0 || cout <= 0 || cprev <= 0) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!y || !coef || !w || !y_prev || !ss_prev || !mean_rstd_prev || !gamma_prev || !dz_prev || !ab_prev || !dgamma_prev || !dbeta_prev)
        return OGC_ERR_INVALID_ARG;
    if (!dz && (!go || !sel)) return OGC_ERR_INVALID_ARG;
    if (nsample != 64 || cprev != 32 || (cout != 32 && cout != 64) || b > 65535) return OGC_ERR_UNSUPPORTED;
    NarrowDxParams q;
    q.P = m * nsample; q.M = m; q.dz = dz; q.go = go; q.sel = sel; q.go_ctotal = go_ctotal; q.go_coff = go_coff;
    q.y = y; q.coef = coef; q.W = w; q.y_prev = y_prev; q.ss_prev = ss_prev; q.mean_rstd_prev = mean_rstd_prev;
    q.gamma_prev = gamma_prev; q.dz_prev = dz_prev; q.ab_prev = ab_prev; q.dgamma_prev = dgamma_prev; q.dbeta_prev = dbeta_prev;
    dim3 grid(narrow_grid_x(b, m), b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t ce = cudaMemcpyToSymbolAsync(cNwW, w, static_cast<size_t>(cout) * 32 * sizeof(float), 0, cudaMemcpyDeviceToDevice, st);
    if (ce != cudaSuccess) return static_cast<int>(ce);
    if (cout == 32) {
        if (dz) narrow_dx_kernel<32, false><<<grid, kNwThreads, 0, st>>>(q);
        else narrow_dx_kernel<32, true><<<grid, kNwThreads, 0, st>>>(q);
    } else {
        if (dz) narrow_dx_kernel<64, false><<<grid, kNwThreads, 0, st>>>(q);
        else narrow_dx_kernel<64, true><<<grid, kNwThreads, 0, st>>>(q);
    }
    OGC_RETURN_LAUNCH_STATUS();
}

extern "C" int ogc_sa_mlp_narrow_dw(int b, int m, int nsample, int cout, int cin, const float *dz, const float *go,
                                    int go_ctotal, int go_coff, const unsigned char *sel, const float *y,
                                    const float *coef, const float *y_prev, const float *ss_prev, float *dw,
                                    void *stream) {
    using namespace ogc;
    if (b < 0 || m <= 0 || nsample <= 0 || cout <= 0 || cin <= 0 || !dw) return OGC_ERR_INVALID_ARG;
    if (b == 0) return OGC_OK;
    if (!y_prev || !ss_prev) return OGC_ERR_INVALID_ARG;
    if (nsample != 64 || cin != 32 || (cout != 32 && cout != 64)) return OGC_ERR_UNSUPPORTED;
    NarrowDwParams q;
    int rc = fill_dy(q.dy, cout, m, nsample, dz, go, go_ctotal, go_coff, sel, y, coef);
    if (rc != OGC_OK) return rc;
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(y_prev) | (dz ? reinterpret_cast<uintptr_t>(dz) : 0)) & 15u)
        return OGC_ERR_UNSUPPORTED;
    q.B = b; q.y_prev = y_prev; q.ss_prev = ss_prev; q.dW = dw;
    const int total = b * (m * nsample / 128);
    const int halves = cout / 32;                       // C_out = 64: two CTA streams, 32 output channels each
    int gx = (kNumSMs * 2) / halves;
    gx = total < gx ? total : gx;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem_tile = (static_cast<size_t>(16) * 264 + 32 * 132) * sizeof(float);
    const size_t smem_red = static_cast<size_t>(2 * kNwWarps) * 32 * 33 * sizeof(float);
    const size_t smem = smem_tile > smem_red ? smem_tile : smem_red;
    cudaError_t e = cudaFuncSetAttribute(narrow_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    narrow_dw_kernel<<<dim3(gx, halves), kNwThreads, smem, st>>>(q);
    OGC_RETURN_LAUNCH_STATUS();
}
