// Shared pieces of the chained set-abstraction kernels (sa_chain_fwd.cu, sa_chain_bwd.cu).
//
// Orientation (round 2): POSITIONS on the MMA's M axis, CHANNELS on N.
//   D[p][co] = sum_k a[p][k] W[co][k]       M = 128 positions (= 2 centres x nsample 64) = the 128 TMEM lanes,
//                                           N = C_out (32..256) accumulator columns, K = C_in
//   A = activations, written to TENSOR MEMORY by the thread that owns the position (tcgen05.st, lane = position,
//       column = channel; hi block then lo block of the 3xTF32 split) and read from there by the MMA (.ts form)
//   B = weights, resident in shared memory for the CTA's lifetime: K-major 128B-swizzled column blocks
//       [2 N rows][32 k], rows 0..N-1 = W_hi, rows N..2N-1 = W_lo
// A layer's accumulator row is read back by the SAME thread (tcgen05.ld, thread = position), which applies
// GroupNorm + ReLU per channel and stores the next layer's A operand straight back into tensor memory: a chain of
// layers never leaves the SM.  Global tensors stay channel-major (B,C,P): for a fixed channel the 32 lanes of a warp
// touch 32 consecutive positions = one 128-byte segment per load / store instruction, no staging.
#pragma once
#include "mlp_common.cuh"
#include "tcgen05.cuh"

namespace ogc {
namespace chain {

constexpr int kTile = 128;        // positions per tile
constexpr int kNS = 64;           // nsample (positions per centre)
constexpr int kEpiWarps = 8;      // warps 0-7 epilogue: lane quadrant = warp & 3, column group = warp >> 2
constexpr int kEpi = kEpiWarps * 32;
constexpr int kProdWarp0 = kEpiWarps;               // warps 8-15 producer: lane quadrant = warp & 3, chunk group = (warp - 8) >> 2
constexpr int kMmaWarp = kProdWarp0 + 8;            // warp 16 MMA issuer / TMEM owner
constexpr int kThreads = (kMmaWarp + 1) * 32;       // 544: one warp per scheduler and role left every stage latency-bound
constexpr int kProdBar = 2, kEpiBar = 4;            // named barriers: 2, 3 = producer groups, 4 = epilogue
constexpr int kDensePitch = kTile * 4 + 16;         // raw-stage row pitch of a stored channel row (128 positions)
constexpr int kMaxC = 256;

__host__ __device__ constexpr int align_up(int v, int a) { return (v + a - 1) / a * a; }

// bytes of the resident B operand of a layer: KB column blocks of [2N rows][128 B]
__host__ __device__ constexpr uint32_t w_tile_bytes(int n_rows, int k) {
    return static_cast<uint32_t>(align_up(k, 32) / 32) * 2u * static_cast<uint32_t>(n_rows) * 128u;
}

// Build the resident B operand of one layer: rows r < n_rows of W (row-major, leading dimension ldw), K-columns
// k -> W[r][kmap(k)] for k < k_real, zero padding up to the next multiple of 32.  All threads of the CTA.
template <typename KMap>
__device__ __forceinline__ void build_weights(uint8_t *dst, const float *__restrict__ W, int ldw, int n_rows, int k_real,
                                              KMap kmap, int tid, int nthreads) {
    const int kp = align_up(k_real, 32);
    const uint32_t blk = 2u * static_cast<uint32_t>(n_rows) * 128u;
    for (int e = tid; e < n_rows * kp; e += nthreads) {
        const int r = e / kp, k = e - r * kp;
        const float v = k < k_real ? __ldg(W + static_cast<size_t>(r) * ldw + kmap(k)) : 0.f;
        const float hi = tc::tf32_hi(v), lo = tc::tf32_hi(v - hi);
        const uint32_t off = static_cast<uint32_t>(k >> 5) * blk + tc::sw128_offset(r, k & 31);
        *reinterpret_cast<float *>(dst + off) = hi;
        *reinterpret_cast<float *>(dst + off + static_cast<uint32_t>(n_rows) * 128u) = lo;
    }
}

// One layer's MMAs (executed by ALL lanes of the issuing warp, one elected lane issues): D[tmem_d] = A[tmem, hi at a_col, lo at a_col + k] x B[smem tile]^T, 3xTF32.
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t tmem_a, int k, uint32_t b_smem, int n) {
    const uint32_t idesc = tc::make_idesc_tf32(kTile, n, 0, 0);
    const uint32_t blk16 = (2u * static_cast<uint32_t>(n) * 128u) >> 4, lo16 = (static_cast<uint32_t>(n) * 128u) >> 4;
    const uint64_t d0 = tc::make_desc_sw128(b_smem, 16, 1024);           // start-address field += byte offset >> 4
    const int kblocks = k >> 5;
    for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const uint64_t bh = d0 + (static_cast<uint32_t>(kb) * blk16 + static_cast<uint32_t>(s) * 2u), bl = bh + lo16;
            const uint32_t ah = tmem_a + static_cast<uint32_t>(kb * 32 + s * 8), al = ah + static_cast<uint32_t>(k);
            tc::mma_tf32_ts_elect(tmem_d, ah, bh, idesc, (kb | s) ? 1u : 0u);
            tc::mma_tf32_ts_elect(tmem_d, ah, bl, idesc, 1u);
            tc::mma_tf32_ts_elect(tmem_d, al, bh, idesc, 1u);
        }
    }
}

// order-preserving map float -> uint32 (for redux.sync max / min) and back
__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ (static_cast<uint32_t>(static_cast<int32_t>(u) >> 31) | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

}  // namespace chain
}  // namespace ogc
