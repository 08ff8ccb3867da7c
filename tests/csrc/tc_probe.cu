// Diagnostic kernel: one 128 x N x K GEMM on the 5th-generation tensor core (tcgen05.mma kind::tf32,
// accumulator in TMEM, operands in 128B-swizzled shared-memory tiles), used by tests/test_gpu_tcgen05.py to
// pin the descriptor conventions of tcgen05.cuh (K-major and MN-major operands, single-pass TF32 and the
// 3xTF32 split) against torch on a B200 before the fused kernels rely on them.
#include "../../ogc_b200/csrc/tcgen05.cuh"

namespace ogc {

// mode 0: A (128,K) row-major, B (N,K) row-major (both K-major operands):   D = A B^T
// mode 1: A (K,128) row-major, B (K,N) row-major (both MN-major operands):  D = A^T B
// mode 2: as mode 0, but A is first written to TENSOR MEMORY (tcgen05.st, lane = row, column = k; hi then lo) and the
//         MMAs read it from there (the .ts form used by the fused MLP kernels for the stationary weight operand)
__global__ void __launch_bounds__(128)
tc_probe_kernel(int mode, int N, int K, int split3, const float *__restrict__ A, const float *__restrict__ B,
                float *__restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = 128;
    // tile sizes in bytes
    const uint32_t a_bytes = static_cast<uint32_t>(M) * K * 4, b_bytes = static_cast<uint32_t>(N) * K * 4;
    uint8_t *a_hi = smem, *b_hi = a_hi + a_bytes, *a_lo = b_hi + b_bytes, *b_lo = a_lo + a_bytes;

    uint32_t ncols = 32;
    while (ncols < static_cast<uint32_t>(N) + (mode == 2 ? 2u * K : 0u)) ncols <<= 1;
    const uint32_t a_col = static_cast<uint32_t>(N);      // TMEM column of A_hi (mode 2); A_lo follows at +K
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, ncols);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    // ---- fill the operand tiles ----
    if (mode == 0 || mode == 2) {
        // column blocks over K: block kb = [rows][32], rows = M (A) or N (B)
        for (int e = tid; e < M * K && mode == 0; e += blockDim.x) {
            const int r = e / K, k = e - r * K;
            const float v = A[e], hi = split3 ? tc::tf32_hi(v) : v;
            const uint32_t off = static_cast<uint32_t>(k >> 5) * (M * 128u) + tc::sw128_offset(r, k & 31);
            *reinterpret_cast<float *>(a_hi + off) = hi;
            if (split3) *reinterpret_cast<float *>(a_lo + off) = tc::tf32_hi(v - hi);
        }
        for (int e = tid; e < N * K; e += blockDim.x) {
            const int r = e / K, k = e - r * K;
            const float v = B[e];
            float hi = split3 ? tc::tf32_hi(v) : v, lo = tc::tf32_hi(v - hi);
            if (split3 && mode == 2) tc::tf32_split(v, hi, lo);
            const uint32_t off = static_cast<uint32_t>(k >> 5) * (N * 128u) + tc::sw128_offset(r, k & 31);
            *reinterpret_cast<float *>(b_hi + off) = hi;
            if (split3) *reinterpret_cast<float *>(b_lo + off) = lo;
        }
    } else {
        // column blocks over M / N: block cb = [K rows][32]
        for (int e = tid; e < K * M; e += blockDim.x) {
            const int k = e / M, m = e - k * M;
            const float v = A[e], hi = split3 ? tc::tf32_hi(v) : v;
            const uint32_t off = static_cast<uint32_t>(m >> 5) * (K * 128u) + tc::sw128_32b_offset(k, m & 31);
            *reinterpret_cast<float *>(a_hi + off) = hi;
            if (split3) *reinterpret_cast<float *>(a_lo + off) = tc::tf32_hi(v - hi);
        }
        for (int e = tid; e < K * N; e += blockDim.x) {
            const int k = e / N, n = e - k * N;
            const float v = B[e], hi = split3 ? tc::tf32_hi(v) : v;
            const uint32_t off = static_cast<uint32_t>(n >> 5) * (K * 128u) + tc::sw128_32b_offset(k, n & 31);
            *reinterpret_cast<float *>(b_hi + off) = hi;
            if (split3) *reinterpret_cast<float *>(b_lo + off) = tc::tf32_hi(v - hi);
        }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    if (mode == 2) {
        // thread = row of A: 32 columns per store
        for (int k0 = 0; k0 < K; k0 += 32) {
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float v = A[tid * K + k0 + j];
                hi[j] = v; lo[j] = 0.f;
                if (split3) tc::tf32_split(v, hi[j], lo[j]);     // the integer split the fused kernels use
            }
            const uint32_t ta = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + a_col + static_cast<uint32_t>(k0);
            tc::tmem_st32(ta, hi);
            if (split3) tc::tmem_st32(ta + static_cast<uint32_t>(K), lo);
        }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }

    if (tid == 0 && mode == 2) {
        const uint32_t idesc = tc::make_idesc_tf32(M, N, 0, 0);
        uint32_t acc = 0;
        for (int s = 0; s < K / 8; ++s) {
            const uint32_t b_off = static_cast<uint32_t>(s >> 2) * (N * 128u) + static_cast<uint32_t>(s & 3) * 32u;
            const uint64_t bh = tc::make_desc_sw128(smem_u32(b_hi) + b_off, 16, 1024);
            const uint64_t bl = tc::make_desc_sw128(smem_u32(b_lo) + b_off, 16, 1024);
            const uint32_t ah = tmem_base + a_col + static_cast<uint32_t>(s * 8), al = ah + static_cast<uint32_t>(K);
            tc::mma_tf32_ts(tmem_base, ah, bh, idesc, acc);
            if (split3) {
                tc::mma_tf32_ts(tmem_base, ah, bl, idesc, 1);
                tc::mma_tf32_ts(tmem_base, al, bh, idesc, 1);
            }
            acc = 1;
        }
        tc::mma_commit(&bar);
    }
    if (tid == 0 && mode != 2) {
        const uint32_t idesc = tc::make_idesc_tf32(M, N, mode, mode);
        const int ksteps = K / 8;
        uint32_t acc = 0;
        for (int s = 0; s < ksteps; ++s) {
            uint32_t a_off, b_off, a_lbo, b_lbo, sbo = mode == 0 ? 1024 : 512;
            const uint32_t layout = mode == 0 ? tc::kLayoutSw128 : tc::kLayoutSw128Base32;
            if (mode == 0) {
                a_off = static_cast<uint32_t>(s >> 2) * (M * 128u) + static_cast<uint32_t>(s & 3) * 32u;
                b_off = static_cast<uint32_t>(s >> 2) * (N * 128u) + static_cast<uint32_t>(s & 3) * 32u;
                a_lbo = b_lbo = 16;
            } else {
                a_off = b_off = static_cast<uint32_t>(s) * 1024u;
                a_lbo = b_lbo = static_cast<uint32_t>(K) * 128u;   // stride between 32-wide M/N column blocks
            }
            const int passes = split3 ? 3 : 1;
            for (int p = 0; p < passes; ++p) {
                const uint8_t *ap = (p == 2) ? a_lo : a_hi;       // hi*hi, hi*lo, lo*hi
                const uint8_t *bp = (p == 1) ? b_lo : b_hi;
                const uint64_t ad = tc::make_desc(smem_u32(ap) + a_off, a_lbo, sbo, layout);
                const uint64_t bd = tc::make_desc(smem_u32(bp) + b_off, b_lbo, sbo, layout);
                tc::mma_tf32(tmem_base, ad, bd, idesc, acc);
                acc = 1;
            }
        }
        tc::mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c0), v);
        const int row = warp * 32 + lane;
        for (int j = 0; j < 32; ++j)
            if (c0 + j < N) D[row * N + c0 + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, ncols);
}

}  // namespace ogc

extern "C" __attribute__((visibility("default"))) int ogc_tc_probe_gemm(int mode, int n, int k, int split3, const float *a, const float *b, float *d,
                                 void *stream) {
    using namespace ogc;
    if (n < 32 || n > 256 || n % 32 != 0 || k < 32 || k % 32 != 0 || !a || !b || !d) return OGC_ERR_INVALID_ARG;
    if (mode == 2 && n + 2 * k > 512) return OGC_ERR_UNSUPPORTED;
    const size_t smem = static_cast<size_t>(split3 ? 2 : 1) * (128 + n) * k * 4 + 1024;
    if (smem > static_cast<size_t>(kMaxSmemPerCta)) return OGC_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    tc_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(mode, n, k, split3, a, b, d);
    OGC_RETURN_LAUNCH_STATUS();
}
