// Diagnostic kernel: issue rate of tcgen05.mma kind::tf32 (M = 128, N = n, K = 8 per instruction) on a B200, for the
// operand sources the fused set-abstraction kernels use.  One thread per CTA issues `reps` x (K/8) k-steps back to back
// into one accumulator and the CTA reports the SM cycles between the first issue and the completion of the last MMA.
//   mode 0: A from tensor memory (.ts), 3xTF32 pattern per k-step: (a_hi, w_hi) (a_hi, w_lo) (a_lo, w_hi)
//   mode 1: A from shared memory (.ss), same pattern
//   mode 2: A from tensor memory, ONE MMA per k-step
//   mode 3: A from shared memory, ONE MMA per k-step
//   mode 4: as mode 0, alternating between two accumulators per k-step
//   mode 5: as mode 0 with the 8 k-steps' descriptors precomputed and the issue loop fully unrolled (K = 64 only)
//   mode 8: as mode 0 with a tcgen05.commit (to a second mbarrier) after every 4 k-steps (cost of frequent commits)
//   mode 6 / 7: as mode 0 while warps 1-3 stream tcgen05.ld / tcgen05.st on other tensor-memory columns (port contention)
// scratch/tc_rate.py prints the table; the measured numbers are quoted in DESIGN.md.
#include "../../ogc_b200/csrc/tcgen05.cuh"

namespace ogc {

__global__ void __launch_bounds__(128)
tc_rate_kernel(int mode, int N, int K, int reps, long long *__restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int done_flag;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t b_bytes = static_cast<uint32_t>(N) * K * 4, a_bytes = 128u * K * 4;
    uint8_t *b_hi = smem, *b_lo = b_hi + b_bytes, *a_hi = b_lo + b_bytes, *a_lo = a_hi + a_bytes;
    const bool ss = mode == 1 || mode == 3;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2, 1);
        mbar_fence_init();
        done_flag = 0;
    }
    // operand contents: small pseudo-random values (the data only matters for power)
    const uint32_t words = (2 * b_bytes + (ss ? 2 * a_bytes : 0)) / 4;
    for (uint32_t e = tid; e < words; e += blockDim.x) {
        uint32_t h = (e + 1u) * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15;
        reinterpret_cast<float *>(smem)[e] = static_cast<float>(h & 1023u) * (1.f / 512.f) - 1.f;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_col = 2u * N <= 512u - 2u * K ? 2u * N : static_cast<uint32_t>(N);   // TMEM column of A_hi; A_lo at +K
    if (!ss) {
        for (int k0 = 0; k0 < 2 * K; k0 += 32) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = static_cast<float>(((tid * 37 + k0 + j) * 2654435761u >> 22) & 1023u) * (1.f / 512.f) - 1.f;
            tc::tmem_st32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + a_col + static_cast<uint32_t>(k0), v);
        }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, N, 0, 0);
        const int ksteps = K / 8;
        const bool three = mode == 0 || mode == 1 || mode == 4 || mode == 5 || mode >= 6;
        if (mode == 5) {
            uint64_t bh[8], bl[8];
            uint32_t ah[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const uint32_t b_off = static_cast<uint32_t>(s >> 2) * (N * 128u) + static_cast<uint32_t>(s & 3) * 32u;
                bh[s] = tc::make_desc_sw128(smem_u32(b_hi) + b_off, 16, 1024);
                bl[s] = tc::make_desc_sw128(smem_u32(b_lo) + b_off, 16, 1024);
                ah[s] = tmem_base + a_col + static_cast<uint32_t>(s * 8);
            }
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    tc::mma_tf32_ts(tmem_base, ah[s], bh[s], idesc, (r | s) ? 1u : 0u);
                    tc::mma_tf32_ts(tmem_base, ah[s], bl[s], idesc, 1u);
                    tc::mma_tf32_ts(tmem_base, ah[s] + static_cast<uint32_t>(K), bh[s], idesc, 1u);
                }
            }
            tc::mma_commit(&bar);
            mbar_wait(&bar, 0);
            t1 = clock64();
            cycles[blockIdx.x] = t1 - t0;
        } else {
        t0 = clock64();
        uint32_t acc = 0;
        for (int r = 0; r < reps; ++r) {
            for (int s = 0; s < ksteps; ++s) {
                const uint32_t b_off = static_cast<uint32_t>(s >> 2) * (N * 128u) + static_cast<uint32_t>(s & 3) * 32u;
                const uint32_t a_off = static_cast<uint32_t>(s >> 2) * (128u * 128u) + static_cast<uint32_t>(s & 3) * 32u;
                const uint64_t bh = tc::make_desc_sw128(smem_u32(b_hi) + b_off, 16, 1024);
                const uint64_t bl = tc::make_desc_sw128(smem_u32(b_lo) + b_off, 16, 1024);
                const uint32_t d = tmem_base + ((mode == 4 && (s & 1)) ? static_cast<uint32_t>(N) : 0u);
                if (ss) {
                    const uint64_t ah = tc::make_desc_sw128(smem_u32(a_hi) + a_off, 16, 1024);
                    const uint64_t al = tc::make_desc_sw128(smem_u32(a_lo) + a_off, 16, 1024);
                    tc::mma_tf32(d, ah, bh, idesc, acc);
                    if (three) {
                        tc::mma_tf32(d, ah, bl, idesc, 1);
                        tc::mma_tf32(d, al, bh, idesc, 1);
                    }
                } else {
                    const uint32_t ah = tmem_base + a_col + static_cast<uint32_t>(s * 8), al = ah + static_cast<uint32_t>(K);
                    tc::mma_tf32_ts(d, ah, bh, idesc, acc);
                    if (three) {
                        tc::mma_tf32_ts(d, ah, bl, idesc, 1);
                        tc::mma_tf32_ts(d, al, bh, idesc, 1);
                    }
                }
                if (mode != 4 || (s & 1)) acc = 1;
                if (mode == 8 && (s & 3) == 3) tc::mma_commit(&bar2);
            }
        }
        tc::mma_commit(&bar);
        mbar_wait(&bar, 0);
        t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
        }
        done_flag = 1;
    } else if ((mode == 6 || mode == 7) && warp > 0) {
        // background tensor-memory traffic on the scratch columns [2N + 2K .. ) of this warp's lane quadrant
        const uint32_t scratch = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + a_col + 2u * K;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = static_cast<float>(tid + j);
        float sink = 0.f;
        while (!done_flag) {
            for (int it = 0; it < 16; ++it) {
                if (mode == 6) { float w[32]; tc::tmem_ld32(scratch, w); sink += w[it & 31]; }
                else tc::tmem_st32(scratch, v);
            }
        }
        if (sink == 12345.678f) cycles[0] = 0;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace ogc

// cycles: one int64 per CTA.  Returns OGC_ERR_UNSUPPORTED when the operands do not fit.
extern "C" __attribute__((visibility("default"))) int ogc_tc_rate(int mode, int n, int k, int reps, int ctas,
                                                                   long long *cycles, void *stream) {
    using namespace ogc;
    if (n < 16 || n > 256 || n % 16 != 0 || k < 32 || k % 32 != 0 || reps < 1 || ctas < 1 || !cycles) return OGC_ERR_INVALID_ARG;
    const bool ss = mode == 1 || mode == 3;
    if (!ss && (mode == 4 ? 2 * n : n) + 2 * k > 512) return OGC_ERR_UNSUPPORTED;
    if ((mode == 6 || mode == 7) && 2 * n + 2 * k + 32 > 512) return OGC_ERR_UNSUPPORTED;
    if (mode == 5 && k != 64) return OGC_ERR_UNSUPPORTED;
    const size_t smem = static_cast<size_t>(2) * n * k * 4 + (ss ? static_cast<size_t>(2) * 128 * k * 4 : 0) + 1024;
    if (smem > static_cast<size_t>(kMaxSmemPerCta)) return OGC_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(tc_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    tc_rate_kernel<<<ctas, 128, smem, static_cast<cudaStream_t>(stream)>>>(mode, n, k, reps, cycles);
    OGC_RETURN_LAUNCH_STATUS();
}
