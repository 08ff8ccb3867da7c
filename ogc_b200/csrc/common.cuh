// Shared device/host helpers for libogc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ogc_b200.h"

#define OGC_FULL_MASK 0xffffffffu

#define OGC_RETURN_LAUNCH_STATUS()                  \
    do {                                            \
        cudaError_t e__ = cudaGetLastError();       \
        return e__ == cudaSuccess ? OGC_OK : (int)e__; \
    } while (0)

namespace ogc {

constexpr int kNumSMs = 148;              // B200
constexpr int kMaxSmemPerCta = 227 * 1024;

// fp32 squared distance in the rounding order of the reference build:
//   nvcc contracts (a-b)*(a-b) + (c-d)*(c-d) + (e-f)*(e-f) into fma(dz,dz, fma(dx,dx, dy*dy))
// (pointnet2/src/sampling_gpu.cu:133, interpolate_gpu.cu:40,108, ball_query_gpu.cu:33).
// Spelled with explicit intrinsics so no compiler flag can change it.
__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Packed fp32 FMA (Blackwell FFMA2): d = a * b + c on both halves, each half an IEEE fp32 fma.  A 3-register FFMA
// issues every other cycle per scheduler; FFMA2 retires two FMAs per issue slot, which is what lets the fp32 SIMT
// kernels approach the 128 FMA/clk/SM peak.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- explicit shared-state-space accesses on 32-bit addresses (smem_u32): one LDS/STS with an immediate offset
// instead of a generic LD/ST with 64-bit address arithmetic when the compiler cannot prove the address space ----
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds_v2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async4_s(uint32_t dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) -------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- per-thread async copies global -> shared (SASS: LDGSTS), tracked with commit / wait groups ----
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` (0..7) of this thread's most recent groups are still in flight
__device__ __forceinline__ void cp_async_wait(int pending) {
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}

// Stage `nfloats` contiguous floats from global memory into shared memory with the TMA bulk-copy
// engine.  `sm` must be 16-byte aligned with room for nfloats + 4 floats.  The source only needs
// 4-byte alignment: the (<= 3 float) unaligned head and tail are moved with ordinary loads, the
// 16-byte aligned interior with one cp.async.bulk per <= 2^19-byte chunk.  Returns the shared
// address of element 0.  Must be called by ALL threads of the CTA.  It contains the mbarrier wait
// but no __syncthreads: the caller issues one __syncthreads() after it (head/tail visibility) and
// one before the buffer is staged again.  `parity` is the mbarrier
// phase to wait for and is flipped on return.
__device__ __forceinline__ const float *stage_floats(float *sm, const float *__restrict__ src, int nfloats,
                                                     uint64_t *bar, uint32_t &parity) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
    const int shift = static_cast<int>((addr & 15u) >> 2);        // floats past a 16 B boundary
    float *dst = sm + shift;                                      // keeps src/dst congruent mod 16 B
    const int head = shift == 0 ? 0 : min(4 - shift, nfloats);    // floats before the first boundary
    const int body = ((nfloats - head) >> 2) << 2;                // multiple of 4 floats = 16 B
    const int tail = nfloats - head - body;
    if (threadIdx.x == 0 && body > 0) {
        const uint32_t bytes = static_cast<uint32_t>(body) * 4u;
        mbar_arrive_expect_tx(bar, bytes);
        uint32_t off = 0;
        while (off < bytes) {
            const uint32_t chunk = min(bytes - off, 1u << 19);
            bulk_g2s(reinterpret_cast<char *>(dst + head) + off, reinterpret_cast<const char *>(src + head) + off,
                     chunk, bar);
            off += chunk;
        }
    }
    for (int i = threadIdx.x; i < head + tail; i += blockDim.x) {
        const int e = i < head ? i : body + i;   // head element i, or tail element (i - head)
        dst[e] = __ldg(src + e);
    }
    if (body > 0) {
        mbar_wait(bar, parity);
        parity ^= 1u;
    }
    return dst;
}

// v[c] (c < 32) per lane  ->  returns sum over the 32 lanes of v[lane]  (31 shuffles instead of 160)
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const bool up = lane & 16;
        const float send = up ? v[i] : v[i + 16];
        const float keep = up ? v[i + 16] : v[i];
        v[i] = keep + __shfl_xor_sync(OGC_FULL_MASK, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool up = lane & 8;
        const float send = up ? v[i] : v[i + 8];
        const float keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(OGC_FULL_MASK, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 4;
        const float send = up ? v[i] : v[i + 4];
        const float keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(OGC_FULL_MASK, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 2;
        const float send = up ? v[i] : v[i + 2];
        const float keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(OGC_FULL_MASK, send, 2);
    }
    {
        const bool up = lane & 1;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(OGC_FULL_MASK, send, 1);
    }
    return v[0];
}

}  // namespace ogc
