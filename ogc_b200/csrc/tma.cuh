// Tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG) helpers: 2-D fp32 tiles global -> shared with mbarrier completion.
// One instruction moves a whole [box_outer][box_inner] tile; a 1-D cp.async.bulk (UBLKCP) of one 512-byte row costs the
// copy engine about as much as a whole tile (measured: ~50-70 cycles per operation per SM), which capped row-by-row
// staging at ~9 B/clk/SM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace ogc {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The driver entry point is resolved through the runtime (no link-time dependency on libcuda).
static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Row-major fp32 matrix (outer rows of `inner` elements, contiguous), tiles of [box_outer][box_inner], no swizzle:
// the tile lands in shared memory row-major, box_inner * 4 bytes per row; with swizzle 1 (box_inner * 4 == 128, destination
// 1024-byte aligned) in the 128-byte-swizzled K-major layout a tcgen05 shared-memory descriptor reads directly, with
// swizzle 2 in the 32-byte-atom variant that MN-major tf32 operands use (UMMA SWIZZLE_128B_BASE32B).  inner * 4 must be a multiple of 16 and
// `base` 16-byte aligned.  Returns false when the driver refuses.
static inline bool make_2d_f32(CUtensorMap *map, const void *base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                               uint32_t box_outer, int swizzle = 0) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {inner, outer};
    const cuuint64_t strides[1] = {inner * 4};
    const cuuint32_t box[2] = {box_inner, box_outer};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE,
              swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
// One thread: tile whose first element is (row y, column x) -> dst (128-byte aligned); completes `bar` with the tile's bytes.
__device__ __forceinline__ void load_2d(void *dst_smem, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem))),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
        : "memory");
}
// One thread: shared-memory tile (layout per the map's swizzle) -> global tile at (row y, column x); completion is tracked
// by the thread's bulk async-group (store_commit / store_wait_read).
__device__ __forceinline__ void store_2d(const CUtensorMap *map, int x, int y, const void *src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(x), "r"(y), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(src_smem)))
                 : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void store_wait_read() {      // at most N of this thread's groups still READING shared memory
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
#endif

}  // namespace tma
}  // namespace ogc
