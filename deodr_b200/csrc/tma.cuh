// TMA (tensor memory accelerator) helpers for the framebuffer tiles: 2-D tensor maps over the caller's image-shaped
// buffers, tile loads into shared memory (cp.async.bulk.tensor -> SASS UTMALDG) and tile stores from it (UTMASTG).
//
// A 16x16 pixel tile of an interleaved [H, W, C] fp32 image is a box of 16 rows x 16*C floats of the 2-D tensor
// [H, W*C]; of the int32 owner map / the fp64 z-buffer, a 16 x 16 box of [H, W].  One elected thread issues the copy,
// the hardware clips boxes that stick out of the image (loads fill the outside with zeros, stores drop it), and the
// 256 threads of the CTA are spared the strided 4-byte accesses (C = 3: three store instructions of 4 bytes at a 12-byte
// stride per pixel).  Requirements of the hardware, checked at encode time: base address and row pitch multiples of 16
// bytes (C = 3 needs W % 4 == 0), inner box <= 256 elements; a buffer that does not qualify keeps the plain path.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace deodr {

struct TileMap {
    alignas(64) CUtensorMap map;
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// 2-D map over `rows` x `row_elems` elements of `elem_bytes` (4: fp32 / int32 as UINT32-compatible, 8: fp64 as UINT64),
// box = box_rows x box_elems.  false: the buffer does not qualify (the caller keeps its plain loads / stores).
static inline bool encode_tile_map(TileMap *out, const void *base, int elem_bytes, bool is_float, int rows, int row_elems,
                                   int box_rows, int box_elems) {
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode || !base) return false;
    const size_t pitch = (size_t)row_elems * elem_bytes;
    if (((uintptr_t)base & 15) || (pitch & 15) || box_elems > 256 || box_rows > 256 || ((size_t)box_elems * elem_bytes & 15))
        return false;
    const CUtensorMapDataType type = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64
                                     : is_float      ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                     : CU_TENSOR_MAP_DATA_TYPE_INT32;
    const cuuint64_t dims[2] = {(cuuint64_t)row_elems, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch};  // byte stride of dimension 1 (dimension 0 is dense)
    const cuuint32_t box[2] = {(cuuint32_t)box_elems, (cuuint32_t)box_rows};
    const cuuint32_t elem_strides[2] = {1, 1};
    return encode(&out->map, type, 2, const_cast<void *>(base), dims, strides, box, elem_strides,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#if defined(__CUDACC__)
static __device__ __forceinline__ uint32_t tma_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tile load: global (tensor map, element coordinate x of dimension 0, row y) -> shared; completion on `bar` (tx bytes)
static __device__ __forceinline__ void tma_load_tile(void *dst, const TileMap *map, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            tma_smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(tma_smem_u32(bar))
        : "memory");
}

// tile store: shared -> global; the caller orders its shared-memory writes before it with fence_proxy_async() and a
// CTA barrier, and keeps the shared memory alive until tma_store_wait_read()
static __device__ __forceinline__ void tma_store_tile(const TileMap *map, int x, int y, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x),
                 "r"(y), "r"(tma_smem_u32(src))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
static __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// makes the generic-proxy writes of this thread to shared memory visible to the async proxy (the TMA engine)
static __device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

}  // namespace deodr
