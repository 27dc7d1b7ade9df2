// TMA (cp.async.bulk.tensor) + mbarrier helpers shared by the FUSED kernels, and the host-side cache of
// 2-D tensor maps (the encoder is fetched through cudaGetDriverEntryPoint: no link against libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <unordered_map>

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// Spin on an mbarrier phase.  No nanosleep back-off: the wait (a tile's inputs) is on the critical path of
// every warp, try_wait already suspends the thread for a hardware-defined interval, and a sleeping warp
// can oversleep the arrival by microseconds.  The loop stays inside one asm block so that the compiler
// sees straight-line code and keeps treating the warp as converged for the shuffles that follow.
__device__ __forceinline__ void mbar_wait_hot(unsigned long long* b, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_HOT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_HOT_DONE;\n"
        "bra MBAR_HOT_LOOP;\n"
        "MBAR_HOT_DONE:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// One 2-D box global -> shared; coordinates in elements (innermost first), out-of-bounds elements arrive as zeros.
__device__ __forceinline__ void tma_load_tile(void* dst_smem, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tma_load_tile_hint(void* dst_smem, const CUtensorMap* map, int c0, int c1, unsigned long long* bar,
                                                   unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*BmfrEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline BmfrEncodeTiledFn bmfr_encode_tiled_fn() {
    static BmfrEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (BmfrEncodeTiledFn)p;
    }
    return fn;
}

// A [rows][row_elems] tensor of `elem_bytes`-sized elements (4: float, 1: byte), row pitch = row_elems * elem_bytes, boxes of
// box_w x box_h elements.  The TMA wants a 16-byte aligned base, a pitch that is a multiple of 16 bytes and a box row of a
// multiple of 16 bytes.  Maps are cached by (pointer, shape, box): a caller that cycles through a fixed set of frame buffers
// encodes each of them once.  (One cache per translation unit: the function is static.)
static bool bmfr_tensor_map_2d(const void* base, int elem_bytes, long long row_elems, int rows, int box_w, int box_h, CUtensorMap* out) {
    struct Key {
        const void* p; long long w; int r, e, bw, bh;
        bool operator==(const Key& o) const { return p == o.p && w == o.w && r == o.r && e == o.e && bw == o.bw && bh == o.bh; }
    };
    struct Hash {
        size_t operator()(const Key& k) const {
            return std::hash<const void*>()(k.p) ^ ((size_t)k.w * 1315423911u) ^ ((size_t)k.r << 20) ^ ((size_t)k.bw << 7) ^ ((size_t)k.bh << 13) ^ (size_t)k.e;
        }
    };
    static std::unordered_map<Key, CUtensorMap, Hash> cache;
    static std::mutex mu;
    BmfrEncodeTiledFn enc = bmfr_encode_tiled_fn();
    if (!enc || !base || ((uintptr_t)base & 15) != 0 || ((row_elems * elem_bytes) & 15) != 0 || ((box_w * elem_bytes) & 15) != 0 || box_w > 256 ||
        box_h > 256 || rows < 1)
        return false;
    std::lock_guard<std::mutex> lock(mu);
    const Key key{base, row_elems, rows, elem_bytes, box_w, box_h};
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return true;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)row_elems, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_elems * (cuuint64_t)elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, elem[2] = {1, 1};
    CUtensorMap m;
    if (enc(&m, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box,
            elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return true;
}
