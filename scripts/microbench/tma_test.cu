// Stand-alone check of the 2-D TMA tile load used by fit_qr_kernel (tensor map inside a __grid_constant__ struct).
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

struct Maps { CUtensorMap a, b; int flag; };
struct Dummy { int x[88]; };   // 352 bytes, like KParams

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ Dummy D, const __grid_constant__ Maps M, float* out, int c0, int c1) {
    extern __shared__ __align__(128) unsigned char raw[];
    float* tile = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tile + 32 * 100);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(32 * 100 * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_u32(tile)),
                     "l"(&M.b), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                     : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nL_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra L_DONE;\nbra L_WAIT;\nL_DONE:\n}\n" ::"r"(
            smem_u32(bar))
        : "memory");
    for (int i = threadIdx.x; i < 32 * 100; i += blockDim.x) out[i] = tile[i] + D.x[0];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int W = 160, rows = 96;
    std::vector<float> h((size_t)rows * W * 3);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, 32 * 100 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    Maps M;
    memset(&M, 0, sizeof(M));
    const cuuint64_t dims[2] = {(cuuint64_t)W * 3, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 3 * sizeof(float)};
    const cuuint32_t box[2] = {100, 32}, elem[2] = {1, 1};
    CUresult r = enc(&M.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d  sizeof(Maps)=%zu alignof=%zu\n", (int)r, sizeof(Maps), alignof(Maps));
    M.a = M.b;
    Dummy D;
    memset(&D, 0, sizeof(D));
    const int smem = 32 * 100 * 4 + 8 + 128;
    for (int c0 : {0, 6, 30, 54, 384}) {
        const int c1 = 16;
        const int c0a = c0 & ~3;
        k<<<1, 128, smem>>>(D, M, out, c0a, c1);
        e = cudaDeviceSynchronize();
        std::vector<float> o(32 * 100);
        cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < 32; ++y)
            for (int x = 0; x < 100 && c0a + x < W * 3; ++x)
                if (o[y * 100 + x] != h[(size_t)(c1 + y) * W * 3 + c0a + x]) ++bad;
        printf("c0=%d: %s, mismatches %d\n", c0, cudaGetErrorString(e), bad);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
