// What does one kernel launch cost when it is timed the way bench.py times a FUSED kernel (CUDA events between launches on one
// stream), and which property of the fit / post launches makes it more expensive than the reprojection's?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bmfr_b200/csrc launch_overhead.cu -o launch_overhead
// Kernels that do nothing but (optionally) what a CTA of the real kernels does before its first useful instruction:
//   small    : 8 bytes of parameters, 2040 CTAs x 256 threads
//   params   : + 1.5 KB of __grid_constant__ parameters (KParams + six tensor maps)
//   smem     : + 60 KB of dynamic shared memory, 444 CTAs
//   tma1/6   : + thread 0 of every CTA requests one / six 34-row boxes by TMA and every thread waits for them
// Printed: microseconds per launch between events, and back to back (1000 launches between one pair of events).
#include <cstdio>
#include <cuda_runtime.h>

#include "bmfr_tma.cuh"

struct Big {
    char pad[760];
    CUtensorMap map[6];
};

__global__ void k_small(int* out) {
    if (out != nullptr && threadIdx.x == 0 && blockIdx.x == 0x7fffffff) *out = 1;
}
__global__ void k_params(const __grid_constant__ Big b, int* out) {
    if (out != nullptr && threadIdx.x == 0 && blockIdx.x == 0x7fffffff) *out = b.pad[3];
}
template <int NMAPS>
__global__ void __launch_bounds__(256, 3) k_smem(const __grid_constant__ Big b, int* out, int rows) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 59 * 1024);
    if (NMAPS > 0) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
            mbar_expect_tx(bar, NMAPS * 34 * 96 * 4);
            for (int i = 0; i < NMAPS; ++i) tma_load_tile(smem + i * 34 * 96 * 4, &b.map[i], 96 * (blockIdx.x % 8), (blockIdx.x * 7) % (rows - 34), bar);
        }
        __syncthreads();
        mbar_wait_hot(bar, 0);
    }
    if (out != nullptr && threadIdx.x == 0 && blockIdx.x == 0x7fffffff) *out = smem[5];
}

template <class F>
static void time_it(const char* name, F launch) {
    cudaEvent_t e[2];
    cudaEventCreate(&e[0]);
    cudaEventCreate(&e[1]);
    for (int i = 0; i < 20; ++i) launch();
    cudaDeviceSynchronize();
    float sep = 0.f;
    const int n = 300;
    for (int i = 0; i < n; ++i) {
        cudaEventRecord(e[0]);
        launch();
        cudaEventRecord(e[1]);
        cudaEventSynchronize(e[1]);
        float ms;
        cudaEventElapsedTime(&ms, e[0], e[1]);
        sep += ms;
    }
    cudaEventRecord(e[0]);
    for (int i = 0; i < 1000; ++i) launch();
    cudaEventRecord(e[1]);
    cudaEventSynchronize(e[1]);
    float b2b;
    cudaEventElapsedTime(&b2b, e[0], e[1]);
    printf("%-10s between events %6.2f us   back to back %6.2f us   (%s)\n", name, 1e3f * sep / n, b2b, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int W = 1920 * 3, rows = 1080;
    float* img[6];
    Big b = {};
    for (int i = 0; i < 6; ++i) {
        cudaMalloc(&img[i], sizeof(float) * W * rows);
        cudaMemset(img[i], 0, sizeof(float) * W * rows);
        if (!bmfr_tensor_map_2d(img[i], 4, W, rows, 96, 34, &b.map[i])) { printf("tensor map failed\n"); return 1; }
    }
    cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    cudaFuncSetAttribute(k_smem<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    time_it("small", [&] { k_small<<<2040, 256>>>(nullptr); });
    time_it("params", [&] { k_params<<<2040, 256>>>(b, nullptr); });
    time_it("smem", [&] { k_smem<0><<<444, 256, 60 * 1024>>>(b, nullptr, rows); });
    time_it("smem2135", [&] { k_smem<0><<<2135, 256, 60 * 1024>>>(b, nullptr, rows); });
    time_it("tma1", [&] { k_smem<1><<<444, 256, 60 * 1024>>>(b, nullptr, rows); });
    time_it("tma6", [&] { k_smem<6><<<444, 256, 60 * 1024>>>(b, nullptr, rows); });
    time_it("tma6x2135", [&] { k_smem<6><<<2135, 256, 60 * 1024>>>(b, nullptr, rows); });
    // alternating carve-outs, as the frame loop does (reprojection: no shared memory; fit / post: 3 x 60 KB)
    time_it("alternate", [&] { k_small<<<2040, 256>>>(nullptr); k_smem<0><<<444, 256, 60 * 1024>>>(b, nullptr, rows); });
    return 0;
}
