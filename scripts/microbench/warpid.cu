// Which hardware warp slots (and so which SM sub-partitions, slot % 4) do the five warps of the 160-thread
// fit CTAs get when three CTAs share an SM?   nvcc -arch=sm_100a -o warpid warpid.cu && ./warpid
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(160, 3) probe(int* out) {
    extern __shared__ char smem[];
    unsigned sm, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if ((threadIdx.x & 31) == 0) {
        int* o = out + (blockIdx.x * 5 + (threadIdx.x >> 5)) * 2;
        o[0] = (int)sm; o[1] = (int)wid;
    }
    // keep the CTA resident long enough that all three CTAs of an SM coexist
    long long t0 = clock64();
    while (clock64() - t0 < 2000000) {}
    if (smem[threadIdx.x] == 77) out[0] = 1;
}
int main() {
    int n = 148 * 3;
    int* d; cudaMalloc(&d, n * 10 * sizeof(int));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    probe<<<n, 160, 60000>>>(d);
    int* h = new int[n * 10];
    cudaMemcpy(h, d, n * 10 * sizeof(int), cudaMemcpyDeviceToHost);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    for (int sm = 0; sm < 3; ++sm) {
        printf("sm %d:", sm);
        for (int b = 0; b < n; ++b)
            if (h[b * 10] == sm) { printf("  cta %d slots", b); for (int w = 0; w < 5; ++w) printf(" %d", h[b * 10 + w * 2 + 1]); }
        printf("\n");
    }
    int hist[4] = {0, 0, 0, 0};
    for (int b = 0; b < n; ++b) hist[h[b * 10 + 4 * 2 + 1] & 3]++;
    printf("solver-warp slot %% 4 histogram: %d %d %d %d\n", hist[0], hist[1], hist[2], hist[3]);
    return 0;
}
