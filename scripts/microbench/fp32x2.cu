// Issue-rate microbenchmark: scalar FFMA / FADD / FMUL against the packed f32x2 forms on sm_100a.
// Prints warp-instructions per clock per SM sub-partition (SMSP) and the implied FMA lanes.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, float seed, long long* clocks) {
    float a[CHAINS], b2[CHAINS];
    unsigned long long p[CHAINS];
    const float m = seed + 1.0f, c = seed * 0.5f;
    for (int i = 0; i < CHAINS; ++i) { a[i] = threadIdx.x * 0.001f + i; b2[i] = a[i] + 1.f; }
    for (int i = 0; i < CHAINS; ++i) {
        float2 v = make_float2(a[i], b2[i]);
        p[i] = *reinterpret_cast<unsigned long long*>(&v);
    }
    float2 mm = make_float2(m, m), cc = make_float2(c, c);
    const unsigned long long m2 = *reinterpret_cast<unsigned long long*>(&mm), c2 = *reinterpret_cast<unsigned long long*>(&cc);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], m, c);
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
            if (MODE == 2) a[i] = __fadd_rn(a[i], c);
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
            if (MODE == 4) a[i] = __fmul_rn(a[i], m);
            if (MODE == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(m2));
            if (MODE == 6) { a[i] = fmaf(a[i], m, c); b2[i] = __fadd_rn(b2[i], c); }   // FFMA + FADD mix
            if (MODE == 7) { a[i] = fmaf(a[i], m, c); b2[i] = fminf(b2[i], a[i]); }    // FFMA + FMNMX (alu pipe)
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < CHAINS; ++i) {
        float2 v = *reinterpret_cast<float2*>(&p[i]);
        s += a[i] + b2[i] + v.x + v.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, float* out, long long* clk) {
    const int blocks = 148 * 2;
    k<MODE><<<blocks, 1024>>>(out, 0.25f, clk);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 1024>>>(out, 0.25f, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[8]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    // per CTA: 32 warps over 4 SMSPs = 8 warps per SMSP (2 CTAs per SM resident -> 16)
    const double inst_per_warp = (double)ITERS * CHAINS * per_iter;
    const double cyc = (double)h[0];
    // with 2 resident CTAs per SM an SMSP holds 16 warps; all run concurrently for ~cyc cycles
    printf("%-22s %8.3f ms  cta cycles %9.0f  warp-inst/clk/SMSP %.3f\n", name, ms, cyc, 16.0 * inst_per_warp / cyc);
}

int main() {
    float* out; long long* clk;
    cudaMalloc(&out, 148 * 2 * 1024 * sizeof(float));
    cudaMalloc(&clk, 148 * 2 * sizeof(long long));
    run<0>("FFMA", 1, out, clk);
    run<1>("FFMA2 (f32x2)", 1, out, clk);
    run<2>("FADD", 1, out, clk);
    run<3>("FADD2 (f32x2)", 1, out, clk);
    run<4>("FMUL", 1, out, clk);
    run<5>("FMUL2 (f32x2)", 1, out, clk);
    run<6>("FFMA+FADD", 2, out, clk);
    run<7>("FFMA+FMNMX", 2, out, clk);
    cudaError_t e = cudaGetLastError();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
