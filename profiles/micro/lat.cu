// Dependent-chain latencies on sm_100a (one warp): nvcc -arch=sm_100a -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int OP>
__global__ void k(float *out, long long *clk, float a, float b) {
    float x = a + threadIdx.x * 1e-3f, y = b;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = x + y;                                   // FADD
        if (OP == 1) x = fmaxf(x, y) + 1e-7f;                     // FMNMX + FADD
        if (OP == 2) x = ex2f(x) * 0.5f;                          // EX2 + FMUL
        if (OP == 3) x = lg2f(x) + 2.f;                           // LG2 + FADD
        if (OP == 4) x = __shfl_down_sync(0xffffffffu, x, 1) + y; // SHFL + FADD
        if (OP == 5) { float m = fmaxf(x, y), d = fminf(x, y) - m; x = m + lg2f(1.f + ex2f(d)); }   // logadd2
        if (OP == 6) { float u = __shfl_down_sync(0xffffffffu, x, 1); float p = x + y, q = u + y; float m = fmaxf(p, q), d = fminf(p, q) - m; x = m + lg2f(1.f + ex2f(d)) - 1.f; }  // full chain step
        if (OP == 7) { float u = __shfl_down_sync(0xffffffffu, x, 1); x = fmaf(x, y, u * y) * 1.0001f; }   // linear-domain step
        if (OP == 8) x = fmaf(x, y, y);                           // FFMA
        if (OP == 9) { float r; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(x)); x = r + y; }  // CREDUX + FADD
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
    float *out; long long *clk, h;
    cudaMalloc(&out, 4096); cudaMalloc(&clk, 8);
    const char *names[] = {"FADD", "FMNMX+FADD", "EX2+FMUL", "LG2+FADD", "SHFL+FADD", "logadd2 (FMNMX,FMNMX,FADD,EX2,FADD,LG2,FADD)",
                           "chain step (SHFL,2 FADD,logadd2,FADD)", "linear step (SHFL,FMUL,FFMA,FMUL)", "FFMA", "CREDUX+FADD"};
#define RUN(OP) k<OP><<<1, 32>>>(out, clk, 0.5f, -0.25f); k<OP><<<1, 32>>>(out, clk, 0.5f, -0.25f); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("%-50s %.1f clk per iteration\n", names[OP], (double)h / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    return 0;
}
