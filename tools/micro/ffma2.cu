// micro-benchmark: issue rate and dependent latency of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int CHAINS, bool PACKED>
__global__ void k(float* out, float a, float b) {
    float x[CHAINS * 2];
    unsigned long long X[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS * 2; ++c) x[c] = threadIdx.x * 1e-3f + c;
    unsigned long long A, B;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b), "f"(b));
    if (PACKED) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) asm("mov.b64 %0, {%1, %2};" : "=l"(X[c]) : "f"(x[2 * c]), "f"(x[2 * c + 1]));
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(X[c]) : "l"(X[c]), "l"(A), "l"(B));
        }
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * c]), "=f"(x[2 * c + 1]) : "l"(X[c]));
    } else {
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int c = 0; c < CHAINS * 2; ++c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[c]) : "f"(x[c]), "f"(a), "f"(b));
        }
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS * 2; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CHAINS, bool PACKED>
void run(const char* name, int blocks, int threads) {
    float* out; cudaMalloc(&out, blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CHAINS, PACKED><<<blocks, threads>>>(out, 0.999f, 1e-3f);
    cudaEventRecord(e0);
    k<CHAINS, PACKED><<<blocks, threads>>>(out, 0.999f, 1e-3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * CHAINS * 2 * ITERS * (double)blocks * threads;
    double inst = (PACKED ? 1.0 : 2.0) * CHAINS * ITERS * (double)blocks * threads / 32;
    printf("%-28s chains=%d blocks=%d threads=%d  %.3f ms  %.1f TFLOP/s  %.2f warp-inst/clk/SMSP (at 1.9 GHz, 148 SMs)\n", name, CHAINS, blocks,
           threads, ms, flops / ms * 1e-9, inst / (ms * 1e-3) / 1.9e9 / (148 * 4));
    cudaFree(out);
}
int main() {
    // throughput: many warps, many independent chains
    run<8, false>("FFMA  throughput", 148 * 8, 256);
    run<8, true>("FFMA2 throughput", 148 * 8, 256);
    // latency: one warp per SMSP, one dependent chain
    run<1, false>("FFMA  2 chains, 1 warp/SMSP", 148, 128);
    run<1, true>("FFMA2 1 chain,  1 warp/SMSP", 148, 128);
    run<2, true>("FFMA2 2 chains, 1 warp/SMSP", 148, 128);
    run<4, false>("FFMA  8 chains, 1 warp/SMSP", 148, 128);
    run<4, true>("FFMA2 4 chains, 1 warp/SMSP", 148, 128);
    return 0;
}
