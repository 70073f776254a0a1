// rf_probe.cu -- register-file operand bandwidth probe for FP64 on sm_100a:
// DFMA with 1, 2 or 3 DISTINCT 64-bit register operands per instruction.
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: a = fma(a, m, c)        m, c shared by all chains (reuse cache friendly)
// MODE 1: a = fma(a, b_i, c)      two distinct register operands
// MODE 2: a = fma(a, b_i, c_i)    three distinct register operands
// MODE 3: a = a * b_i  (DMUL, 2 distinct)      MODE 4: a = a + b_i (DADD, 2 distinct)
// MODE 5: a = fma(a, K, c_i) with K a compile-time constant (constant-bank / immediate operand)
template <int MODE>
__global__ void __launch_bounds__(256) probe(long long iters, double* sink) {
    double a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-9 + i; b[i] = 1.0 + (threadIdx.x + i) * 1e-12; c[i] = 1e-9 * (i + 1 + threadIdx.x); }
    double m = 1.0000001 + threadIdx.x * 1e-12, cc = 1e-9 + threadIdx.x * 1e-15;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(cc));
            if (MODE == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(cc));
            if (MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(c[i]));
            if (MODE == 3) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[i]));
            if (MODE == 4) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[i]));
            if (MODE == 5) asm volatile("fma.rn.f64 %0, %0, 0d3FF000001AD7F29B, %1;" : "+d"(a[i]) : "d"(c[i]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i] + c[i];
    if (s == 12345.678) sink[0] = s;
}

static double base_ms = 0;
template <int MODE> void run(int sm, const char* name) {
    double* sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    long long iters = 20000;
    probe<MODE><<<sm * 8, 256>>>(iters / 10, sink);
    cudaEventRecord(e0);
    probe<MODE><<<sm * 8, 256>>>(iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (MODE == 0) base_ms = ms;
    printf("%-46s %8.3f ms  = %5.2f cycles/SMSP per instr\n", name, ms, 2.0 * ms / base_ms);
}
int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    run<0>(sm, "DFMA a,a,m,c   (shared m,c)");
    run<1>(sm, "DFMA a,a,b_i,c (2 distinct regs)");
    run<2>(sm, "DFMA a,a,b_i,c_i (3 distinct regs)");
    run<3>(sm, "DMUL a,a,b_i");
    run<4>(sm, "DADD a,a,b_i");
    run<5>(sm, "DFMA a,a,CONST,c_i");
    return 0;
}
