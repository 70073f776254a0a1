// issue_probe.cu -- issue-slot / pipe-rate probe for sm_100a.
// Per loop trip and per chain: 1 DFMA + KA x LOP3 (alu pipe) + KI x IMAD
// (fma pipe) + KW x IMAD.WIDE, all independent across 8 chains.
#include <cstdio>
#include <cuda_runtime.h>

template <int KD, int KA, int KI, int KW>
__global__ void __launch_bounds__(256) probe(long long iters, double* sink, unsigned* isink) {
    double a[8];
    unsigned u[8], v[8];
    unsigned long long w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-9 + i; u[i] = threadIdx.x * 7 + i; v[i] = threadIdx.x * 13 + i; w[i] = threadIdx.x * 17 + i; }
    double m = 1.0000001 + threadIdx.x * 1e-12, c = 1e-9;
    unsigned k1 = 0x9E3779B9u + threadIdx.x, k2 = 0xD2511F53u;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int k = 0; k < KD; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
#pragma unroll
            for (int k = 0; k < KA; ++k) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(k1), "r"(v[i]));
#pragma unroll
            for (int k = 0; k < KI; ++k) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(k2), "r"(k1));
#pragma unroll
            for (int k = 0; k < KW; ++k) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(v[i]), "r"(k2));
        }
    }
    double s = 0; unsigned x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += a[i]; x ^= u[i] ^ v[i] ^ (unsigned)w[i] ^ (unsigned)(w[i] >> 32); }
    if (s == 12345.678) sink[0] = s;
    if (x == 0x12345678u) isink[0] = x;
}

static double base_ms = 0;

template <int KD, int KA, int KI, int KW>
void run(int sm) {
    double* sink; unsigned* isink;
    cudaMalloc(&sink, 8); cudaMalloc(&isink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    long long iters = 20000;
    probe<KD, KA, KI, KW><<<sm * 8, 256>>>(iters / 10, sink, isink);
    cudaEventRecord(e0);
    probe<KD, KA, KI, KW><<<sm * 8, 256>>>(iters, sink, isink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (KD == 1 && KA + KI + KW == 0) base_ms = ms;
    // cycles per chain-trip per SMSP, calibrated on DFMA-only = 2 cycles
    printf("dfma=%d lop3=%d imad=%d imadwide=%d : %8.3f ms  = %5.2f cycles/SMSP per trip\n",
           KD, KA, KI, KW, ms, base_ms > 0 ? 2.0 * ms / base_ms : 0.0);
}

int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    run<1, 0, 0, 0>(sm); run<1, 0, 0, 0>(sm);
    run<0, 0, 0, 1>(sm); run<0, 0, 0, 2>(sm);
    run<0, 1, 0, 1>(sm); run<0, 2, 0, 2>(sm);
    run<1, 0, 0, 1>(sm); run<2, 0, 0, 1>(sm); run<3, 0, 0, 1>(sm); run<1, 0, 0, 2>(sm);
    run<3, 1, 0, 1>(sm); run<6, 2, 0, 2>(sm); run<3, 1, 1, 1>(sm); run<3, 2, 0, 1>(sm);
    run<3, 0, 0, 0>(sm); run<3, 1, 0, 0>(sm); run<3, 2, 0, 0>(sm); run<3, 3, 0, 0>(sm);
    return 0;
}
