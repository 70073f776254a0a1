// mix_probe.cu -- does the lean Heston kernel's instruction mix overlap its FP64 and
// integer work at the kernel's own occupancy (2 x 256-thread CTAs per SM, CH
// independent chains per thread)?  Per trip: K2 two-register FP64 instructions (DMUL
// a, a, b_i), K3 three-register DFMAs (a, b_i, c_i), KA LOP3, KI IMAD, KW IMAD.WIDE,
// spread evenly over the trip and round-robin over the chains.  MODE 0 mixed,
// 1 FP64 only, 2 integer only.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mix_probe tools/mix_probe.cu && /tmp/mix_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, int K2, int K3, int KA, int KI, int KW, int MODE>
__global__ void __launch_bounds__(256, 2) probe(long long iters, double* sink, unsigned* isink) {
    double a[CH], b[CH], c[CH];
    unsigned u[CH], v[CH];
    unsigned long long w[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        a[i] = 1.0 + threadIdx.x*1e-9 + i; b[i] = 1.0 + (threadIdx.x + i)*1e-12;
        c[i] = 1e-9*(i + 1 + threadIdx.x);
        u[i] = threadIdx.x*7 + i; v[i] = threadIdx.x*13 + i; w[i] = threadIdx.x*17 + i;
    }
    unsigned k1 = 0x9E3779B9u + threadIdx.x, k2 = 0xD2511F53u;
    enum { KD = K2 + K3, KF = (MODE == 2) ? 0 : KD, N = (MODE == 1) ? KD : (KD > 0 ? KD : 1) };
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < KD; ++k) {
            const int i = k % CH;
            if (MODE != 2) {
                if (k*K3/KD != (k + 1)*K3/KD)
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(c[i]));
                else
                    asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[i]));
            }
            if (MODE != 1) {
                if (k*KA/KD != (k + 1)*KA/KD)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(k1), "r"(v[(i + 1) % CH]));
                if (k*KI/KD != (k + 1)*KI/KD)
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(k2), "r"(k1));
                if (k*KW/KD != (k + 1)*KW/KD)
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(v[i]), "r"(k2));
            }
        }
    }
    double s = 0; unsigned x = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { s += a[i] + b[i] + c[i]; x ^= u[i] ^ v[i] ^ (unsigned)w[i] ^ (unsigned)(w[i] >> 32); }
    if (s == 12345.678) sink[0] = s;
    if (x == 0x12345678u) isink[0] = x;
}

template <int CH, int K2, int K3, int KA, int KI, int KW, int MODE>
void run(int sm, const char* label) {
    double* sink; unsigned* isink;
    cudaMalloc(&sink, 8); cudaMalloc(&isink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long iters = 20000;
    probe<CH, K2, K3, KA, KI, KW, MODE><<<2*sm, 256>>>(iters/10, sink, isink);
    cudaEventRecord(e0);
    probe<CH, K2, K3, KA, KI, KW, MODE><<<2*sm, 256>>>(iters, sink, isink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double cyc = ms*1e-3*khz*1e3/(double)(iters*4);     // 4 warps per sub-partition
    printf("%-10s chains=%d fp64=%2d+%2d(3-reg) lop3=%2d imad=%2d wide=%2d : %7.1f cycles per warp-trip per SMSP\n",
           label, CH, K2, K3, KA, KI, KW, cyc);
    cudaFree(sink); cudaFree(isink);
}

template <int CH, int K2, int K3, int KA, int KI, int KW>
void group(int sm) {
    run<CH, K2, K3, KA, KI, KW, 1>(sm, "fp64 only");
    run<CH, K2, K3, KA, KI, KW, 2>(sm, "int only");
    run<CH, K2, K3, KA, KI, KW, 0>(sm, "mixed");
}

int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    // the product kernel's mix per path-step (profiles/r02_ncu_pipe_instructions.csv):
    // 45 FP64 (10 three-register DFMAs), 31 ALU-pipe (LOP3, SHF, LEA ...), 8 IMAD, 10 IMAD.WIDE
    group<2, 35, 10, 31, 8, 10>(sm);
    group<4, 35, 10, 31, 8, 10>(sm);
    group<8, 35, 10, 31, 8, 10>(sm);
    // the mix at the start of the round: 53 FP64 (9 three-register), 36 ALU, 11 IMAD, 10 WIDE
    group<2, 44, 9, 36, 11, 10>(sm);
    group<4, 44, 9, 36, 11, 10>(sm);
    return 0;
}
