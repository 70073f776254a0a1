// ws_probe.cu -- does warp specialisation pay on sm_100a for the Heston step mix?
//
// integrate_lean_kernel<Heston> issues, per path-step, ~52 FP64-pipe, ~26 ALU and
// ~13 IMAD-class (10 of them IMAD.WIDE) instructions from EVERY warp, in program
// order; FP64 and integer pipes are each busy about half of the time and overlap
// poorly (profiles/r01_summary.md, r01_ablation.md).  This probe times the same
// instruction mix at the same occupancy (2 x 256-thread CTAs per SM) in two layouts:
//   mixed        every warp: per trip KD x DFMA + KA x LOP3 + KI x IMAD + KW x IMAD.WIDE,
//                interleaved the way ptxas interleaves them in the real loop
//   specialised  even warps: 2 x KD DFMA per trip; odd warps: 2 x (KA + KI + KW) integer
//                instructions per trip  (same work per SM, one role per warp)
// If `specialised` is clearly faster, a producer/consumer split of the kernel
// (integer warps filling a shared-memory ring with Philox bits and table indices,
// FP64 warps consuming it) is worth building; if not, the pipes share a resource
// and only removing work helps.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ws_probe tools/ws_probe.cu && /tmp/ws_probe
#include <cstdio>
#include <cuda_runtime.h>

#define DFMA(x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(m), "d"(c))
#define LOP3(x, y) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(k1), "r"(y))
#define IMAD(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(k2), "r"(k1))
#define WIDE(w, x) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(x), "r"(k2))

enum { CH = 4 };   // independent chains per thread (the real loop has ~2-4 way ILP)

template <int KD, int KA, int KI, int KW, int MODE>   // MODE 0 mixed, 1 specialised, 2 fp64 only, 3 int only
__global__ void __launch_bounds__(256, 2) probe(long long iters, double* sink, unsigned* isink) {
    double a[CH];
    unsigned u[CH], v[CH];
    unsigned long long w[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        a[i] = threadIdx.x*1e-9 + i; u[i] = threadIdx.x*7 + i; v[i] = threadIdx.x*13 + i;
        w[i] = threadIdx.x*17 + i;
    }
    double m = 1.0000001 + threadIdx.x*1e-12, c = 1e-9;
    unsigned k1 = 0x9E3779B9u + threadIdx.x, k2 = 0xD2511F53u;
    const bool fp_role = MODE == 2 || (MODE == 1 && ((threadIdx.x >> 5) & 1) == 0);
    const bool int_role = MODE == 3 || (MODE == 1 && ((threadIdx.x >> 5) & 1) == 1);
    for (long long it = 0; it < iters; ++it) {
        if (MODE == 0) {
            // one FP64 instruction, then its share of the integer ones, round robin over chains
#pragma unroll
            for (int k = 0; k < KD; ++k) {
                DFMA(a[k % CH]);
                if (k*KA/KD != (k + 1)*KA/KD) LOP3(u[k % CH], v[(k + 1) % CH]);
                if (k*KI/KD != (k + 1)*KI/KD) IMAD(v[k % CH]);
                if (k*KW/KD != (k + 1)*KW/KD) WIDE(w[k % CH], v[k % CH]);
            }
        } else {
            if (fp_role) {
#pragma unroll
                for (int k = 0; k < (MODE == 1 ? 2 : 1)*KD; ++k) DFMA(a[k % CH]);
            }
            if (int_role) {
#pragma unroll
                for (int r = 0; r < (MODE == 1 ? 2 : 1); ++r) {
#pragma unroll
                    for (int k = 0; k < KA; ++k) LOP3(u[k % CH], v[(k + 1) % CH]);
#pragma unroll
                    for (int k = 0; k < KI; ++k) IMAD(v[k % CH]);
#pragma unroll
                    for (int k = 0; k < KW; ++k) WIDE(w[k % CH], v[k % CH]);
                }
            }
        }
    }
    double s = 0; unsigned x = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { s += a[i]; x ^= u[i] ^ v[i] ^ (unsigned)w[i] ^ (unsigned)(w[i] >> 32); }
    if (s == 12345.678) sink[0] = s;
    if (x == 0x12345678u) isink[0] = x;
}

template <int KD, int KA, int KI, int KW, int MODE>
float run(int sm, const char* label) {
    double* sink; unsigned* isink;
    cudaMalloc(&sink, 8); cudaMalloc(&isink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long iters = 20000;
    probe<KD, KA, KI, KW, MODE><<<2*sm, 256>>>(iters/10, sink, isink);
    cudaEventRecord(e0);
    probe<KD, KA, KI, KW, MODE><<<2*sm, 256>>>(iters, sink, isink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // cycles per trip per sub-partition: 4 warps per SMSP, each doing `iters` trips
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double cyc = ms*1e-3*khz*1e3/(double)(iters*4);
    printf("%-12s dfma=%2d lop3=%2d imad=%2d wide=%2d : %8.3f ms = %6.1f cycles per warp-trip per SMSP\n",
           label, KD, KA, KI, KW, ms, cyc);
    cudaFree(sink); cudaFree(isink);
    return ms;
}

// ---------------------------------------------------------------------------
// second probe: the mixed layout plus the step's OTHER instructions, one kind at a
// time -- XU (2 MUFU.RSQ64H + 2 I2F + 1 FLO per step), shared-memory table look-ups
// (2 random LDS.128 of a 4 KB table + 1 uniform LDS.128 per step: ~23 wavefronts with
// the bank conflicts) and the 3 rarely-taken branches -- to see which of them
// accounts for the distance between `mixed` (133 cycles) and the real kernel (190).
// ---------------------------------------------------------------------------
template <int XU, int LDS, int BR>
__global__ void __launch_bounds__(256, 2) probe2(long long iters, double* sink, unsigned* isink) {
    enum { KD = 52, KA = 26, KI = 3, KW = 10 };
    __shared__ __align__(16) double tab[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tab[i] = 1.0 + i*1e-6;
    __syncthreads();
    double a[CH];
    unsigned u[CH], v[CH];
    unsigned long long w[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        a[i] = threadIdx.x*1e-9 + i + 1; u[i] = threadIdx.x*7 + i; v[i] = threadIdx.x*13 + i;
        w[i] = threadIdx.x*17 + i;
    }
    double m = 1.0000001 + threadIdx.x*1e-12, c = 1e-9;
    unsigned k1 = 0x9E3779B9u + threadIdx.x, k2 = 0xD2511F53u;
    unsigned saddr = (unsigned)__cvta_generic_to_shared(tab);
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < KD; ++k) {
            DFMA(a[k % CH]);
            if (k*KA/KD != (k + 1)*KA/KD) LOP3(u[k % CH], v[(k + 1) % CH]);
            if (k*KI/KD != (k + 1)*KI/KD) IMAD(v[k % CH]);
            if (k*KW/KD != (k + 1)*KW/KD) WIDE(w[k % CH], v[k % CH]);
            if (XU && k % 10 == 3) {          // 5 XU instructions per trip
                if (k % 20 == 3) {
                    double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[k % CH]));
                    a[(k + 1) % CH] += y*1e-30;
                } else {
                    double y = (double)(int)u[k % CH];
                    a[(k + 1) % CH] += y*1e-30;
                }
            }
            if (LDS && (k == 5 || k == 25)) { // 2 random 16-byte look-ups per trip
                double t0, t1;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t0), "=d"(t1)
                             : "r"(saddr + ((u[k % CH] >> 8) & 0xFF0u)));
                a[(k + 2) % CH] += (t0 + t1)*1e-30;
            }
            if (BR && (k == 10 || k == 30 || k == 45)) {   // rarely taken, like the tail / tiny paths
                if ((u[k % CH] & 0xFFF00000u) == 0xABC00000u) {
                    a[k % CH] = sqrt(a[k % CH] + 3.0);
                }
            }
        }
    }
    double s = 0; unsigned x = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { s += a[i]; x ^= u[i] ^ v[i] ^ (unsigned)w[i] ^ (unsigned)(w[i] >> 32); }
    if (s == 12345.678) sink[0] = s;
    if (x == 0x12345678u) isink[0] = x;
}

template <int XU, int LDS, int BR>
void run2(int sm, const char* label) {
    double* sink; unsigned* isink;
    cudaMalloc(&sink, 8); cudaMalloc(&isink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long iters = 20000;
    probe2<XU, LDS, BR><<<2*sm, 256>>>(iters/10, sink, isink);
    cudaEventRecord(e0);
    probe2<XU, LDS, BR><<<2*sm, 256>>>(iters, sink, isink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("mixed %-22s : %8.3f ms = %6.1f cycles per warp-trip per SMSP\n", label, ms,
           ms*1e-3*khz*1e3/(double)(iters*4));
    cudaFree(sink); cudaFree(isink);
}

int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    // the Heston step: 52 FP64, 26 ALU, 3 IMAD, 10 IMAD.WIDE
    run<52, 26, 3, 10, 2>(sm, "fp64 only");
    run<52, 26, 3, 10, 3>(sm, "int only");
    float mixed = run<52, 26, 3, 10, 0>(sm, "mixed");
    float spec = run<52, 26, 3, 10, 1>(sm, "specialised");
    printf("specialised / mixed = %.3f (per-SM work identical)\n", spec/mixed);
    // a lighter integer side (Philox4x32-7 would be 52 / 23 / 3 / 7)
    mixed = run<52, 23, 3, 7, 0>(sm, "mixed r7");
    spec = run<52, 23, 3, 7, 1>(sm, "special. r7");
    printf("specialised / mixed = %.3f\n", spec/mixed);
    run2<0, 0, 0>(sm, "(plain)");
    run2<1, 0, 0>(sm, "+ 5 XU");
    run2<0, 1, 0>(sm, "+ 2 random LDS.128");
    run2<0, 0, 1>(sm, "+ 3 branches");
    run2<1, 1, 1>(sm, "+ XU + LDS + branches");
    return 0;
}
