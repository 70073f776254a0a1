// heston_ilp_probe.cu -- does instruction-level parallelism across PATHS help the
// Heston step?  Stand-alone loop built from the engine's own device functions
// (Philox4x32-10, normal_pair, HestonSDE<1,false>::step; csrc/sde_engine.cuh), without
// the engine's staging / store / statistics machinery: each thread integrates PPT
// independent paths side by side, so that ptxas can interleave PPT dependence chains
// (profiles/r01_ablation.md, section D: the real kernel needs 190 cycles per warp-step
// for a pipe mix that runs in 133 with four independent chains).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xptxas=-v \
//        -o /tmp/heston_ilp_probe tools/heston_ilp_probe.cu && /tmp/heston_ilp_probe
//
// Prints path-steps/s for PPT = 1 (the product's layout) and PPT = 2 at the
// occupancies the register counts allow, plus a checksum of the terminal states (the
// same paths are integrated in every variant: the sums must agree).
#include <cstdio>
#include <cuda_runtime.h>
#include "../sdepy_b200/csrc/sde_engine.cuh"

using namespace sdeb;

struct PArgs {
    NrmK nk;
    u32 rkey[20];
    double pc[9];          // mu, sigma^2/2, sigma, theta, k, xi, L00, L10, L11
    double ds, sq;
    long long n_paths;
    int n_steps;
    double* sum;           // [2] sum of log x_T, sum of y_T
    unsigned long long* neg;
};

template <int PPT, int MINB, int COPIES>
__global__ void __launch_bounds__(256, MINB) heston_probe(const PArgs a) {
    extern __shared__ __align__(16) double tab_mem[];
    fill_tables_t<COPIES>(tab_mem);
    __syncthreads();
    const TabT<COPIES> tab(tab_mem, a.nk.v[14]);
    typedef HestonSDE<1, false> M;
    double sx = 0.0, sy = 0.0;
    unsigned long long negs = 0;
    const long long stride = (long long)gridDim.x*blockDim.x*PPT;
    for (long long base = ((long long)blockIdx.x*blockDim.x + threadIdx.x)*PPT;
         base < a.n_paths; base += stride) {
        double x[PPT][2];
        int cnt[PPT][2];
        Rng rng[PPT];
        U4 blk[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            x[q][0] = 4.605170185988092; x[q][1] = 0.04;     // log 100, y0
            cnt[q][0] = cnt[q][1] = 0;
            const u64 g = (u64)(base + q);
            rng[q].rk = a.rkey; rng[q].c_x = (u32)g; rng[q].c_y = (u32)(g >> 32) & 0xFFu;
            rng[q].step = 0;
            blk[q] = rng[q].block(0u);
        }
        for (int n = 0; n < a.n_steps; n += 2) {
            U4 cur[PPT];
#pragma unroll
            for (int q = 0; q < PPT; ++q) cur[q] = blk[q];
            // step n (first pair of the block)
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                double z[2];
                rng[q].step = (u32)(n >> 1);
                normal_pair(cur[q].x, cur[q].y, tab, a.nk, a.sq, z[0], z[1]);
                z[1] = fma(a.pc[8], z[1], a.pc[7]*z[0]);
                M::step<false>(x[q], a.pc, a.ds, z, z, cnt[q], tab.k375);
            }
            // next period's block, drawn while step n+1 runs (as the engine does)
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                rng[q].step = (u32)(n >> 1) + 1;
                blk[q] = rng[q].block(0u);
            }
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                double z[2];
                rng[q].step = (u32)(n >> 1);
                normal_pair(cur[q].z, cur[q].w, tab, a.nk, a.sq, z[0], z[1]);
                z[1] = fma(a.pc[8], z[1], a.pc[7]*z[0]);
                M::step<false>(x[q], a.pc, a.ds, z, z, cnt[q], tab.k375);
            }
        }
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            if (base + q < a.n_paths) { sx += x[q][0]; sy += x[q][1]; negs += cnt[q][0]; }
        }
    }
    // block reduction (order differs between variants: compare to ~1e-9 relative)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sx += __shfl_down_sync(0xffffffffu, sx, off);
        sy += __shfl_down_sync(0xffffffffu, sy, off);
        negs += __shfl_down_sync(0xffffffffu, negs, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&a.sum[0], sx); atomicAdd(&a.sum[1], sy); atomicAdd(a.neg, negs);
    }
}

template <int PPT, int MINB, int COPIES>
void run(const PArgs& base, int sm) {
    const size_t smem = (size_t)TAB_DOUBLES*COPIES*8;
    auto kern = heston_probe<PPT, MINB, COPIES>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    PArgs a = base;
    cudaMemset(a.sum, 0, 16); cudaMemset(a.neg, 0, 8);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    const int grid = sm*(occ > 0 ? occ : 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    PArgs w = a; w.n_paths = a.n_paths/8;
    kern<<<grid, 256, smem>>>(w);                 // warm-up
    cudaMemset(a.sum, 0, 16); cudaMemset(a.neg, 0, 8);
    cudaEventRecord(e0);
    kern<<<grid, 256, smem>>>(a);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double h[2]; unsigned long long neg;
    cudaMemcpy(h, a.sum, 16, cudaMemcpyDeviceToHost); cudaMemcpy(&neg, a.neg, 8, cudaMemcpyDeviceToHost);
    printf("paths/thread=%d min_blocks=%d table copies=%d: %3d regs, %d CTAs/SM, %8.3f ms, %.4g path-steps/s  "
           "mean log x_T=%.12f mean y_T=%.12f neg=%llu  (%s)\n", PPT, MINB, COPIES, fa.numRegs, occ, ms,
           (double)a.n_paths*a.n_steps/(ms*1e-3), h[0]/a.n_paths, h[1]/a.n_paths, neg,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    PArgs a;
    const NrmK nk = {SDEB_NRMK_VALUES};
    a.nk = nk;
    philox_round_keys(1234ull, a.rkey);
    const double mu = .03, sigma = 1., theta = .04, k = 2., xi = .3, rho = -.7;
    const double pc[9] = {mu, sigma*sigma/2, sigma, theta, k, xi, 1.0, rho, 0.714142842854285};
    for (int i = 0; i < 9; ++i) a.pc[i] = pc[i];
    a.n_steps = 252; a.ds = 1.0/252; a.sq = 0.06299407883487121;   // sqrt(1/252)
    a.n_paths = 20000000;
    cudaMalloc(&a.sum, 16); cudaMalloc(&a.neg, 8);
    run<1, 2, 1>(a, sm);       // round 1's layout: 1 path per thread, 2 CTAs per SM, one table copy
    run<1, 2, 8>(a, sm);       // ... with 8 interleaved table copies (conflict-free look-ups)
    run<1, 3, 8>(a, sm);
    run<2, 1, 1>(a, sm);       // 2 paths per thread, whatever occupancy the registers allow
    run<2, 1, 8>(a, sm);
    run<2, 2, 1>(a, sm);       // 2 paths per thread squeezed under 128 registers
    run<2, 2, 8>(a, sm);
    run<3, 1, 8>(a, sm);
    run<4, 1, 8>(a, sm);
    return 0;
}
