// sdeb.cu -- C ABI (include/sdeb.h) over the sm_100a kernels of sde_engine.cuh.
// Build: see __graft_entry__.build() / sdepy_b200/_build.py
//   nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "sdeb_internal.h"

using namespace sdeb;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t e_ = (expr);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            cudaGetLastError();   /* do not leave it for the next call's check */   \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver \
                            ? SDEB_ENODEV : SDEB_ECUDA,                             \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));        \
        }                                                                           \
    } while (0)

extern "C" int sdeb_abi_version(void) { return SDEB_ABI_VERSION; }
extern "C" const char* sdeb_last_error(void) { return g_err.c_str(); }

extern "C" int sdeb_device_info(int64_t* sm_count, int64_t* cc_major, int64_t* cc_minor,
                                int64_t* total_mem_bytes) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem_bytes) *total_mem_bytes = (int64_t)prop.totalGlobalMem;
    return SDEB_OK;
}

// ---------------------------------------------------------------------------
// model registry: pre-instantiated presets live in sdeb_models_*.cu
// ---------------------------------------------------------------------------
struct JitModule {
    std::string source, name;
    cudaLibrary_t lib;                 // stage 0: lean entry + dims
    cudaKernel_t kernel_lean;
    cudaLibrary_t glib[10];            // general entry, one library per sweep variant
    cudaKernel_t gkernel[10];          // (0..5), stream entry per variant (6..9)
    bool ghave[10];
    ModelInfo mi;
};
static std::mutex g_jit_mutex;
static std::map<int64_t, JitModule> g_jit;
static int64_t g_jit_next = 1;

static bool lookup_model(int64_t model, int64_t n, int64_t jit_handle, ModelInfo& mi) {
    if (model == SDEB_MODEL_JIT) {
        std::lock_guard<std::mutex> lock(g_jit_mutex);
        auto it = g_jit.find(jit_handle);
        if (it == g_jit.end()) return false;
        mi = it->second.mi;
        return true;
    }
    return sdeb_lookup_unit_0(model, n, mi) || sdeb_lookup_unit_1(model, n, mi) ||
           sdeb_lookup_unit_2(model, n, mi) || sdeb_lookup_unit_3(model, n, mi) ||
           sdeb_lookup_unit_4(model, n, mi) || sdeb_lookup_unit_5(model, n, mi) ||
           sdeb_lookup_unit_6(model, n, mi) || sdeb_lookup_unit_7(model, n, mi) ||
           sdeb_lookup_unit_8(model, n, mi) || sdeb_lookup_unit_9(model, n, mi) ||
           sdeb_lookup_unit_10(model, n, mi);
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
static const int kThreads = SDEB_THREADS;
static const int64_t kMaxStatsSmem = 96 * 1024;   // accumulators kept in smem up to here

static int64_t smem_bytes(const ModelInfo& mi, const sdeb_problem* p, bool stats_in_kernel) {
    int64_t nch = mi.ndw > 1 ? (int64_t)mi.ndw * (mi.ndw + 1) / 2 : 0;
    int64_t npt = mi.npc + nch;
    // dynamic part only; the staged step block is static __shared__.  The generator
    // tables start on the first TAB_ALIGN boundary of the dynamic window (reserve a
    // whole TAB_ALIGN in front of them); parameter records (staged only when
    // time-dependent and shared by the paths) and the warp scratch live inside that
    // gap when they fit its first TAB_FRONT bytes (sde_engine.cuh:integrate_body)
    const bool staged = p->n_psteps > 1 && !p->params_per_path;
    int64_t small = ((staged ? (int64_t)STEP_CHUNK * npt : 0) + 8 * NSTAT * mi.nx) * 8;
    int64_t bytes = TAB_ALIGN + (int64_t)TAB_DOUBLES * SDEB_TAB_COPIES * 8 +
                    (small > TAB_FRONT ? small : 0);
    if (stats_in_kernel) bytes += p->n_rows * p->n_groups * mi.nx * NSTAT * 8;
    if (p->noise == SDEB_NOISE_REPLAY)       // cp.async ring of the replay table
        bytes += (int64_t)replay_depth(mi.ndw) * mi.ndw * kThreads * 8;
    return bytes;
}

// The stream kernel serves the full-path output mode of diffusions without jumps:
// paths stored (float64), no statistics / dumps / antithetic pairing / per-path
// records, 16-byte aligned rows.  SDEB_NO_STREAM=1 in the environment sends
// everything to the general kernel (used by the tests to compare the two).
static bool aligned16(const void* q) { return (((uintptr_t)q) & 15) == 0; }
static bool stream_shape(const sdeb_problem* p, const ModelInfo& mi) {
    if ((mi.jumps && p->noise != SDEB_NOISE_REPLAY) || !p->out || p->stats ||
        p->out_dtype != 0 || p->params_per_path ||
        p->dW_dump || p->dJ_dump || p->dN_dump || p->anti_dw_half || p->anti_dj_half ||
        (p->pitch & 1) || !aligned16(p->out) || p->n_steps < 1)
        return false;
    if (p->noise == SDEB_NOISE_REPLAY &&
        (!aligned16(p->dW) || (mi.jumps && (!aligned16(p->dJ) || !aligned16(p->dN)))))
        return false;
    const char* off = getenv("SDEB_NO_STREAM");
    return !(off && off[0] == '1');
}
static int stream_variant(const sdeb_problem* p) {
    return 2 * (p->noise == SDEB_NOISE_REPLAY ? 1 : 0) + (p->n_psteps > 1 ? 1 : 0);
}
static int64_t stream_smem_bytes(const ModelInfo& mi, const sdeb_problem* p) {
    int64_t nch = mi.ndw > 1 ? (int64_t)mi.ndw * (mi.ndw + 1) / 2 : 0;
    int64_t par = p->n_psteps > 1 ? (int64_t)STEP_CHUNK * ((mi.npc + nch + 1) & ~(int64_t)1) * 8 : 0;
    if (p->noise != SDEB_NOISE_REPLAY)      // tables on a TAB_ALIGN boundary, records in the gap
        return TAB_ALIGN + (int64_t)TAB_DOUBLES * SDEB_TAB_COPIES * 8 + (par > TAB_FRONT ? par : 0);
    const int64_t ne = mi.ndw + (mi.jumps ? 2 * (int64_t)mi.jumps * mi.nw : 0);
    return par + (int64_t)stream_depth((int)ne) * ne * kThreads * 2 * 8;
}

// the lean kernel serves the hot configuration: Philox draws, one
// time-invariant parameter record (passed through the constant bank), no dump
static bool use_lean(const sdeb_problem* p, const ModelInfo& mi) {
    return mi.fn_lean && p->noise == SDEB_NOISE_PHILOX && p->params_host &&
           p->n_psteps == 1 && p->n_groups == 1 && !p->dW_dump && !p->dJ_dump && !p->dN_dump &&
           !p->anti_dw_half && !p->anti_dj_half && !p->params_per_path;
}

// index of the sweep variant integrate_body takes for this problem (SDEB_SWEEPS bit)
static int sweep_variant(const sdeb_problem* p) {
    int noise = p->noise == SDEB_NOISE_REPLAY ? 1 : (p->dW_dump ? 2 : 0);
    return 2 * noise + (p->n_psteps > 1 ? 1 : 0);
}
static int jit_general_kernel(int64_t handle, int variant, const void** fn);

static int plan_impl(const sdeb_problem* p, sdeb_plan_t* plan, ModelInfo& mi, bool need_device) {
    if (!p || !plan) return fail(SDEB_EINVAL, "null problem/plan");
    if (p->abi_version != SDEB_ABI_VERSION)
        return fail(SDEB_EINVAL, "sdeb_problem.abi_version mismatch");
    if (!lookup_model(p->model, p->ncomp, p->jit_handle, mi)) {
        char buf[160];
        snprintf(buf, sizeof buf, "model %lld with ncomp=%lld is not pre-instantiated "
                 "(use the JIT path)", (long long)p->model, (long long)p->ncomp);
        return fail(SDEB_EINVAL, buf);
    }
    if (p->n_paths < 0 || p->n_steps < 0 || p->n_groups < 1 || p->n_rows < 0 ||
        p->pitch < p->n_paths)
        return fail(SDEB_EINVAL, "inconsistent sizes (n_paths/n_steps/n_groups/n_rows/pitch)");
    if (p->n_groups >= (1 << 24)) return fail(SDEB_EINVAL, "n_groups must be < 2^24");
    if (p->n_psteps != 1 && p->n_psteps != p->n_steps)
        return fail(SDEB_EINVAL, "n_psteps must be 1 or n_steps");
    memset(plan, 0, sizeof *plan);
    plan->nw = mi.nw; plan->ndw = mi.ndw; plan->nx = mi.nx; plan->npc = mi.npc;
    plan->npt = mi.npc + (mi.ndw > 1 ? (int64_t)mi.ndw * (mi.ndw + 1) / 2 : 0);
    plan->ncnt = mi.ncnt; plan->jumps = mi.jumps;
    plan->threads = kThreads;
    int64_t acc = p->n_rows * p->n_groups * mi.nx * NSTAT * 8;
    plan->stats_in_kernel = (acc <= kMaxStatsSmem) ? 1 : 0;
    bool sik = p->stats != NULL && plan->stats_in_kernel;
    plan->smem_bytes = smem_bytes(mi, p, sik);
    // 227 KB per CTA on sm_100a, ~2 KB of it static (step block, store rows)
    if (plan->smem_bytes > (227 - 2) * 1024)
        return fail(SDEB_EINVAL,
                    "this model does not fit the per-block shared memory (" +
                    std::to_string(plan->smem_bytes / 1024) + " KB of staged parameter records "
                    "and replay ring needed, 225 KB available): too many correlated components "
                    "for time-dependent parameters and/or replayed increments");
    int64_t tiles = ((p->n_paths + kThreads - 1) / kThreads) * p->n_groups;
    int sm = 148, occ = 2;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && p->model == SDEB_MODEL_JIT && !use_lean(p, mi) && p->n_paths > 0 &&
        !(stream_shape(p, mi) && stream_smem_bytes(mi, p) <= (227 - 2) * 1024)) {
        // NVRTC models: compile the general kernel of this run's sweep variant
        // (a problem without paths is a query for the model's dimensions)
        int rcj = jit_general_kernel(p->jit_handle, sweep_variant(p), &mi.fn);
        if (rcj) return rcj;
    }
    // kernel choice: lean (hot Philox configuration) > stream (full-path output) > general
    bool stream = !use_lean(p, mi) && stream_shape(p, mi) &&
                  (p->model == SDEB_MODEL_JIT || mi.fn_stream[stream_variant(p)]) &&
                  stream_smem_bytes(mi, p) <= (227 - 2) * 1024;
    if (stream && e == cudaSuccess && p->model == SDEB_MODEL_JIT) {
        int rcj = jit_general_kernel(p->jit_handle, 6 + stream_variant(p),
                                     &mi.fn_stream[stream_variant(p)]);
        if (rcj) return rcj;
    }
    if (stream) {
        plan->smem_bytes = stream_smem_bytes(mi, p);
        tiles = ((p->n_paths + 2 * kThreads - 1) / (2 * kThreads)) * p->n_groups;
    }
    plan->kernel = stream ? 2 : (use_lean(p, mi) ? 1 : 0);
    const void* fn = stream ? mi.fn_stream[stream_variant(p)]
                            : (use_lean(p, mi) ? mi.fn_lean : mi.fn);
    if (e == cudaSuccess && fn) {
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
        if (plan->smem_bytes > 32 * 1024)   // + ~2 KB of static shared memory
            cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)plan->smem_bytes);
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn, kThreads,
                                                          (size_t)plan->smem_bytes) == cudaSuccess && o > 0)
            occ = o;
    } else if (e != cudaSuccess) {
        cudaGetLastError();
        if (need_device) return fail(SDEB_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    int64_t blocks = (int64_t)sm * occ;          // persistent grid: resident CTAs only
    if (p->max_blocks > 0 && blocks > p->max_blocks) blocks = p->max_blocks;
    if (blocks > tiles) blocks = tiles;
    if (blocks < 1) blocks = 1;
    plan->blocks = blocks;
    plan->workspace_bytes = blocks * acc;
    return SDEB_OK;
}

extern "C" int sdeb_plan(const sdeb_problem* p, sdeb_plan_t* plan) {
    ModelInfo mi;
    return plan_impl(p, plan, mi, false);
}

// fold per-block statistics partials (deterministic: fixed lane assignment and
// shuffle tree).  One warp per output element; lane l takes blocks l, l+32, ...
__global__ void fold_partials_kernel(const double* partials, int64_t n_blocks, int64_t len,
                                     double* stats) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= len) return;
    const int k = (int)(i % NSTAT);
    double acc = (k == 4) ? __longlong_as_double(0x7FF0000000000000LL)
               : (k == 5) ? __longlong_as_double(0xFFF0000000000000LL) : 0.0;
    for (int64_t b = lane; b < n_blocks; b += 32) {
        double o = partials[b * len + i];
        acc = (k == 4) ? fmin(acc, o) : (k == 5) ? fmax(acc, o) : acc + o;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double o = __shfl_down_sync(0xffffffffu, acc, off);
        acc = (k == 4) ? fmin(acc, o) : (k == 5) ? fmax(acc, o) : acc + o;
    }
    if (lane == 0) stats[i] = acc;
}

extern "C" int sdeb_integrate(const sdeb_problem* p, void* stream_) {
    ModelInfo mi;
    sdeb_plan_t plan;
    int rc = plan_impl(p, &plan, mi, true);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (p->n_paths == 0) return SDEB_OK;
    if (!p->steps && p->n_steps > 0) return fail(SDEB_EINVAL, "steps table is NULL");
    if (!p->params || !p->w0) return fail(SDEB_EINVAL, "params / w0 is NULL");
    if (p->noise == SDEB_NOISE_REPLAY) {
        if (!p->dW) return fail(SDEB_EINVAL, "replay mode needs dW");
        if (mi.jumps && !p->dJ) return fail(SDEB_EINVAL, "replay mode needs dJ for jump models");
    }
    if (p->stats) {
        if (!plan.stats_in_kernel)
            return fail(SDEB_EINVAL, "statistics accumulators exceed shared memory: "
                        "store the rows and use sdeb_moments");
        if (!p->workspace || p->workspace_bytes < plan.workspace_bytes)
            return fail(SDEB_EINVAL, "workspace too small (see sdeb_plan)");
        if (!p->centre) return fail(SDEB_EINVAL, "stats need a centre vector");
    }
    KArgs a;
    memset(&a, 0, sizeof a);
    { static const NrmK nk = {SDEB_NRMK_VALUES}; a.nk = nk; }
    philox_round_keys(p->seed, a.rkey);
    const bool lean = use_lean(p, mi);
    if (lean) {
        // single time-invariant record: serve it from the constant bank
        for (int k = 0; k < plan.npt; ++k) a.pc[k] = p->params_host[k];
    }
    a.n_paths = p->n_paths; a.path_offset = p->path_offset; a.pitch = p->pitch;
    a.n_steps = (int)p->n_steps; a.n_groups = (int)p->n_groups; a.n_rows = (int)p->n_rows;
    a.row0 = (int)p->row0; a.n_psteps = (int)p->n_psteps; a.w0_per_path = (int)p->w0_per_path;
    a.noise = (int)p->noise; a.params_pp = (int)p->params_per_path;
    if (p->out_dtype < 0 || p->out_dtype > 2)
        return fail(SDEB_EINVAL, "out_dtype must be 0 (float64), 1 (float32) or 2 (float16)");
    a.out_dtype = (int)p->out_dtype;
    a.payoff_kind = (int)p->payoff_kind; a.payoff_strike = p->payoff_strike;
    a.payoff_scale = p->payoff_scale;
    a.seed = p->seed;
    a.steps = p->steps; a.store_row = p->store_row; a.params = p->params; a.w0 = p->w0;
    a.dW = p->dW; a.dJ = p->dJ; a.dN = (const i64*)p->dN;
    a.out = p->out; a.partials = p->stats ? (double*)p->workspace : NULL;
    a.centre = p->centre; a.counter = (i64*)p->counter; a.dn_sum = (i64*)p->dn_sum;
    a.dW_dump = p->dW_dump; a.dJ_dump = p->dJ_dump; a.dN_dump = (i64*)p->dN_dump;
    a.anti_dw_half = p->anti_dw_half; a.anti_dj_half = p->anti_dj_half;

    void* args[] = {&a};
    const void* fn = plan.kernel == 2 ? mi.fn_stream[stream_variant(p)]
                                      : (lean ? mi.fn_lean : mi.fn);
    CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)plan.blocks), dim3(kThreads), args,
                              (size_t)plan.smem_bytes, stream));
    if (p->stats) {
        int64_t len = p->n_rows * p->n_groups * mi.nx * NSTAT;
        fold_partials_kernel<<<(unsigned)((len * 32 + 127) / 128), 128, 0, stream>>>(
            (const double*)p->workspace, plan.blocks, len, p->stats);
        CUDA_TRY(cudaGetLastError());
    }
    return SDEB_OK;
}

// ---------------------------------------------------------------------------
// across-path statistics of stored rows
// ---------------------------------------------------------------------------
// Streaming reader of one row: 16-byte loads, four independent loads in flight
// per thread (8 CTAs x 256 threads x 64 B = 128 KB per SM: HBM latency x
// bandwidth needs ~45 KB), evict-first.  f(v) is applied to every element
// exactly once; which thread sees which element is fixed by the launch
// geometry, so the block partials -- folded in block order -- are reproducible.
template <class F>
__device__ __forceinline__ void stream_row(const double* __restrict__ xr, int64_t n, F&& f) {
    const int64_t head = ((((uintptr_t)xr) & 15) != 0 && n > 0) ? 1 : 0;
    const int64_t nvec = (n - head) >> 1;
    const double2* __restrict__ xv = (const double2*)(xr + head);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = tid;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const double2 a = __ldcs(xv + i), b = __ldcs(xv + i + stride);
        const double2 c = __ldcs(xv + i + 2 * stride), d = __ldcs(xv + i + 3 * stride);
        f(a.x); f(a.y); f(b.x); f(b.y); f(c.x); f(c.y); f(d.x); f(d.y);
    }
    for (; i < nvec; i += stride) {
        const double2 a = __ldcs(xv + i);
        f(a.x); f(a.y);
    }
    if (tid == 0) {
        if (head) f(xr[0]);
        if ((n - head) & 1) f(xr[n - 1]);
    }
}

// block reduction of one NSTAT vector -> partials[block][row][NSTAT]
__device__ __forceinline__ void block_fold_stats(double (&st)[NSTAT], double* partials, int row) {
    __shared__ double s_warp[8][NSTAT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < NSTAT; ++k) {
            double o = __shfl_down_sync(0xffffffffu, st[k], off);
            st[k] = (k == 4) ? fmin(st[k], o) : (k == 5) ? fmax(st[k], o) : st[k] + o;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NSTAT; ++k) s_warp[warp][k] = st[k];
    }
    __syncthreads();
    if (threadIdx.x < NSTAT) {
        int k = threadIdx.x;
        double acc = s_warp[0][k];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            double o = s_warp[w][k];
            acc = (k == 4) ? fmin(acc, o) : (k == 5) ? fmax(acc, o) : acc + o;
        }
        // layout [block][row][NSTAT] so that fold_partials_kernel applies
        partials[((int64_t)blockIdx.x * gridDim.y + row) * NSTAT + k] = acc;
    }
}

// RANGE_ONLY: sum, min, max (pass 1 of montecarlo's first update: the centring
// constant and the histogram range); else S1..S4 about `centre`, min, max
template <bool RANGE_ONLY>
__global__ void __launch_bounds__(256)
moments_kernel(const double* x, int64_t n_paths, int64_t pitch, const double* centre,
               double* partials) {
    const int row = blockIdx.y;
    const double c = centre ? centre[row] : 0.0;
    const double* xr = x + (int64_t)row * pitch;
    double st[NSTAT];
    st[0] = st[1] = st[2] = st[3] = st[6] = st[7] = 0.0;
    st[4] = __longlong_as_double(0x7FF0000000000000LL);
    st[5] = __longlong_as_double(0xFFF0000000000000LL);
    stream_row(xr, n_paths, [&](double v) {
        if (RANGE_ONLY) {
            st[0] += v;
        } else {
            double d = v - c, d2 = d * d;
            st[0] += d; st[1] += d2; st[2] = fma(d2, d, st[2]); st[3] = fma(d2, d2, st[3]);
        }
        st[4] = fmin(st[4], v); st[5] = fmax(st[5], v);
    });
    block_fold_stats(st, partials, row);
}

// grid.x per row: ~8 resident CTAs per SM over all rows, at least 2 per row
static int64_t mom_blocks(int64_t n_rows, int64_t n) {
    int64_t per = 1184 / n_rows;
    if (per < 2) per = 2;
    int64_t need = (n + 2047) / 2048;
    if (need < 1) need = 1;
    return need < per ? need : per;
}

extern "C" int64_t sdeb_moments_workspace(int64_t n_rows) {
    if (n_rows < 1) n_rows = 1;
    int64_t per = 1184 / n_rows;
    if (per < 2) per = 2;
    return per * n_rows * NSTAT * 8;
}

static int moments_impl(bool range_only, const double* x, int64_t n_rows, int64_t n_paths,
                        int64_t pitch, const double* centre, double* stats, void* workspace,
                        int64_t workspace_bytes, void* stream_) {
    if (!x || !stats || n_rows < 1 || n_paths < 1 || pitch < n_paths)
        return fail(SDEB_EINVAL, "sdeb_moments: bad arguments");
    if (n_rows > 65535) return fail(SDEB_EINVAL, "sdeb_moments: n_rows > 65535 (chunk the rows)");
    if (!workspace || workspace_bytes < sdeb_moments_workspace(n_rows))
        return fail(SDEB_EINVAL, "sdeb_moments: workspace too small");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int blocks = (int)mom_blocks(n_rows, n_paths);
    if (range_only)
        moments_kernel<true><<<dim3(blocks, (unsigned)n_rows), 256, 0, stream>>>(
            x, n_paths, pitch, NULL, (double*)workspace);
    else
        moments_kernel<false><<<dim3(blocks, (unsigned)n_rows), 256, 0, stream>>>(
            x, n_paths, pitch, centre, (double*)workspace);
    CUDA_TRY(cudaGetLastError());
    int64_t len = n_rows * NSTAT;
    fold_partials_kernel<<<(unsigned)((len * 32 + 127) / 128), 128, 0, stream>>>(
        (const double*)workspace, blocks, len, stats);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

extern "C" int sdeb_moments(const double* x, int64_t n_rows, int64_t n_paths, int64_t pitch,
                            const double* centre, double* stats, void* workspace,
                            int64_t workspace_bytes, void* stream_) {
    return moments_impl(false, x, n_rows, n_paths, pitch, centre, stats, workspace,
                        workspace_bytes, stream_);
}

extern "C" int sdeb_mc_range(const double* x, int64_t n_rows, int64_t n, int64_t pitch,
                             double* stats, void* workspace, int64_t workspace_bytes,
                             void* stream_) {
    return moments_impl(true, x, n_rows, n, pitch, NULL, stats, workspace, workspace_bytes,
                        stream_);
}

// ---------------------------------------------------------------------------
// montecarlo update in ONE pass over the sample (reference infrastructure.py:
// 2924-3021): centred power sums S1..S4 AND the histogram of every row.
//  * centre: given, or sum/n of the range pass (first sample, 2934)
//  * edges : given, or built here exactly as numpy.linspace(lo, hi, nbins + 1)
//            does for numpy.histogram(range=None | (lo, hi)) (2999-3004):
//            step = (hi - lo)/nbins, e_i = fl(fl(i*step) + lo), e_nbins = hi;
//            lo == hi widens to (lo - 1/2, hi + 1/2)
//  * bins  : numpy.histogram semantics (half-open, last closed), arithmetic
//            guess + edge fix-up against the edges in shared memory; counts in
//            WARP-PRIVATE 32-bit shared counters (one atomic per element hits
//            only the lanes of its own warp that share the bin), flushed with
//            64-bit global atomics -- integer, hence order-independent.
// ---------------------------------------------------------------------------
enum { MC_EDGES_GIVEN = 0, MC_EDGES_MINMAX = 1, MC_EDGES_RANGE = 2 };

__global__ void __launch_bounds__(256)
mc_update_kernel(const double* x, int64_t n, int64_t pitch, const double* centre,
                 double* centre_out,
                 const double* range_stats, double range_lo, double range_hi, int edges_mode,
                 double* edges, int nbins, int uniform, int copies,
                 double* partials, unsigned long long* counts, unsigned long long* outside) {
    extern __shared__ double sh[];
    __shared__ int s_exact;       // edges are exactly linspace(lo, hi, nbins + 1)
    const int row = blockIdx.y;
    double* s_edges = sh;                                       // nbins + 1
    unsigned int* s_cnt = (unsigned int*)(sh + nbins + 1);      // copies x (nbins + 1)
    const int cstride = nbins + 1;                              // last slot: outside
    const double* xr = x + (int64_t)row * pitch;
    const double c = centre ? centre[row] : __ddiv_rn(range_stats[row * NSTAT], (double)n);
    if (centre_out && blockIdx.x == 0 && threadIdx.x == 0) centre_out[row] = c;
    if (threadIdx.x == 0) s_exact = uniform;
    __syncthreads();
    if (nbins > 0) {
        double* e = edges + (int64_t)row * (nbins + 1);
        if (edges_mode == MC_EDGES_GIVEN) {
            const double glo = e[0], ghi = e[nbins];
            const double gstep = __ddiv_rn(__dsub_rn(ghi, glo), (double)nbins);
            for (int i = threadIdx.x; i <= nbins; i += blockDim.x) {
                const double v = e[i];
                s_edges[i] = v;
                const double w = (i == nbins) ? ghi : __dadd_rn(__dmul_rn((double)i, gstep), glo);
                if (uniform && !(v == w)) s_exact = 0;          // benign race: all write 0
            }
        } else {
            double lo = range_lo, hi = range_hi;
            if (edges_mode == MC_EDGES_MINMAX) {
                lo = range_stats[row * NSTAT + 4]; hi = range_stats[row * NSTAT + 5];
                // a NaN in the sample poisons the sum: numpy's range is then NaN
                if (range_stats[row * NSTAT] != range_stats[row * NSTAT]) lo = hi = range_stats[row * NSTAT];
            }
            if (lo == hi) { lo = __dsub_rn(lo, 0.5); hi = __dadd_rn(hi, 0.5); }
            const double step = __ddiv_rn(__dsub_rn(hi, lo), (double)nbins);
            for (int i = threadIdx.x; i <= nbins; i += blockDim.x) {
                double v = (i == nbins) ? hi : __dadd_rn(__dmul_rn((double)i, step), lo);
                s_edges[i] = v;
                if (blockIdx.x == 0) e[i] = v;
            }
        }
        for (int i = threadIdx.x; i < copies * cstride; i += blockDim.x) s_cnt[i] = 0;
    }
    __syncthreads();
    const double lo = nbins > 0 ? s_edges[0] : 0.0, hi = nbins > 0 ? s_edges[nbins] : 0.0;
    const double scale = nbins / (hi - lo);
    const double step = __ddiv_rn(__dsub_rn(hi, lo), (double)(nbins > 0 ? nbins : 1));
    const bool exact = s_exact != 0;
    // counter copy of this lane: private to its warp, odd and even lanes apart
    unsigned int* my_cnt = s_cnt + ((((threadIdx.x >> 5) << 1) | (threadIdx.x & 1)) % copies) * cstride;
    double st[NSTAT];
    st[0] = st[1] = st[2] = st[3] = st[6] = st[7] = 0.0;
    st[4] = __longlong_as_double(0x7FF0000000000000LL);
    st[5] = __longlong_as_double(0xFFF0000000000000LL);
    if (nbins > 0 && exact) {
        // Uniform bins.  t = (v - lo) * nbins/(hi - lo) - 1/2, r = rint(t) (two
        // additions with the 1.5 * 2^52 magic number), f = t - r: f is the
        // position of v inside bin r, in bin widths from the bin CENTRE.  The
        // arithmetic t and the exactly-rounded linspace edges e_i disagree with
        // the ideal affine map by at most `margin` bin widths (a few ulps of
        // max|lo|,|hi| over the bin width, plus the rounding of t), so an
        // element farther than that from both edges of bin r is in bin r by
        // numpy's edge comparisons too: no edge is touched.  The few elements
        // within `margin` of an edge (and NaN / Inf, whose f is NaN) take the
        // exact path below.  Out-of-range elements have r outside [0, nbins).
        const double big = fmax(fabs(lo), fabs(hi));
        const double margin = (16.0 * big / step + 16.0 * nbins) * 1.1102230246251565e-16;
        const double f_ok = 0.5 - margin;               // <= 0: everything goes the exact way
        const double magic = 6755399441055744.0;        // 1.5 * 2^52
        stream_row(xr, n, [&](double v) {
            double d = v - c, d2 = d * d;
            st[0] += d; st[1] += d2; st[2] = fma(d2, d, st[2]); st[3] = fma(d2, d2, st[3]);
            const double t = fma(v - lo, scale, -0.5);
            const double r = __dadd_rn(__dadd_rn(t, magic), -magic);
            const double f = t - r;
            int idx;
            if (fabs(f) < f_ok) {
                const unsigned int u = (unsigned int)__double2int_rn(r);   // negative -> huge
                idx = (int)(u < (unsigned int)nbins ? u : (unsigned int)nbins);
            } else if (!(v >= lo && v <= hi)) {
                idx = nbins;                                    // outside (or NaN)
            } else {
                idx = (int)((v - lo) * scale);
                idx = idx < 0 ? 0 : (idx > nbins - 1 ? nbins - 1 : idx);
                while (idx > 0 && v < s_edges[idx]) --idx;
                while (idx < nbins - 1 && v >= s_edges[idx + 1]) ++idx;
            }
            atomicAdd(&my_cnt[idx], 1u);
        });
    } else {
        stream_row(xr, n, [&](double v) {
            double d = v - c, d2 = d * d;
            st[0] += d; st[1] += d2; st[2] = fma(d2, d, st[2]); st[3] = fma(d2, d2, st[3]);
            if (nbins > 0) {
                int idx;
                if (!(v >= lo && v <= hi)) {
                    idx = nbins;                                // outside (or NaN)
                } else {
                    if (uniform) {
                        idx = (int)((v - lo) * scale);
                        idx = idx < 0 ? 0 : (idx > nbins - 1 ? nbins - 1 : idx);
                    } else {
                        int a = 0, b = nbins;                   // largest idx with edges[idx] <= v
                        while (b - a > 1) { int m = (a + b) >> 1; if (s_edges[m] <= v) a = m; else b = m; }
                        idx = a;
                    }
                    while (idx > 0 && v < s_edges[idx]) --idx;
                    while (idx < nbins - 1 && v >= s_edges[idx + 1]) ++idx;
                }
                atomicAdd(&my_cnt[idx], 1u);
            }
        });
    }
    block_fold_stats(st, partials, row);
    if (nbins > 0) {
        __syncthreads();
        for (int i = threadIdx.x; i <= nbins; i += blockDim.x) {
            unsigned int t = 0;
            for (int k = 0; k < copies; ++k) t += s_cnt[k * cstride + i];
            if (t) atomicAdd(i < nbins ? &counts[(int64_t)row * nbins + i] : &outside[row],
                             (unsigned long long)t);
        }
    }
}

extern "C" int sdeb_mc_update(const double* x, int64_t n_rows, int64_t n, int64_t pitch,
                              const double* centre, double* centre_out, const double* range_stats,
                              double range_lo, double range_hi, int64_t edges_mode,
                              double* edges, int64_t nbins, int64_t uniform_edges,
                              double* stats, int64_t* counts, int64_t* outside,
                              void* workspace, int64_t workspace_bytes, void* stream_) {
    if (!x || !stats || n_rows < 1 || n_rows > 65535 || n < 1 || pitch < n || nbins < 0)
        return fail(SDEB_EINVAL, "sdeb_mc_update: bad arguments");
    if (!centre && !range_stats)
        return fail(SDEB_EINVAL, "sdeb_mc_update: needs a centre or the statistics of the range pass");
    if (nbins > 0) {
        if (!edges || !counts || !outside)
            return fail(SDEB_EINVAL, "sdeb_mc_update: histogram buffers are NULL");
        if (nbins > 4000) return fail(SDEB_EINVAL, "sdeb_mc_update: at most 4000 bins");
        if (edges_mode < MC_EDGES_GIVEN || edges_mode > MC_EDGES_RANGE ||
            (edges_mode == MC_EDGES_MINMAX && !range_stats))
            return fail(SDEB_EINVAL, "sdeb_mc_update: bad edges_mode");
    }
    if (!workspace || workspace_bytes < sdeb_moments_workspace(n_rows))
        return fail(SDEB_EINVAL, "sdeb_mc_update: workspace too small");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int blocks = (int)mom_blocks(n_rows, n);
    // warp-private counter copies (odd / even lanes apart): as many as fit ~24 KB
    // (8 CTAs per SM stay resident)
    int copies = nbins > 0 ? (int)(24 * 1024 / (4 * (nbins + 1))) : 1;
    copies = copies < 1 ? 1 : (copies > 16 ? 16 : copies);
    size_t smem = nbins > 0 ? (size_t)(nbins + 1) * (8 + 4 * copies) : 0;
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(mc_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    mc_update_kernel<<<dim3(blocks, (unsigned)n_rows), 256, smem, stream>>>(
        x, n, pitch, centre, centre_out, range_stats, range_lo, range_hi, (int)edges_mode, edges, (int)nbins,
        (int)uniform_edges, copies, (double*)workspace, (unsigned long long*)counts,
        (unsigned long long*)outside);
    CUDA_TRY(cudaGetLastError());
    int64_t len = n_rows * NSTAT;
    fold_partials_kernel<<<(unsigned)((len * 32 + 127) / 128), 128, 0, stream>>>(
        (const double*)workspace, blocks, len, stats);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

// ---------------------------------------------------------------------------
// antithetic fold: out[r][k] = (x[r][k] + sign * x[r][half + k]) / 2
// (montecarlo(use='even'|'odd'), infrastructure.py:2905-2914)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
antithetic_fold_kernel(const double* x, int64_t half, int64_t pitch_in, int64_t pitch_out,
                       double sign, double* out) {
    const int row = blockIdx.y;
    const double* xr = x + (int64_t)row * pitch_in;
    double* o = out + (int64_t)row * pitch_out;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < half;
         i += (int64_t)gridDim.x * blockDim.x)
        o[i] = __ddiv_rn(__dadd_rn(xr[i], __dmul_rn(sign, xr[half + i])), 2.0);
}

extern "C" int sdeb_antithetic_fold(const double* x, int64_t n_rows, int64_t half,
                                    int64_t pitch_in, int64_t pitch_out, int64_t sign,
                                    double* out, void* stream_) {
    if (!x || !out || n_rows < 1 || n_rows > 65535 || half < 1 || pitch_in < 2 * half ||
        pitch_out < half || (sign != 1 && sign != -1))
        return fail(SDEB_EINVAL, "sdeb_antithetic_fold: bad arguments");
    int64_t need = (half + 255) / 256;
    int blocks = (int)(need < 1184 ? need : 1184);
    antithetic_fold_kernel<<<dim3(blocks, (unsigned)n_rows), 256, 0, (cudaStream_t)stream_>>>(
        x, half, pitch_in, pitch_out, (double)sign, out);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

// ---------------------------------------------------------------------------
// process.cdf / process.chf / process interpolation over the paths of one row
// (reference infrastructure.py:544-633, 1125-1209).  The row at time t is
// y = w_hi*y_hi + w_lo*y_lo, w_hi = (t - t_lo)/(t_hi - t_lo),
// w_lo = (t_hi - t)/(t_hi - t_lo) (weights evaluated on the host), each
// product and the sum separately rounded -- the formula of
// scipy.interpolate.interp1d._call_linear (SciPy 1.18, a third-party
// dependency of the reference) -- so that cdf counts match the reference
// exactly.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double row_value(const double* lo, const double* hi, int64_t i,
                                            double w_lo, double w_hi, double, int interp) {
    double y = lo[i];
    if (interp) y = __dadd_rn(__dmul_rn(w_hi, hi[i]), __dmul_rn(w_lo, y));
    return y;
}

enum { EVAL_QCHUNK = 8, CDF_MAX_Q = 2048 };

// One row, possibly interpolated between two stored rows, streamed with 16-byte
// loads (scalar accesses when a row is not 16-byte aligned): f(y) once per path.
template <class F>
__device__ __forceinline__ void stream_interp(const double* __restrict__ lo,
                                              const double* __restrict__ hi, double w_lo,
                                              double w_hi, int interp, int64_t n, F&& f) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool vec = ((((uintptr_t)lo) | (interp ? (uintptr_t)hi : 0)) & 15) == 0;
    if (vec) {
        const int64_t nvec = n >> 1;
        const double2* __restrict__ lv = (const double2*)lo;
        const double2* __restrict__ hv = (const double2*)hi;
        int64_t i = tid;
        for (; i + stride < nvec; i += 2 * stride) {
            double2 a = __ldcs(lv + i), b = __ldcs(lv + i + stride);
            if (interp) {
                const double2 c = __ldcs(hv + i), d = __ldcs(hv + i + stride);
                a.x = __dadd_rn(__dmul_rn(w_hi, c.x), __dmul_rn(w_lo, a.x));
                a.y = __dadd_rn(__dmul_rn(w_hi, c.y), __dmul_rn(w_lo, a.y));
                b.x = __dadd_rn(__dmul_rn(w_hi, d.x), __dmul_rn(w_lo, b.x));
                b.y = __dadd_rn(__dmul_rn(w_hi, d.y), __dmul_rn(w_lo, b.y));
            }
            f(a.x); f(a.y); f(b.x); f(b.y);
        }
        for (; i < nvec; i += stride) {
            double2 a = __ldcs(lv + i);
            if (interp) {
                const double2 c = __ldcs(hv + i);
                a.x = __dadd_rn(__dmul_rn(w_hi, c.x), __dmul_rn(w_lo, a.x));
                a.y = __dadd_rn(__dmul_rn(w_hi, c.y), __dmul_rn(w_lo, a.y));
            }
            f(a.x); f(a.y);
        }
        if (tid == 0 && (n & 1)) f(row_value(lo, hi, n - 1, w_lo, w_hi, 0.0, interp));
    } else {
        for (int64_t i = tid; i < n; i += stride) f(row_value(lo, hi, i, w_lo, w_hi, 0.0, interp));
    }
}

// process.cdf: counts[j] += #{paths : y <= q[j]} for ANY number and order of
// thresholds in ONE pass over the row.  Each block rank-sorts the thresholds in
// shared memory; a value's bucket is the first sorted threshold >= y (binary
// search, log2(nq) steps instead of nq comparisons), counted in warp-private
// shared counters; a prefix sum over the buckets gives the counts (integers:
// exact, order-independent).  NaN values fall in no bucket, NaN thresholds count 0.
__global__ void __launch_bounds__(256)
path_cdf_kernel(const double* lo, const double* hi, double w_lo, double w_hi, int interp,
                int64_t n_paths, const double* q, int nq, unsigned long long* counts) {
    extern __shared__ double sh[];
    double* s_q = sh;                                   // sorted thresholds (NaNs last)
    int* s_perm = (int*)(sh + nq);                      // sorted position -> caller's index
    unsigned int* s_cnt = (unsigned int*)(s_perm + nq); // 8 warps x (nq + 1) buckets
    const int cstride = nq + 1;
    for (int j = threadIdx.x; j < nq; j += blockDim.x) {
        const double qj = q[j];
        const bool nanj = qj != qj;
        int r = 0;
        for (int k = 0; k < nq; ++k) {
            const double qk = q[k];
            const bool nank = qk != qk;
            r += nanj ? (!nank || k < j) : (!nank && (qk < qj || (qk == qj && k < j)));
        }
        s_q[r] = qj; s_perm[r] = j;
    }
    for (int i = threadIdx.x; i < 8 * cstride; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    unsigned int* my = s_cnt + (threadIdx.x >> 5) * cstride;
    stream_interp(lo, hi, w_lo, w_hi, interp, n_paths, [&](double y) {
        int a = 0, b = nq;                              // first index with s_q[idx] >= y
        while (a < b) { const int m = (a + b) >> 1; if (s_q[m] >= y) b = m; else a = m + 1; }
        // (a NaN y compares false everywhere: a ends at nq, the "no threshold" bucket)
        atomicAdd(&my[a], 1u);
    });
    __syncthreads();
    for (int i = threadIdx.x; i <= nq; i += blockDim.x) {
        unsigned int t = 0;
        for (int w = 0; w < 8; ++w) t += s_cnt[w * cstride + i];
        s_cnt[i] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int j = 0; j < nq; ++j) {
            run += s_cnt[j];
            if (run && s_q[j] == s_q[j]) atomicAdd(&counts[s_perm[j]], run);
        }
    }
}

// process.chf: sums of cos(u_j y), sin(u_j y) over the paths.  When the
// frequencies form an arithmetic progression (the usual grid), each chunk of 8
// is evaluated exactly at its first frequency and ROTATED by (cos, sin)(du y)
// for the next seven (7 rotations: ~1e-15) -- 1 + nq/8 sincos per path instead of nq.
__global__ void __launch_bounds__(256)
path_chf_kernel(const double* lo, const double* hi, double w_lo, double w_hi, int interp,
                int64_t n_paths, const double* q, int nq, double* partials) {
    __shared__ double s_red[8][2 * EVAL_QCHUNK];
    __shared__ int s_grid;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        int ok = nq >= 3;
        const double u0 = q[0], du = nq > 1 ? q[1] - q[0] : 0.0;
        double big = 0.0;
        for (int j = 0; j < nq; ++j) big = fmax(big, fabs(q[j]));
        for (int j = 2; j < nq && ok; ++j) ok = fabs(q[j] - (u0 + j * du)) <= 8.9e-16 * big;
        s_grid = ok;
    }
    __syncthreads();
    const bool grid = s_grid != 0;
    const double du = nq > 1 ? q[1] - q[0] : 0.0;
    for (int q0 = 0; q0 < nq; q0 += EVAL_QCHUNK) {
        double qv[EVAL_QCHUNK];
        double a0[EVAL_QCHUNK], a1[EVAL_QCHUNK];
#pragma unroll
        for (int j = 0; j < EVAL_QCHUNK; ++j) {
            qv[j] = (q0 + j < nq) ? q[q0 + j] : 0.0;
            a0[j] = a1[j] = 0.0;
        }
        stream_interp(lo, hi, w_lo, w_hi, interp, n_paths, [&](double y) {
            if (grid) {
                double sn, cs, sd, cd;
                sincos(qv[0] * y, &sn, &cs);
                sincos(du * y, &sd, &cd);
#pragma unroll
                for (int j = 0; j < EVAL_QCHUNK; ++j) {
                    a0[j] += cs; a1[j] += sn;
                    const double c2 = cs * cd - sn * sd;
                    sn = sn * cd + cs * sd; cs = c2;
                }
            } else {
#pragma unroll
                for (int j = 0; j < EVAL_QCHUNK; ++j) {
                    double sn, cs;
                    sincos(qv[j] * y, &sn, &cs);
                    a0[j] += cs; a1[j] += sn;
                }
            }
        });
#pragma unroll
        for (int j = 0; j < EVAL_QCHUNK; ++j) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                a0[j] += __shfl_down_sync(0xffffffffu, a0[j], off);
                a1[j] += __shfl_down_sync(0xffffffffu, a1[j], off);
            }
            if (lane == 0) { s_red[warp][2*j] = a0[j]; s_red[warp][2*j + 1] = a1[j]; }
        }
        __syncthreads();
        if (threadIdx.x < 2 * EVAL_QCHUNK && q0 + (int)threadIdx.x / 2 < nq) {
            double acc = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) acc += s_red[w][threadIdx.x];
            // [block][nq][2]
            partials[((int64_t)blockIdx.x * nq + q0) * 2 + threadIdx.x] = acc;
        }
        __syncthreads();
    }
}

__global__ void fold_sum_kernel(const double* partials, int64_t n_blocks, int64_t len, double* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double acc = 0.0;
    for (int64_t b = 0; b < n_blocks; ++b) acc += partials[b * len + i];
    out[i] = acc;
}

__global__ void __launch_bounds__(256)
path_interp_kernel(const double* lo, const double* hi, double t_lo /* w_lo */,
                   double dt_knots /* w_hi */, double t,
                   int64_t n_paths, double* out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_paths;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = row_value(lo, hi, i, t_lo, dt_knots, t, 1);
}

static const int kEvalBlocks = 592;

extern "C" int64_t sdeb_path_eval_workspace(int64_t nq) { return (int64_t)kEvalBlocks * nq * 2 * 8; }

extern "C" int sdeb_path_cdf(const double* y_lo, const double* y_hi, double w_lo, double w_hi,
                             int64_t interp, int64_t n_paths, const double* q,
                             int64_t nq, int64_t* counts, void* stream_) {
    if (!y_lo || (interp && !y_hi) || !q || !counts || n_paths < 1 || nq < 1)
        return fail(SDEB_EINVAL, "sdeb_path_cdf: bad arguments");
    int64_t need = (n_paths + 1023) / 1024;
    int blocks = (int)(need < 1184 ? need : 1184);
    for (int64_t q0 = 0; q0 < nq; q0 += CDF_MAX_Q) {       // one pass per 2048 thresholds
        const int m = (int)(nq - q0 < CDF_MAX_Q ? nq - q0 : CDF_MAX_Q);
        size_t smem = (size_t)m * 12 + (size_t)8 * (m + 1) * 4;
        if (smem > 48 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(path_cdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
        path_cdf_kernel<<<blocks, 256, smem, (cudaStream_t)stream_>>>(
            y_lo, y_hi, w_lo, w_hi, (int)interp, n_paths, q + q0, m,
            (unsigned long long*)counts + q0);
        CUDA_TRY(cudaGetLastError());
    }
    return SDEB_OK;
}

extern "C" int sdeb_path_chf(const double* y_lo, const double* y_hi, double w_lo, double w_hi,
                             int64_t interp, int64_t n_paths, const double* u,
                             int64_t nq, double* sums, void* workspace, int64_t workspace_bytes,
                             void* stream_) {
    if (!y_lo || (interp && !y_hi) || !u || !sums || n_paths < 1 || nq < 1)
        return fail(SDEB_EINVAL, "sdeb_path_chf: bad arguments");
    if (!workspace || workspace_bytes < sdeb_path_eval_workspace(nq))
        return fail(SDEB_EINVAL, "sdeb_path_chf: workspace too small");
    cudaStream_t stream = (cudaStream_t)stream_;
    int64_t need = (n_paths + 1023) / 1024;
    int blocks = (int)(need < kEvalBlocks ? need : kEvalBlocks);
    path_chf_kernel<<<blocks, 256, 0, stream>>>(y_lo, y_hi, w_lo, w_hi, (int)interp, n_paths, u,
                                                (int)nq, (double*)workspace);
    CUDA_TRY(cudaGetLastError());
    int64_t len = nq * 2;
    fold_sum_kernel<<<(unsigned)((len + 127) / 128), 128, 0, stream>>>(
        (const double*)workspace, blocks, len, sums);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

extern "C" int sdeb_path_interp(const double* y_lo, const double* y_hi, double w_lo, double w_hi,
                                int64_t n_paths, double* out, void* stream_) {
    if (!y_lo || !y_hi || !out || n_paths < 1)
        return fail(SDEB_EINVAL, "sdeb_path_interp: bad arguments");
    int64_t need = (n_paths + 255) / 256;
    int blocks = (int)(need < 1184 ? need : 1184);
    path_interp_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(y_lo, y_hi, w_lo, w_hi, 0.0,
                                                                  n_paths, out);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

// ---------------------------------------------------------------------------
// reductions and scans along an inner axis of a resident slab
// x is [outer][R][cols] (cols contiguous): process.tmin/tmax/tsum/tmean/tvar
// (outer = 1, R = time points, cols = values x paths; infrastructure.py:
// 963-1030) and vmin..vvar (outer = time points, R = values, cols = paths;
// 894-960).  One thread owns one (outer, col) column and walks R in order:
// NumPy reduces a non-contiguous axis sequentially, so sums are bit-equal.
// HBM-bound: 8 independent loads in flight per thread.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double np_min(double a, double v) { return (v < a || v != v) ? v : a; }
__device__ __forceinline__ double np_max(double a, double v) { return (v > a || v != v) ? v : a; }

__global__ void __launch_bounds__(256)
axis_reduce_kernel(const double* __restrict__ x, int64_t outer, int64_t R, int64_t cols,
                   double* omin, double* omax, double* osum, double* ossd,
                   double sum_div, double ssd_div, int ssd_sqrt) {
    const int64_t total = outer * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = i / cols, c = i - o * cols;
        const double* col = x + o * R * cols + c;
        double first = col[0];
        double mn = first, mx = first, sum = first;
        int64_t r = 1;
        for (; r + 8 <= R; r += 8) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = col[(r + k) * cols];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                mn = np_min(mn, v[k]); mx = np_max(mx, v[k]); sum = __dadd_rn(sum, v[k]);
            }
        }
        for (; r < R; ++r) {
            double v = col[r * cols];
            mn = np_min(mn, v); mx = np_max(mx, v); sum = __dadd_rn(sum, v);
        }
        if (omin) omin[i] = mn;
        if (omax) omax[i] = mx;
        if (osum) osum[i] = __ddiv_rn(sum, sum_div);
        if (ossd) {
            // numpy.var: mean = sum/R, then add.reduce((x - mean)*(x - mean))
            const double mean = __ddiv_rn(sum, (double)R);
            double d = __dsub_rn(first, mean);
            double acc = __dmul_rn(d, d);
            r = 1;
            for (; r + 8 <= R; r += 8) {
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = col[(r + k) * cols];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    d = __dsub_rn(v[k], mean);
                    acc = __dadd_rn(acc, __dmul_rn(d, d));
                }
            }
            for (; r < R; ++r) {
                d = __dsub_rn(col[r * cols], mean);
                acc = __dadd_rn(acc, __dmul_rn(d, d));
            }
            acc = __ddiv_rn(acc, ssd_div);
            ossd[i] = ssd_sqrt ? __dsqrt_rn(acc) : acc;
        }
    }
}

extern "C" int sdeb_axis_reduce(const double* x, int64_t outer, int64_t n_reduce, int64_t cols,
                                double* out_min, double* out_max, double* out_sum,
                                double* out_ssd, double sum_div, double ssd_div,
                                int64_t ssd_sqrt, void* stream_) {
    if (!x || outer < 1 || n_reduce < 1 || cols < 1 ||
        !(out_min || out_max || out_sum || out_ssd))
        return fail(SDEB_EINVAL, "sdeb_axis_reduce: bad arguments");
    int64_t need = (outer * cols + 255) / 256;
    int blocks = (int)(need < 2368 ? need : 2368);
    axis_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(
        x, outer, n_reduce, cols, out_min, out_max, out_sum, out_ssd, sum_div, ssd_div,
        (int)ssd_sqrt);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

// process.tcumsum / tint / tdiff (infrastructure.py:990-997, 1032-1122) over
// x [rows][cols]; w holds one weight per time interval (device, rows-1).
//   SCAN_CUMSUM : out[r] = x[0] + ... + x[r]
//   SCAN_INT    : out[0] = 0, out[r] = out[r-1] + x[r-1]*w[r-1]
//   SCAN_DIFF   : out[r] = (x[r+1] - x[r]) (/ w[r] when w != NULL), rows-1 rows
enum { SCAN_CUMSUM = 0, SCAN_INT = 1, SCAN_DIFF = 2 };

template <int MODE>
__global__ void __launch_bounds__(256)
time_scan_kernel(const double* __restrict__ x, int64_t rows, int64_t cols,
                 const double* __restrict__ w, double* __restrict__ out) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cols;
         c += (int64_t)gridDim.x * blockDim.x) {
        const double* col = x + c;
        double* o = out + c;
        if (MODE == SCAN_DIFF) {
            double prev = col[0];
            int64_t r = 1;
            for (; r + 8 <= rows; r += 8) {
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = col[(r + k) * cols];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    double d = __dsub_rn(v[k], prev);
                    if (w) d = __ddiv_rn(d, w[r + k - 1]);
                    o[(r + k - 1) * cols] = d;
                    prev = v[k];
                }
            }
            for (; r < rows; ++r) {
                double v = col[r * cols];
                double d = __dsub_rn(v, prev);
                if (w) d = __ddiv_rn(d, w[r - 1]);
                o[(r - 1) * cols] = d;
                prev = v;
            }
        } else {
            double acc = 0.0;
            int64_t r = 0;
            if (MODE == SCAN_CUMSUM) { acc = col[0]; o[0] = acc; r = 1; }
            else { o[0] = 0.0; }
            // SCAN_INT consumes x[r] to produce out[r+1]
            const int64_t last = MODE == SCAN_INT ? rows - 1 : rows;
            for (; r + 8 <= last; r += 8) {
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = col[(r + k) * cols];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (MODE == SCAN_INT) {
                        acc = __dadd_rn(acc, __dmul_rn(v[k], w[r + k]));
                        o[(r + k + 1) * cols] = acc;
                    } else {
                        acc = __dadd_rn(acc, v[k]);
                        o[(r + k) * cols] = acc;
                    }
                }
            }
            for (; r < last; ++r) {
                double v = col[r * cols];
                if (MODE == SCAN_INT) {
                    acc = __dadd_rn(acc, __dmul_rn(v, w[r]));
                    o[(r + 1) * cols] = acc;
                } else {
                    acc = __dadd_rn(acc, v);
                    o[r * cols] = acc;
                }
            }
        }
    }
}

extern "C" int sdeb_time_scan(const double* x, int64_t rows, int64_t cols, int64_t mode,
                              const double* weights, double* out, void* stream_) {
    if (!x || !out || rows < 1 || cols < 1 || mode < SCAN_CUMSUM || mode > SCAN_DIFF)
        return fail(SDEB_EINVAL, "sdeb_time_scan: bad arguments");
    if (mode == SCAN_INT && rows > 1 && !weights)
        return fail(SDEB_EINVAL, "sdeb_time_scan: SCAN_INT needs the interval weights");
    if (mode == SCAN_DIFF && rows < 2) return SDEB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    int64_t need = (cols + 255) / 256;
    int blocks = (int)(need < 2368 ? need : 2368);
    if (mode == SCAN_CUMSUM)
        time_scan_kernel<SCAN_CUMSUM><<<blocks, 256, 0, stream>>>(x, rows, cols, weights, out);
    else if (mode == SCAN_INT)
        time_scan_kernel<SCAN_INT><<<blocks, 256, 0, stream>>>(x, rows, cols, weights, out);
    else
        time_scan_kernel<SCAN_DIFF><<<blocks, 256, 0, stream>>>(x, rows, cols, weights, out);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

// ---------------------------------------------------------------------------
// histogram with numpy.histogram bin semantics
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
histogram_kernel(const double* x, int64_t n, const double* edges, int nbins, int uniform,
                 unsigned long long* counts, unsigned long long* outside) {
    extern __shared__ double sh[];
    double* s_edges = sh;                                   // nbins + 1
    unsigned int* s_cnt = (unsigned int*)(sh + nbins + 1);  // nbins + 1 (last: outside)
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x) { s_edges[i] = edges[i]; s_cnt[i] = 0; }
    __syncthreads();
    const double lo = s_edges[0], hi = s_edges[nbins];
    const double scale = nbins / (hi - lo);
    // each block handles a contiguous span so that 32-bit bin counters cannot overflow
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        int idx;
        if (!(v >= lo && v <= hi)) {
            idx = nbins;                                    // outside (or NaN)
        } else {
            if (uniform) {
                idx = (int)((v - lo) * scale);
                idx = idx < 0 ? 0 : (idx > nbins - 1 ? nbins - 1 : idx);
            } else {
                int a = 0, b = nbins;                       // largest idx with edges[idx] <= v
                while (b - a > 1) { int m = (a + b) >> 1; if (s_edges[m] <= v) a = m; else b = m; }
                idx = a;
            }
            while (idx > 0 && v < s_edges[idx]) --idx;
            while (idx < nbins - 1 && v >= s_edges[idx + 1]) ++idx;
        }
        atomicAdd(&s_cnt[idx], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x) {
        unsigned int c = s_cnt[i];
        if (c) atomicAdd(i < nbins ? &counts[i] : outside, (unsigned long long)c);
    }
}

extern "C" int sdeb_histogram(const double* x, int64_t n, const double* edges, int64_t nbins,
                              int64_t uniform_edges, int64_t* counts, int64_t* outside,
                              void* stream_) {
    if (!x || !edges || !counts || !outside || nbins < 1 || n < 0)
        return fail(SDEB_EINVAL, "sdeb_histogram: bad arguments");
    if (nbins > 4000) return fail(SDEB_EINVAL, "sdeb_histogram: at most 4000 bins");
    if (n == 0) return SDEB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    int64_t need = (n + 255) / 256;
    // >= n / 2^31 blocks keeps the 32-bit shared counters from overflowing
    int blocks = (int)(need < 1184 ? need : 1184);
    size_t smem = (size_t)(nbins + 1) * 12;
    histogram_kernel<<<blocks, 256, smem, stream>>>(x, n, edges, (int)nbins, (int)uniform_edges,
                                                    (unsigned long long*)counts,
                                                    (unsigned long long*)outside);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

// ---------------------------------------------------------------------------
// standalone source draws
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
draw_wiener_kernel(const NrmK nk, double* out, int n_groups, int ndw, int64_t n_paths, int64_t pitch,
                   int64_t path_offset, u64 seed, u32 step, double sq, const double* chol) {
    __shared__ __align__(16) double tab_mem[TAB_DOUBLES];
    fill_tables(tab_mem);
    const Tab tab(tab_mem);
    __syncthreads();
    const int64_t path = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (path >= n_paths) return;
    const u64 gpath = (u64)(path_offset + path);
    u32 rk[20];
    philox_round_keys(seed, rk);
    Rng rng;
    rng.rk = rk;
    rng.c_x = (u32)gpath; rng.c_y = ((u32)(gpath >> 32) & 0xFFu) | ((u32)g << 8);
    rng.step = step;
    double z[36];
    for (int b = 0; b < (ndw + 3) / 4; ++b) {      // one block = two pairs
        U4 w = rng.block((u32)b);
        normal_pair(w.x, w.y, tab, nk, 1.0, z[4*b], z[4*b + 1]);
        normal_pair(w.z, w.w, tab, nk, 1.0, z[4*b + 2], z[4*b + 3]);
    }
    for (int r = ndw - 1; r >= 0; --r) {
        double acc = z[r];
        if (chol && r > 0) {
            acc = 0.0;
            for (int c = 0; c <= r; ++c) acc = fma(chol[r*(r+1)/2 + c], z[c], acc);
        }
        out[((int64_t)g * ndw + r) * pitch + path] = xmul(acc, sq);
        z[r] = acc;
    }
}

extern "C" int sdeb_draw_wiener(double* out, int64_t n_groups, int64_t ndw, int64_t n_paths,
                                int64_t pitch, int64_t path_offset, uint64_t seed, int64_t step,
                                double sqrt_abs_dt, const double* chol, void* stream_) {
    if (!out || n_groups < 1 || n_groups > 65535 || ndw < 1 || ndw > 32 || n_paths < 1 ||
        pitch < n_paths)
        return fail(SDEB_EINVAL, "sdeb_draw_wiener: bad arguments (ndw <= 32, n_groups <= 65535)");
    cudaStream_t stream = (cudaStream_t)stream_;
    dim3 grid((unsigned)((n_paths + 255) / 256), (unsigned)n_groups);
    static const NrmK nk = {SDEB_NRMK_VALUES};
    draw_wiener_kernel<<<grid, 256, 0, stream>>>(nk, out, (int)n_groups, (int)ndw, n_paths, pitch,
                                                 path_offset, seed, (u32)step, sqrt_abs_dt, chol);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

// Brownian bridge / extension step of a Wiener source with memory
// (true_wiener_source.new_outside / new_inside, infrastructure.py:2460-2499):
//   out = M1 w1 + M2 w2 + Ly z,   z iid N(0,1) from Philox (counter = path, group, step)
// with ndw x ndw matrices mats[3][ndw][ndw] (row-major; M2 ignored when w2 == NULL).
__global__ void __launch_bounds__(256)
bridge_wiener_kernel(const NrmK nk, double* out, const double* w1, const double* w2,
                     const double* mats, int ndw, int64_t n_paths, int64_t pitch,
                     int64_t path_offset, u64 seed, u32 step) {
    __shared__ __align__(16) double tab_mem[TAB_DOUBLES];
    __shared__ double s_m[3 * 32 * 32];
    fill_tables(tab_mem);
    const Tab tab(tab_mem);
    for (int i = threadIdx.x; i < 3 * ndw * ndw; i += blockDim.x) s_m[i] = mats[i];
    __syncthreads();
    const int64_t path = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (path >= n_paths) return;
    const u64 gpath = (u64)(path_offset + path);
    u32 rk[20];
    philox_round_keys(seed, rk);
    Rng rng;
    rng.rk = rk;
    rng.c_x = (u32)gpath; rng.c_y = ((u32)(gpath >> 32) & 0xFFu) | ((u32)g << 8);
    rng.step = step;
    double z[36];
    for (int b = 0; b < (ndw + 3) / 4; ++b) {      // one block = two pairs
        U4 w = rng.block((u32)b);
        normal_pair(w.x, w.y, tab, nk, 1.0, z[4*b], z[4*b + 1]);
        normal_pair(w.z, w.w, tab, nk, 1.0, z[4*b + 2], z[4*b + 3]);
    }
    const double* M1 = s_m; const double* M2 = s_m + ndw * ndw; const double* Ly = s_m + 2 * ndw * ndw;
    for (int r = 0; r < ndw; ++r) {
        double acc = 0.0;
        for (int c = 0; c < ndw; ++c) {
            int64_t at = ((int64_t)g * ndw + c) * pitch + path;
            acc += M1[r * ndw + c] * w1[at];
            if (w2) acc += M2[r * ndw + c] * w2[at];
            acc += Ly[r * ndw + c] * z[c];
        }
        out[((int64_t)g * ndw + r) * pitch + path] = acc;
    }
}

extern "C" int sdeb_bridge_wiener(double* out, const double* w1, const double* w2,
                                  const double* mats, int64_t n_groups, int64_t ndw,
                                  int64_t n_paths, int64_t pitch, int64_t path_offset,
                                  uint64_t seed, int64_t step, void* stream_) {
    if (!out || !w1 || !mats || n_groups < 1 || n_groups > 65535 || ndw < 1 || ndw > 32 ||
        n_paths < 1 || pitch < n_paths)
        return fail(SDEB_EINVAL, "sdeb_bridge_wiener: bad arguments (ndw <= 32)");
    static const NrmK nk = {SDEB_NRMK_VALUES};
    dim3 grid((unsigned)((n_paths + 255) / 256), (unsigned)n_groups);
    bridge_wiener_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(
        nk, out, w1, w2, mats, (int)ndw, n_paths, pitch, path_offset, seed, (u32)step);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

__global__ void __launch_bounds__(256)
draw_cpoisson_kernel(const NrmK nk, double* dj, i64* dn, int64_t n_paths, int64_t pitch, int64_t path_offset,
                     u64 seed, u32 step, double lamdt, double explam, int sign, int law,
                     double a, double b, double pa) {
    __shared__ __align__(16) double tab_mem[TAB_DOUBLES];
    fill_tables(tab_mem);
    const Tab tab(tab_mem);
    __syncthreads();
    const int64_t path = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (path >= n_paths) return;
    const u64 gpath = (u64)(path_offset + path);
    u32 rk[20];
    philox_round_keys(seed, rk);
    Rng rng;
    rng.rk = rk;
    rng.c_x = (u32)gpath; rng.c_y = ((u32)(gpath >> 32) & 0xFFu) | ((u32)g << 8);
    rng.step = step;
    U4 w = rng.block((u32)STREAM_POISSON);
    int k = poisson_draw(u01(w.x, w.y), lamdt, rng, 0u);
    double sum = 0.0;
    for (int j = 0; j < k; ++j) {
        U4 wj = rng.block((u32)(STREAM_JUMP + (j & 0x3FFF)));
        double yj = jump_size(wj, tab, nk, law, a, b, pa);
        sum = (j == 0) ? yj : sum + yj;
    }
    if (dj) dj[(int64_t)g * pitch + path] = sign * sum;
    if (dn) dn[(int64_t)g * pitch + path] = (i64)sign * k;
}

extern "C" int sdeb_draw_cpoisson(double* dj, int64_t* dn, int64_t n_lanes, int64_t n_paths,
                                  int64_t pitch, int64_t path_offset, uint64_t seed,
                                  int64_t step, double lam_abs_dt, int64_t sign, int64_t law,
                                  double a, double b, double pa, void* stream_) {
    if ((!dj && !dn) || n_lanes < 1 || n_lanes > 65535 || n_paths < 1 || pitch < n_paths ||
        lam_abs_dt < 0)
        return fail(SDEB_EINVAL, "sdeb_draw_cpoisson: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    dim3 grid((unsigned)((n_paths + 255) / 256), (unsigned)n_lanes);
    static const NrmK nk = {SDEB_NRMK_VALUES};
    draw_cpoisson_kernel<<<grid, 256, 0, stream>>>(nk, dj, (i64*)dn, n_paths, pitch, path_offset, seed,
                                                   (u32)step, lam_abs_dt, exp(-lam_abs_dt),
                                                   (int)sign, (int)law, a, b, pa);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

// ---------------------------------------------------------------------------
// self tests / measurement
// ---------------------------------------------------------------------------
__global__ void test_normals_kernel(const NrmK nk, u64 seed, int64_t n, double* zf, double* zl) {
    __shared__ __align__(16) double tab_mem[TAB_DOUBLES];
    fill_tables(tab_mem);
    const Tab tab(tab_mem);
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 rk[20];
    philox_round_keys(seed, rk);
    Rng rng;
    rng.rk = rk;
    rng.c_x = (u32)i; rng.c_y = (u32)(i >> 32); rng.step = 0;
    U4 w = rng.block(0);
    const bool full = i >= n / 2;      // second half: the 96-bit full-resolution map
    if ((i % (n / 2 > 0 ? n / 2 : 1)) < 128) {   // force the extreme corners of the bit space through both maps
        if (i & 1) { w.x = 0; w.z = 0; }                          // smallest u: deepest tail
        if (i & 2) { w.x = 0xFFFFFFFFu; w.z = 0xFFFFFFFFu; }      // largest u
        if (i & 4) { w.x &= 0xFFF00000u; w.z &= 0x00000FFFu; }    // top mantissa bits clear
        if (i & 8) w.x |= 0x00080000u;                            // u >= 1/2: e = 1
        if (i & 16) w.y |= 0x00FFFFFFu;                           // largest offset in the sector
        if (i & 32) w.y &= 0xFF000000u;                           // most negative offset
        if (i & 64) w.y |= 0xFF000000u;                           // last sector
    }
    if (full) normal_pair_full(w.x, w.z, w.y, tab, nk, 1.0, zf[2*i], zf[2*i + 1]);
    else normal_pair(w.x, w.y, tab, nk, 1.0, zf[2*i], zf[2*i + 1]);
    normal_pair_libdevice(w.x, w.z, w.y, full, zl[2*i], zl[2*i + 1]);
}

extern "C" int sdeb_test_normals(uint64_t seed, int64_t n, double* z_fast, double* z_libdevice,
                                 void* stream_) {
    if (!z_fast || !z_libdevice || n < 1) return fail(SDEB_EINVAL, "sdeb_test_normals: bad arguments");
    static const NrmK nk = {SDEB_NRMK_VALUES};
    test_normals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
        nk, seed, n, z_fast, z_libdevice);
    CUDA_TRY(cudaGetLastError());
    return SDEB_OK;
}

// host-side Philox (same code path as the device one) for known-answer tests
extern "C" int sdeb_test_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    return SDEB_OK;
}

__global__ void __launch_bounds__(256)
fp64_peak_kernel(int64_t iters, double* sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int64_t i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) sink[0] = s;
}

extern "C" int sdeb_fp64_peak(int64_t iters, double* dfma_per_second, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sm = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    double* sink = NULL;
    CUDA_TRY(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    int blocks = sm * 8;
    fp64_peak_kernel<<<blocks, 256, 0, stream>>>(iters / 8 + 1, sink);   // warm-up
    CUDA_TRY(cudaEventRecord(e0, stream));
    fp64_peak_kernel<<<blocks, 256, 0, stream>>>(iters, sink);
    CUDA_TRY(cudaEventRecord(e1, stream));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    if (dfma_per_second) *dfma_per_second = (double)blocks * 256.0 * 8.0 * (double)iters / (ms * 1e-3);
    return SDEB_OK;
}

// ---------------------------------------------------------------------------
// NVRTC JIT (libnvrtc is dlopen'ed on first use; module loading goes through
// the runtime's cudaLibrary API so that libcuda is never a link dependency)
// ---------------------------------------------------------------------------
typedef int nvrtcResult_t;
typedef struct _nvrtcProgram* nvrtcProgram_t;
struct NvrtcApi {
    void* h;
    nvrtcResult_t (*CreateProgram)(nvrtcProgram_t*, const char*, const char*, int,
                                   const char* const*, const char* const*);
    nvrtcResult_t (*CompileProgram)(nvrtcProgram_t, int, const char* const*);
    nvrtcResult_t (*GetProgramLogSize)(nvrtcProgram_t, size_t*);
    nvrtcResult_t (*GetProgramLog)(nvrtcProgram_t, char*);
    nvrtcResult_t (*GetCUBINSize)(nvrtcProgram_t, size_t*);
    nvrtcResult_t (*GetCUBIN)(nvrtcProgram_t, char*);
    nvrtcResult_t (*DestroyProgram)(nvrtcProgram_t*);
    const char* (*GetErrorString)(nvrtcResult_t);
};
static NvrtcApi g_nvrtc;

static int load_nvrtc() {
    if (g_nvrtc.h) return SDEB_OK;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so",
                           "/usr/local/cuda/lib64/libnvrtc.so.12", NULL};
    void* h = NULL;
    for (int i = 0; names[i] && !h; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(SDEB_EJIT, std::string("cannot dlopen libnvrtc: ") + dlerror());
#define NVRTC_SYM(field, name)                                              \
    *(void**)(&g_nvrtc.field) = dlsym(h, name);                             \
    if (!g_nvrtc.field) return fail(SDEB_EJIT, "libnvrtc lacks " name);
    NVRTC_SYM(CreateProgram, "nvrtcCreateProgram")
    NVRTC_SYM(CompileProgram, "nvrtcCompileProgram")
    NVRTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    NVRTC_SYM(GetProgramLog, "nvrtcGetProgramLog")
    NVRTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    NVRTC_SYM(GetCUBIN, "nvrtcGetCUBIN")
    NVRTC_SYM(DestroyProgram, "nvrtcDestroyProgram")
    NVRTC_SYM(GetErrorString, "nvrtcGetErrorString")
#undef NVRTC_SYM
    g_nvrtc.h = h;
    return SDEB_OK;
}

static int nvrtc_cubin(const std::string& source, const std::string& name,
                       const std::vector<std::string>& defines, std::vector<char>& cubin,
                       std::string& plog) {
    int rc = load_nvrtc();
    if (rc) return rc;
    nvrtcProgram_t prog;
    nvrtcResult_t r = g_nvrtc.CreateProgram(&prog, source.c_str(), name.c_str(), 0, NULL, NULL);
    if (r) return fail(SDEB_EJIT, std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r));
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17",
                                     "-default-device"};
    for (const std::string& d : defines) opts.push_back(d.c_str());
    r = g_nvrtc.CompileProgram(prog, (int)opts.size(), opts.data());
    size_t lsz = 0;
    g_nvrtc.GetProgramLogSize(prog, &lsz);
    plog.assign(lsz, 0);
    if (lsz > 1) g_nvrtc.GetProgramLog(prog, &plog[0]);
    if (r) {
        g_nvrtc.DestroyProgram(&prog);
        return fail(SDEB_EJIT, std::string("nvrtcCompileProgram: ") + g_nvrtc.GetErrorString(r) +
                    "\n" + plog);
    }
    size_t csz = 0;
    g_nvrtc.GetCUBINSize(prog, &csz);
    cubin.resize(csz);
    g_nvrtc.GetCUBIN(prog, cubin.data());
    g_nvrtc.DestroyProgram(&prog);
    return SDEB_OK;
}

// `source` = engine header + model definition + the entry wrappers
//   extern "C" __global__ void sdeb_jit_entry(const sdeb::KArgs a)       (#ifndef SDEB_JIT_NO_GENERAL)
//   extern "C" __global__ void sdeb_jit_entry_lean(const sdeb::KArgs a)  (#ifndef SDEB_JIT_NO_LEAN)
// plus  extern "C" __constant__ int sdeb_jit_dims[6] = {NW,NDW,NX,NPC,NCNT,JUMPS};
// (the Python side generates all of it, sdepy_b200/_jit.py); `model_type` is only
// used as the NVRTC program name.  Compiled here: the lean entry and the
// dimensions.  The general entry is compiled by the first sdeb_integrate that
// needs it, one sweep variant (noise mode x record residency) at a time.
extern "C" int sdeb_jit_compile(const char* source, const char* model_type, int64_t* handle,
                                char* log, int64_t log_bytes) {
    if (!source || !handle) return fail(SDEB_EINVAL, "sdeb_jit_compile: null argument");
    if (log && log_bytes > 0) log[0] = 0;
    JitModule jm;
    jm.source = source;
    jm.name = model_type ? model_type : "sdeb_jit.cu";
    for (int k = 0; k < 10; ++k) jm.ghave[k] = false;
    std::vector<char> cubin;
    std::string plog;
    // (this stage holds the lean entry only: contracted arithmetic, SDEB_CONTRACT)
    int rc = nvrtc_cubin(jm.source, jm.name, {"-DSDEB_JIT_NO_GENERAL", "-DSDEB_CONTRACT=1"}, cubin, plog);
    if (log && log_bytes > 0) { strncpy(log, plog.c_str(), (size_t)log_bytes - 1); log[log_bytes - 1] = 0; }
    if (rc) return rc;
    CUDA_TRY(cudaLibraryLoadData(&jm.lib, cubin.data(), NULL, NULL, 0, NULL, NULL, 0));
    void* dptr = NULL;
    size_t dbytes = 0;
    CUDA_TRY(cudaLibraryGetGlobal(&dptr, &dbytes, jm.lib, "sdeb_jit_dims"));
    int dims[6];
    if (dbytes < sizeof dims) return fail(SDEB_EJIT, "sdeb_jit_dims has the wrong size");
    CUDA_TRY(cudaMemcpy(dims, dptr, sizeof dims, cudaMemcpyDeviceToHost));
    jm.mi.fn = NULL;
    jm.mi.fn_lean = NULL;
    for (int k = 0; k < 4; ++k) jm.mi.fn_stream[k] = NULL;
    if (cudaLibraryGetKernel(&jm.kernel_lean, jm.lib, "sdeb_jit_entry_lean") == cudaSuccess &&
        dims[3] + (dims[1] > 1 ? dims[1] * (dims[1] + 1) / 2 : 0) <= MAX_CBANK_PARAMS)
        jm.mi.fn_lean = (const void*)jm.kernel_lean;
    else
        cudaGetLastError();
    jm.mi.nw = dims[0]; jm.mi.ndw = dims[1]; jm.mi.nx = dims[2]; jm.mi.npc = dims[3];
    jm.mi.ncnt = dims[4]; jm.mi.jumps = dims[5];
    std::lock_guard<std::mutex> lock(g_jit_mutex);
    *handle = g_jit_next++;
    g_jit[*handle] = jm;
    return SDEB_OK;
}

static int jit_general_kernel(int64_t handle, int variant, const void** fn) {
    std::lock_guard<std::mutex> lock(g_jit_mutex);
    auto it = g_jit.find(handle);
    if (it == g_jit.end()) return fail(SDEB_EINVAL, "unknown JIT handle");
    JitModule& jm = it->second;
    if (variant < 0 || variant >= 10) return fail(SDEB_EINVAL, "bad sweep variant");
    if (!jm.ghave[variant]) {
        char def[64];
        std::vector<std::string> defs = {"-DSDEB_JIT_NO_LEAN"};
        const char* entry = "sdeb_jit_entry";
        if (variant < 6) {
            snprintf(def, sizeof def, "-DSDEB_SWEEPS=%d", 1 << variant);
        } else {        // stream kernel: variant - 6 = 2*replay + time-dependent records
            defs.push_back("-DSDEB_JIT_NO_GENERAL");
            snprintf(def, sizeof def, "-DSDEB_JIT_STREAM=%d", variant - 6);
            entry = "sdeb_jit_entry_stream";
        }
        defs.push_back(def);
        // the rounding rule (sde_engine.cuh:clamp_tiny): a variant that serves plain
        // Philox runs -- general sweeps 0, 1, stream variants 0, 1 -- is a translation
        // unit of its own and compiles the traced step with plain, contractible operators
        if (variant < 2 || variant == 6 || variant == 7) defs.push_back("-DSDEB_CONTRACT=1");
        std::vector<char> cubin;
        std::string plog;
        int rc = nvrtc_cubin(jm.source, jm.name, defs, cubin, plog);
        if (rc) return rc;
        CUDA_TRY(cudaLibraryLoadData(&jm.glib[variant], cubin.data(), NULL, NULL, 0, NULL, NULL, 0));
        CUDA_TRY(cudaLibraryGetKernel(&jm.gkernel[variant], jm.glib[variant], entry));
        jm.ghave[variant] = true;
    }
    *fn = (const void*)jm.gkernel[variant];
    return SDEB_OK;
}

extern "C" int sdeb_jit_release(int64_t handle) {
    std::lock_guard<std::mutex> lock(g_jit_mutex);
    auto it = g_jit.find(handle);
    if (it == g_jit.end()) return fail(SDEB_EINVAL, "sdeb_jit_release: unknown handle");
    cudaLibraryUnload(it->second.lib);
    for (int k = 0; k < 10; ++k)
        if (it->second.ghave[k]) cudaLibraryUnload(it->second.glib[k]);
    g_jit.erase(it);
    return SDEB_OK;
}
