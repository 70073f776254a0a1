/*
 * sdeb.h -- C ABI of libsdeb.so, the B200 (sm_100a) SDE path-integration
 * library behind the sdepy-compatible Python surface in sdepy_b200/.
 *
 * The reference (sdepy, pure Python) has no FFI; its extension surface is a
 * set of cooperating Python classes.  Each entry point below replaces the
 * reference code cited next to it (paths relative to /root/reference) and is
 * what a maintainer would bind from sdepy itself (ctypes stub: INTEGRATION.md).
 *
 * Conventions
 *  - plain C: fixed-width integers, doubles and raw DEVICE pointers only; no
 *    torch / C++ types.  Every struct field is 8 bytes wide (no padding).
 *  - the library never allocates or frees caller memory; scratch space is
 *    passed in (`workspace`), its size comes from sdeb_plan().
 *  - all calls are asynchronous on the given CUDA stream (a cudaStream_t cast
 *    to void*; NULL = legacy default stream) and re-entrant per stream.
 *  - return value 0 = success, non-zero = error; sdeb_last_error() returns a
 *    thread-local message.  No exceptions cross the ABI, no CPU fallback:
 *    without a CUDA device every compute entry point fails with SDEB_ENODEV.
 */
#ifndef SDEB_H
#define SDEB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDEB_ABI_VERSION 2

enum {
    SDEB_OK = 0,
    SDEB_EINVAL = 1,     /* bad argument / unsupported shape              */
    SDEB_ENODEV = 2,     /* no usable CUDA device                         */
    SDEB_ECUDA = 3,      /* CUDA runtime error (message has the detail)   */
    SDEB_EJIT = 4        /* NVRTC compilation / module load failure       */
};

/* preset models = the reference's preset `sde` bodies */
enum {
    SDEB_MODEL_LINEAR = 1,        /* wiener_SDE  integration.py:2069        */
    SDEB_MODEL_LINEAR_LOG = 2,    /* lognorm_SDE integration.py:2129 (exp at store) */
    SDEB_MODEL_JUMPDIFF = 3,      /* jumpdiff_SDE integration.py:2594-2608  */
    SDEB_MODEL_MEANREV = 4,       /* ornstein_uhlenbeck_SDE integration.py:2198 */
    SDEB_MODEL_HULL_WHITE = 5,    /* hull_white_SDE integration.py:2268-2272 */
    SDEB_MODEL_CIR = 6,           /* cox_ingersoll_ross_SDE integration.py:2351 */
    SDEB_MODEL_HESTON = 7,        /* heston_SDE integration.py:2528-2541    */
    SDEB_MODEL_HESTON_FULL = 8,   /* full_heston_SDE integration.py:2416-2444 */
    SDEB_MODEL_JIT = 100          /* NVRTC-compiled model, see sdeb_jit_compile */
};

enum { SDEB_NOISE_PHILOX = 0, SDEB_NOISE_REPLAY = 1 };
enum { SDEB_LAW_NORMAL = 1, SDEB_LAW_UNIFORM = 2, SDEB_LAW_EXP = 3, SDEB_LAW_DOUBLE_EXP = 4 };
enum { SDEB_PAYOFF_NONE = 0, SDEB_PAYOFF_CALL = 1, SDEB_PAYOFF_PUT = 2 };
enum { SDEB_F64 = 0, SDEB_F32 = 1, SDEB_F16 = 2 };

/* statistics vector per (row, component): centred power sums S1..S4 of
 * (v - centre), min, max, payoff sum, payoff square sum */
#define SDEB_NSTAT 8

/*
 * One integration launch = paths_generator._generate_paths
 * (integration.py:392-474) for `n_paths` paths over `n_steps` steps, including
 * the sources' draws (infrastructure.py:1503-1560, 1617-1633, 2017-2040), the
 * Euler update (integration.py:707-723), SDE.store/let (1175-1186, 1528) and
 * the exit transform (1193-1197).  Layouts follow the reference: per-path
 * arrays are [.., component, path] with the path axis contiguous
 * (integration.py:550-555); `pitch` is the allocated length of that axis.
 */
typedef struct sdeb_problem {
    int64_t abi_version;      /* SDEB_ABI_VERSION                               */
    int64_t model;            /* SDEB_MODEL_*                                   */
    int64_t ncomp;            /* per-lane size of the coupled (last working) axis:
                                 factors (hull-white), vshape[-1] (others, N for heston) */
    int64_t jit_handle;       /* SDEB_MODEL_JIT: handle from sdeb_jit_compile   */
    int64_t noise;            /* SDEB_NOISE_*                                   */
    int64_t n_paths;          /* paths integrated by this launch                */
    int64_t path_offset;      /* global index of local path 0 (Philox counter;
                                 makes results independent of sharding)        */
    int64_t pitch;            /* allocated path-axis length of per-path arrays  */
    int64_t n_steps;          /* len(steps_tt) - 1                              */
    int64_t n_groups;         /* independent lane groups = prod(leading working axes) */
    int64_t n_rows;           /* rows of `out` / `stats`                        */
    int64_t row0;             /* row receiving the initial state, or -1         */
    int64_t n_psteps;         /* 1: time-invariant params; n_steps: per-step    */
    int64_t w0_per_path;      /* w0 carries a trailing path axis                */
    int64_t params_per_path;  /* records carry a trailing path axis:
                                 params[n_psteps][n_groups][npt][pitch] (parameters that
                                 vary along the paths axis, e.g. process-valued ones) */
    uint64_t seed;            /* Philox4x32-10 key.  Counter = (global path low 32
                                 bits, path bits 32..39 | group << 8, step word,
                                 stream | jump lane << 16).  Normals: step word =
                                 n / P with P in {1, 2, 4} the number of steps that
                                 consume whole blocks (a block = two Box-Muller pairs
                                 of 64 bits: 32-bit radius uniform + 32-bit angle; one
                                 pair of 96 bits in kernels compiled with
                                 -DSDEB_DRAW_FULL=1), stream = block index; Poisson
                                 counts: step word = even step, stream 0x100 (further
                                 blocks 0x4000 + j when lam|dt| > 30 and the draw is
                                 split into Poisson(lam|dt|/m) terms); jump sizes:
                                 stream 0x200 + j.  Results depend on (seed, global
                                 path, group) only -- not on the sharding, the grid or
                                 the launch geometry or which kernel (sdeb_plan_t.kernel)
                                 runs.  Rounding: replayed increments, and Philox runs
                                 with dW_dump set, round every product and sum of the
                                 step like NumPy; plain Philox runs contract the preset
                                 steps into FMAs (~1e-14 apart over a few hundred steps) */
    const double* steps;      /* [n_steps][2]: dt = t[n+1]-t[n] (integration.py:714),
                                 sqrt|dt| (infrastructure.py:1558-1559)         */
    const int32_t* store_row; /* [n_steps]: row storing the state after step n
                                 (integration.py:386 exact-equality store), -1 = none */
    const double* params;     /* [n_psteps][n_groups][npt] records (+[pitch] when
                                 params_per_path), see sdeb_plan                   */
    const double* params_host;/* optional HOST copy of the same records: a single
                                 time-invariant record (n_psteps == n_groups == 1)
                                 is then passed in the kernel-argument constant
                                 bank (no registers, no loads in the step loop) */
    const double* w0;         /* initial WORKING state (after init/log,
                                 integration.py:1161-1164): [n_groups][nw] (+[pitch]) */
    const double* dW;         /* replay: [n_steps][n_groups*ndw][pitch]         */
    const double* dJ;         /* replay: [n_steps][n_groups*jumps*nw][pitch] (jump slot
                                 s, component c at lane s*nw + c)                  */
    const int64_t* dN;        /* replay: same layout, optional                  */
    double* out;              /* [n_rows][n_groups*nx][pitch], may be NULL; elements
                                 of type out_dtype (float64 unless stated)         */
    double* stats;            /* [n_rows][n_groups*nx][SDEB_NSTAT], may be NULL */
    const double* centre;     /* [n_groups*nx]: shift of the power sums         */
    int64_t payoff_kind;      /* SDEB_PAYOFF_*                                  */
    double payoff_strike;
    double payoff_scale;      /* e.g. discount factor                           */
    int64_t* counter;         /* [n_groups*ncnt][pitch]: += negative_y_count
                                 (integration.py:2435-2439) / jump_count (2618) */
    int64_t* dn_sum;          /* [n_steps]: += sum over lanes of dn (jump_rate,
                                 integration.py:2616-2617), may be NULL         */
    double* dW_dump;          /* philox: write generated dW (layout of dW)      */
    double* dJ_dump;          /* philox: write generated dJ                     */
    int64_t* dN_dump;         /* philox: write generated dN                     */
    void* workspace;          /* >= plan.workspace_bytes when stats != NULL     */
    int64_t workspace_bytes;
    int64_t max_blocks;       /* 0 = auto (persistent grid, multiple of SM count) */
    int64_t anti_dw_half;     /* K > 0: global paths p >= K reuse the Wiener stream of
                                 p - K with the sign reversed (odd_wiener_source,
                                 infrastructure.py:2095-2110); 0 = off          */
    int64_t anti_dj_half;     /* K > 0: paths p >= K repeat the jumps of p - K
                                 (even_cpoisson_source, 2133-2150); 0 = off     */
    int64_t out_dtype;        /* storage type of `out` (paths_generator dtype=,
                                 integration.py:495-496, 550): SDEB_F64 / SDEB_F32 /
                                 SDEB_F16.  The state and all arithmetic stay fp64 in
                                 registers; values are narrowed at the store        */
} sdeb_problem;

typedef struct sdeb_plan_t {
    int64_t nw;               /* working components per lane                    */
    int64_t ndw;              /* Wiener increments per lane per step            */
    int64_t nx;               /* stored components per lane                     */
    int64_t npc;              /* model parameters per record                    */
    int64_t npt;              /* record length = npc + (ndw > 1 ? ndw(ndw+1)/2 : 0):
                                 when ndw > 1 every record ends with the row-major
                                 lower Cholesky factor of corr (identity if none),
                                 used by the Philox draws (infrastructure.py:1532) */
    int64_t ncnt;             /* counters per lane                              */
    int64_t jumps;            /* compound-Poisson terms per component: 0, 1, or 2 (a
                                 traced model with both 'dj' and 'dn' differentials) */
    int64_t blocks;           /* grid size that will be launched                */
    int64_t threads;          /* block size                                     */
    int64_t smem_bytes;       /* dynamic shared memory per block                */
    int64_t workspace_bytes;  /* scratch needed when stats != NULL              */
    int64_t stats_in_kernel;  /* 0: stats accumulators do not fit in smem       */
    int64_t kernel;           /* which kernel runs: 0 general, 1 lean (Philox, one
                                 time-invariant record), 2 stream (full-path output) */
} sdeb_plan_t;

int sdeb_abi_version(void);
const char* sdeb_last_error(void);

/* device query: returns SDEB_ENODEV when no CUDA device is usable */
int sdeb_device_info(int64_t* sm_count, int64_t* cc_major, int64_t* cc_minor,
                     int64_t* total_mem_bytes);

/* shapes, launch geometry and scratch requirement for a problem (host only).
 * SDEB_EINVAL (message naming the sizes) when the per-block shared memory the
 * problem needs -- parameter records staged per step block when time-dependent
 * (n_psteps > 1, not per path), the replay ring, in-kernel statistics, on top of
 * the 72 KB of generator tables and the 32 KB alignment gap in front of them
 * (records up to 28 KB live inside the gap) -- exceeds what an sm_100a block can
 * have (227 KB, ~2 KB of it static). */
int sdeb_plan(const sdeb_problem* p, sdeb_plan_t* plan);

/* the fused integration launch (+ deterministic fold of the per-block
 * statistics partials into p->stats) */
int sdeb_integrate(const sdeb_problem* p, void* stream);

/*
 * Across-path statistics of a device-resident array x[n_rows][pitch]:
 * process.pmean/pvar/pstd (infrastructure.py:861-889) and the power sums of
 * montecarlo._update_moments (2924-2959).  stats[row][SDEB_NSTAT] as above
 * (payoff slots zero).  workspace >= sdeb_moments_workspace(n_rows) bytes.
 */
int64_t sdeb_moments_workspace(int64_t n_rows);
int sdeb_moments(const double* x, int64_t n_rows, int64_t n_paths, int64_t pitch,
                 const double* centre /* [n_rows] device, may be NULL = 0 */,
                 double* stats, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * montecarlo in two HBM passes (montecarlo.update, infrastructure.py:2869-3021).
 *  sdeb_mc_range : pass 1 of the FIRST sample -- stats[row] = {sum, 0, 0, 0, min,
 *      max, 0, 0}: the centring constant (first-sample mean, 2934) and the
 *      histogram range of numpy.histogram(range=None) (2999-3004).
 *  sdeb_mc_update : ONE pass giving the power sums S1..S4 of (x - centre)
 *      (stats[row][0..3], 2940-2946) AND the histogram (counts/outside are
 *      ACCUMULATED, 3010-3013), every row with its own edges[row][nbins+1].
 *      centre == NULL: centre = range_stats[row][0] / n (written to centre_out, if given,
 *      so that later updates cumulate about the same constant).
 *      edges_mode SDEB_MC_EDGES_GIVEN: edges are read; _MINMAX / _RANGE: edges are
 *      BUILT on the device exactly as numpy.linspace(lo, hi, nbins + 1) does
 *      (step = (hi - lo)/nbins, e_i = fl(fl(i*step) + lo), e_nbins = hi; lo == hi
 *      widens by 1/2 each side) from the min/max of range_stats or from
 *      (range_lo, range_hi), and written back to `edges` -- no host round trip
 *      between the two passes.  nbins == 0: moments only.
 *  workspace >= sdeb_moments_workspace(n_rows) bytes for both.
 */
#define SDEB_MC_EDGES_GIVEN  0
#define SDEB_MC_EDGES_MINMAX 1
#define SDEB_MC_EDGES_RANGE  2
int sdeb_mc_range(const double* x, int64_t n_rows, int64_t n, int64_t pitch, double* stats,
                  void* workspace, int64_t workspace_bytes, void* stream);
int sdeb_mc_update(const double* x, int64_t n_rows, int64_t n, int64_t pitch,
                   const double* centre /* [n_rows] device or NULL */,
                   double* centre_out /* [n_rows] device or NULL: the centre used */,
                   const double* range_stats /* [n_rows][SDEB_NSTAT] device or NULL */,
                   double range_lo, double range_hi, int64_t edges_mode,
                   double* edges /* [n_rows][nbins+1] device */, int64_t nbins,
                   int64_t uniform_edges, double* stats /* [n_rows][SDEB_NSTAT] */,
                   int64_t* counts /* [n_rows][nbins] */, int64_t* outside /* [n_rows] */,
                   void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Antithetic halves: out[r][k] = (x[r][k] + sign*x[r][half+k])/2, sign = +1 ('even')
 * or -1 ('odd') -- montecarlo(use=...), infrastructure.py:2905-2914.
 */
int sdeb_antithetic_fold(const double* x, int64_t n_rows, int64_t half, int64_t pitch_in,
                         int64_t pitch_out, int64_t sign, double* out, void* stream);

/*
 * process.cdf / process.chf / process interpolation (infrastructure.py:544-633,
 * 1125-1209) over the n_paths values of ONE row (one time point, one component).
 * With interp != 0 the row is the linear interpolation y = w_hi*y_hi + w_lo*y_lo
 * between two stored rows, w_hi = (t - t_lo)/(t_hi - t_lo), w_lo = (t_hi - t)/(t_hi - t_lo)
 * evaluated by the caller: scipy.interpolate.interp1d's formula with every
 * operation separately rounded; with interp == 0 it is y_lo itself.
 *   sdeb_path_cdf : counts[j] += #{paths : y <= q[j]}            (exact integers)
 *   sdeb_path_chf : sums[j][0..1] = sum over paths of cos(u[j] y), sin(u[j] y)
 *   sdeb_path_interp : out[path] = y
 */
int64_t sdeb_path_eval_workspace(int64_t nq);
int sdeb_path_cdf(const double* y_lo, const double* y_hi, double w_lo, double w_hi,
                  int64_t interp, int64_t n_paths, const double* q, int64_t nq,
                  int64_t* counts, void* stream);
int sdeb_path_chf(const double* y_lo, const double* y_hi, double w_lo, double w_hi,
                  int64_t interp, int64_t n_paths, const double* u, int64_t nq, double* sums,
                  void* workspace, int64_t workspace_bytes, void* stream);
int sdeb_path_interp(const double* y_lo, const double* y_hi, double w_lo, double w_hi,
                     int64_t n_paths, double* out, void* stream);

/*
 * Reductions along an inner axis of a resident slab x[outer][n_reduce][cols]
 * (cols contiguous), one output per (outer, col): process.tmin/tmax/tsum/tmean/
 * tvar/tstd (outer = 1, n_reduce = time points; infrastructure.py:963-1030) and
 * vmin..vstd (outer = time points, n_reduce = values, cols = paths; 894-960).
 * Each output pointer may be NULL.  The sum S accumulates in index order (bit-equal
 * to NumPy's reduction of a non-contiguous axis); out_sum = S / sum_div (1 for the
 * sum, n_reduce for the mean); out_ssd = sum of (x - S/n_reduce)^2, divided by
 * ssd_div (n_reduce - ddof: numpy.var) and square-rooted when ssd_sqrt != 0
 * (numpy.std).  NaNs propagate through min / max as in NumPy.
 */
int sdeb_axis_reduce(const double* x, int64_t outer, int64_t n_reduce, int64_t cols,
                     double* out_min, double* out_max, double* out_sum, double* out_ssd,
                     double sum_div, double ssd_div, int64_t ssd_sqrt, void* stream);

/*
 * Scans along the time axis of x[rows][cols] (infrastructure.py:990-997, 1032-1122):
 *   SDEB_SCAN_CUMSUM  out[r] = x[0] + ... + x[r]                       (tcumsum)
 *   SDEB_SCAN_INT     out[0] = 0, out[r] = out[r-1] + x[r-1]*w[r-1]    (tint, w = dt)
 *   SDEB_SCAN_DIFF    out[r] = (x[r+1] - x[r]) / w[r], rows-1 rows     (tdiff, w = dt**dt_exp
 *                     or NULL for plain increments)
 * weights: device array of rows-1 doubles.
 */
#define SDEB_SCAN_CUMSUM 0
#define SDEB_SCAN_INT    1
#define SDEB_SCAN_DIFF   2
int sdeb_time_scan(const double* x, int64_t rows, int64_t cols, int64_t mode,
                   const double* weights, double* out, void* stream);

/*
 * 1-D histogram of x[n] on given edges[nbins+1] with numpy.histogram
 * semantics (half-open bins, last one closed; montecarlo._update_histogram,
 * infrastructure.py:2961-3021).  counts[nbins] and outside[1] are
 * ACCUMULATED (+=) so that chunked updates cumulate (3010-3013).
 */
int sdeb_histogram(const double* x, int64_t n, const double* edges, int64_t nbins,
                   int64_t uniform_edges, int64_t* counts, int64_t* outside, void* stream);

/* standalone source draws (wiener_source.__call__, infrastructure.py:1503):
 * out[ncomp_total][pitch] = sqrt|dt| * L z, one Philox step index per call */
int sdeb_draw_wiener(double* out, int64_t n_groups, int64_t ndw, int64_t n_paths,
                     int64_t pitch, int64_t path_offset, uint64_t seed, int64_t step,
                     double sqrt_abs_dt, const double* chol /* device, NULL = iid */,
                     void* stream);

/* Wiener source with memory (true_wiener_source.new_outside / new_inside,
 * infrastructure.py:2460-2499): out = M1 w1 + M2 w2 + Ly z with fresh iid normals z;
 * mats = device array [3][ndw][ndw] (M1, M2, Ly row-major), w2 may be NULL. */
int sdeb_bridge_wiener(double* out, const double* w1, const double* w2, const double* mats,
                       int64_t n_groups, int64_t ndw, int64_t n_paths, int64_t pitch,
                       int64_t path_offset, uint64_t seed, int64_t step, void* stream);

/* compound-Poisson draw (cpoisson_source.__call__, infrastructure.py:2017) */
int sdeb_draw_cpoisson(double* dj, int64_t* dn, int64_t n_lanes, int64_t n_paths,
                       int64_t pitch, int64_t path_offset, uint64_t seed, int64_t step,
                       double lam_abs_dt, int64_t sign, int64_t law,
                       double a, double b, double pa, void* stream);

/* self-test / measurement helpers */
int sdeb_test_normals(uint64_t seed, int64_t n, double* z_fast, double* z_libdevice,
                      void* stream);   /* z_*[2n]: hand-rolled vs libdevice map */
int sdeb_test_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int sdeb_fp64_peak(int64_t iters, double* dfma_per_second, void* stream);

/*
 * Run-time compilation of a user SDE (the `integrate` decorator path,
 * integration.py:1843-1955): `model_source` defines `struct UserModel` with
 * the functor interface of sde_engine.cuh; the engine source is prepended by
 * the caller.  Returns a handle for sdeb_problem.jit_handle.
 * Compilation is staged: this call builds the hot-configuration kernel and the
 * model dimensions (-DSDEB_JIT_NO_GENERAL); sdeb_plan / sdeb_integrate compile
 * the general kernel on first use, one sweep variant (noise mode x record
 * residency) at a time (-DSDEB_JIT_NO_LEAN -DSDEB_SWEEPS=<bit>), and keep it
 * with the handle.  A problem with n_paths == 0 only queries the dimensions.
 */
int sdeb_jit_compile(const char* source, const char* model_type,
                     int64_t* handle, char* log, int64_t log_bytes);
int sdeb_jit_release(int64_t handle);

#ifdef __cplusplus
}
#endif
#endif /* SDEB_H */
