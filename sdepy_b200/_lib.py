"""ctypes binding of libsdeb.so (include/sdeb.h).

There is NO CPU fallback: if the CUDA library is missing the import fails,
and without a CUDA device every compute entry point raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SDEB_LIB') or os.path.join(HERE, 'csrc', 'libsdeb.so')

ABI_VERSION = 2
NSTAT = 8
(MODEL_LINEAR, MODEL_LINEAR_LOG, MODEL_JUMPDIFF, MODEL_MEANREV,
 MODEL_HULL_WHITE, MODEL_CIR, MODEL_HESTON, MODEL_HESTON_FULL) = range(1, 9)
MODEL_JIT = 100
NOISE_PHILOX, NOISE_REPLAY = 0, 1
LAW_NORMAL, LAW_UNIFORM, LAW_EXP, LAW_DOUBLE_EXP = 1, 2, 3, 4
PAYOFF_NONE, PAYOFF_CALL, PAYOFF_PUT = 0, 1, 2
F64, F32, F16 = 0, 1, 2

i64, u64, f64, ptr = C.c_int64, C.c_uint64, C.c_double, C.c_void_p


class Problem(C.Structure):
    """struct sdeb_problem (include/sdeb.h) -- every field is 8 bytes."""
    _fields_ = [
        ('abi_version', i64), ('model', i64), ('ncomp', i64),
        ('jit_handle', i64), ('noise', i64), ('n_paths', i64),
        ('path_offset', i64), ('pitch', i64), ('n_steps', i64),
        ('n_groups', i64), ('n_rows', i64), ('row0', i64),
        ('n_psteps', i64), ('w0_per_path', i64), ('params_per_path', i64),
        ('seed', u64),
        ('steps', ptr), ('store_row', ptr), ('params', ptr),
        ('params_host', ptr), ('w0', ptr),
        ('dW', ptr), ('dJ', ptr), ('dN', ptr), ('out', ptr), ('stats', ptr),
        ('centre', ptr),
        ('payoff_kind', i64), ('payoff_strike', f64), ('payoff_scale', f64),
        ('counter', ptr), ('dn_sum', ptr), ('dW_dump', ptr),
        ('dJ_dump', ptr), ('dN_dump', ptr),
        ('workspace', ptr), ('workspace_bytes', i64), ('max_blocks', i64),
        ('anti_dw_half', i64), ('anti_dj_half', i64), ('out_dtype', i64),
    ]


class Plan(C.Structure):
    """struct sdeb_plan_t (include/sdeb.h)."""
    _fields_ = [(k, i64) for k in (
        'nw', 'ndw', 'nx', 'npc', 'npt', 'ncnt', 'jumps', 'blocks',
        'threads', 'smem_bytes', 'workspace_bytes', 'stats_in_kernel', 'kernel')]


class SdebError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'sdepy_b200: the CUDA library %s is missing. Build it with '
            '`python sdepy_b200/_build.py` (needs nvcc); there is no CPU '
            'fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.sdeb_last_error.restype = C.c_char_p
    lib.sdeb_abi_version.restype = C.c_int
    lib.sdeb_moments_workspace.restype = i64
    lib.sdeb_moments_workspace.argtypes = [i64]
    lib.sdeb_device_info.argtypes = [C.POINTER(i64)]*4
    lib.sdeb_plan.argtypes = [C.POINTER(Problem), C.POINTER(Plan)]
    lib.sdeb_integrate.argtypes = [C.POINTER(Problem), ptr]
    lib.sdeb_moments.argtypes = [ptr, i64, i64, i64, ptr, ptr, ptr, i64, ptr]
    lib.sdeb_histogram.argtypes = [ptr, i64, ptr, i64, i64, ptr, ptr, ptr]
    lib.sdeb_mc_range.argtypes = [ptr, i64, i64, i64, ptr, ptr, i64, ptr]
    lib.sdeb_mc_update.argtypes = [ptr, i64, i64, i64, ptr, ptr, ptr, f64, f64, i64, ptr, i64, i64,
                                   ptr, ptr, ptr, ptr, i64, ptr]
    lib.sdeb_path_eval_workspace.restype = i64
    lib.sdeb_path_eval_workspace.argtypes = [i64]
    lib.sdeb_path_cdf.argtypes = [ptr, ptr, f64, f64, i64, i64, ptr, i64, ptr, ptr]
    lib.sdeb_path_chf.argtypes = [ptr, ptr, f64, f64, i64, i64, ptr, i64, ptr, ptr, i64, ptr]
    lib.sdeb_path_interp.argtypes = [ptr, ptr, f64, f64, i64, ptr, ptr]
    lib.sdeb_axis_reduce.argtypes = [ptr, i64, i64, i64, ptr, ptr, ptr, ptr, f64, f64, i64, ptr]
    lib.sdeb_time_scan.argtypes = [ptr, i64, i64, i64, ptr, ptr, ptr]
    lib.sdeb_antithetic_fold.argtypes = [ptr, i64, i64, i64, i64, i64, ptr, ptr]
    lib.sdeb_draw_wiener.argtypes = [ptr, i64, i64, i64, i64, i64, u64, i64,
                                     f64, ptr, ptr]
    lib.sdeb_bridge_wiener.argtypes = [ptr, ptr, ptr, ptr, i64, i64, i64, i64, i64, u64, i64, ptr]
    lib.sdeb_draw_cpoisson.argtypes = [ptr, ptr, i64, i64, i64, i64, u64, i64,
                                       f64, i64, i64, f64, f64, f64, ptr]
    lib.sdeb_test_normals.argtypes = [u64, i64, ptr, ptr, ptr]
    lib.sdeb_test_philox.argtypes = [C.POINTER(C.c_uint32)]*3
    lib.sdeb_fp64_peak.argtypes = [i64, C.POINTER(f64), ptr]
    lib.sdeb_jit_compile.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(i64),
                                     C.c_char_p, i64]
    lib.sdeb_jit_release.argtypes = [i64]
    if lib.sdeb_abi_version() != ABI_VERSION:
        raise ImportError('libsdeb.so ABI version mismatch: rebuild it')
    return lib


lib = _load()

EXPORTS = ('sdeb_abi_version', 'sdeb_last_error', 'sdeb_device_info',
           'sdeb_plan', 'sdeb_integrate', 'sdeb_moments_workspace',
           'sdeb_moments', 'sdeb_histogram', 'sdeb_mc_range', 'sdeb_mc_update', 'sdeb_antithetic_fold', 'sdeb_path_eval_workspace', 'sdeb_path_cdf',
           'sdeb_path_chf', 'sdeb_path_interp', 'sdeb_axis_reduce', 'sdeb_time_scan', 'sdeb_draw_wiener', 'sdeb_bridge_wiener',
           'sdeb_draw_cpoisson', 'sdeb_test_normals', 'sdeb_test_philox',
           'sdeb_fp64_peak', 'sdeb_jit_compile', 'sdeb_jit_release')


SCAN_CUMSUM, SCAN_INT, SCAN_DIFF = 0, 1, 2
MC_EDGES_GIVEN, MC_EDGES_MINMAX, MC_EDGES_RANGE = 0, 1, 2


def check(rc):
    if rc != 0:
        msg = lib.sdeb_last_error().decode('utf-8', 'replace')
        raise SdebError('libsdeb error %d: %s' % (rc, msg))


def plan(problem):
    p = Plan()
    check(lib.sdeb_plan(C.byref(problem), C.byref(p)))
    return p


def philox(ctr, key):
    out = (C.c_uint32*4)()
    check(lib.sdeb_test_philox((C.c_uint32*4)(*ctr), (C.c_uint32*2)(*key), out))
    return tuple(out)
