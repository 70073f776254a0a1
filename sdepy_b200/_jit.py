"""
Run-time compilation of user-defined SDEs (the ``integrate`` decorator and
direct ``SDE`` / ``SDEs`` subclasses with a Python ``sde`` method, reference
``integration.py:1216-1225, 1711-1730, 1843-1955``).

The Python ``sde(t, x, ..., **params)`` function cannot run inside a kernel,
so it is TRACED: it is called once with symbolic state variables; every NumPy
operation that touches a symbolic value is recorded as a node of an expression
DAG, while sub-expressions that involve parameters only are evaluated by NumPy
itself, exactly as the reference would, and enter the DAG as numeric LEAVES.
The DAG is emitted as a step functor for ``csrc/sde_engine.cuh`` whose
arithmetic uses the never-contracted ``__dmul_rn/__dadd_rn/...`` intrinsics in
the recorded order (so replay mode stays bit-exact with the reference), and the
leaves become the per-lane parameter record (re-traced at every step when they
depend on time).  The source is compiled by NVRTC for sm_100a through
``sdeb_jit_compile``.

``method='milstein'`` adds the (1/2) b b' (dw^2 - dt) correction with b'
obtained by symbolic differentiation of the diffusion node w.r.t. the
equation's own variable (diagonal-noise Milstein).
"""
import ctypes as C
import hashlib
import os

import numpy as np

from . import _engine, _lib
from .infrastructure import lane_values, stack_lane_columns, wiener_source

HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------
# expression DAG
# --------------------------------------------------------------------------

UNARY = {'negative': 'neg', 'positive': 'pos', 'absolute': 'abs', 'fabs': 'abs',
         'sqrt': 'sqrt', 'exp': 'exp', 'log': 'log', 'sin': 'sin', 'cos': 'cos',
         'tanh': 'tanh', 'square': 'square', 'log1p': 'log1p', 'expm1': 'expm1'}
BINARY = {'add': 'add', 'subtract': 'sub', 'multiply': 'mul',
          'true_divide': 'div', 'divide': 'div', 'maximum': 'max',
          'minimum': 'min', 'power': 'pow', 'less': 'lt', 'less_equal': 'le',
          'greater': 'gt', 'greater_equal': 'ge'}


class node:
    """One value of the traced computation: a state variable, a numeric leaf
    or an operation on other nodes.  Behaves like a NumPy scalar in
    arithmetic so that user code runs unmodified."""

    __array_priority__ = 1000.

    def __init__(self, tracer, op, args=(), value=None, index=None):
        self.tracer, self.op, self.args = tracer, op, tuple(args)
        self.value, self.index = value, index
        self.id = len(tracer.nodes)
        tracer.nodes.append(self)

    # ---- NumPy protocol ---------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__' or kwargs.get('out') is not None:
            return NotImplemented
        name = ufunc.__name__
        t = self.tracer
        if name in UNARY and len(inputs) == 1:
            return t.apply(UNARY[name], inputs[0])
        if name in BINARY and len(inputs) == 2:
            return t.apply(BINARY[name], inputs[0], inputs[1])
        raise TypeError(
            'numpy.{} is not supported inside a traced sde function'.format(name))

    def _bin(op):
        def f(self, other):
            return self.tracer.apply(op, self, other)

        def r(self, other):
            return self.tracer.apply(op, other, self)
        return f, r

    __add__, __radd__ = _bin('add')
    __sub__, __rsub__ = _bin('sub')
    __mul__, __rmul__ = _bin('mul')
    __truediv__, __rtruediv__ = _bin('div')
    __pow__, __rpow__ = _bin('pow')
    # comparisons give 0/1-valued nodes (counters of info_next, (y < 0)*1. ...)
    __lt__, __gt__ = _bin('lt')
    __le__, __ge__ = _bin('le')
    del _bin

    def __neg__(self):
        return self.tracer.apply('neg', self)

    def __pos__(self):
        return self

    def __abs__(self):
        return self.tracer.apply('abs', self)

    def __bool__(self):
        raise TypeError('the truth value of a traced state variable is not '
                        'defined: data-dependent Python branches cannot be '
                        'compiled (use numpy.maximum / minimum)')

    def __getitem__(self, key):
        raise TypeError('indexing the state inside a traced sde function is '
                        'not supported: the function is applied per component')

    shape = ()
    ndim = 0
    dtype = np.dtype(float)


class tracer:
    def __init__(self):
        self.nodes, self.leaves = [], []

    def var(self, index):
        return node(self, 'var', index=index)

    def leaf(self, value):
        v = np.asarray(value, dtype=float)
        n = node(self, 'leaf', value=v, index=len(self.leaves))
        # python scalars are structural constants, arrays are parameters
        n.literal = isinstance(value, (int, float)) and not isinstance(value, (bool, np.generic))
        self.leaves.append(n)
        return n

    def lift(self, z):
        if isinstance(z, _sym_state):
            z = z._only()
        return z if isinstance(z, node) else self.leaf(z)

    def apply(self, op, *args):
        if op == 'pos':
            return self.lift(args[0])
        if op == 'pow':
            base, ex = args
            if not isinstance(ex, node):
                exv = np.asarray(ex, dtype=float)
                if exv.shape == () and float(exv) == 2.0:
                    # numpy evaluates x**2 as x*x (fast scalar power)
                    return self.apply('square', base)
                if exv.shape == () and float(exv) == 1.0:
                    return self.lift(base)
                if exv.shape == () and float(exv) == 0.5:
                    return self.apply('sqrt', base)
        return node(self, op, [self.lift(a) for a in args])

    def signature(self, roots):
        """Structural fingerprint (ops and wiring, not leaf values)."""
        parts = []
        for n in self.nodes:
            parts.append('%s:%s:%s' % (n.op, ','.join(str(a.id) for a in n.args),
                                       n.index if n.op in ('var', 'leaf') else ''))
        parts.append('|' + ';'.join(','.join('%s=%d' % (k, v.id) for k, v in r)
                                    for r in roots))
        return '\n'.join(parts)


# symbolic derivative d(node)/d(var index)
def diff(t, n, k, memo):
    key = (n.id, k)
    if key in memo:
        return memo[key]
    op, a = n.op, n.args
    zero, one = 0.0, 1.0

    def mul(u, v):
        if isinstance(u, float) and u == 0.0 or isinstance(v, float) and v == 0.0:
            return 0.0
        if isinstance(u, float) and u == 1.0:
            return v
        if isinstance(v, float) and v == 1.0:
            return u
        return t.apply('mul', u, v)

    def add(u, v):
        if isinstance(u, float) and u == 0.0:
            return v
        if isinstance(v, float) and v == 0.0:
            return u
        return t.apply('add', u, v)

    if op == 'var':
        r = one if n.index == k else zero
    elif op == 'leaf':
        r = zero
    else:
        d = [diff(t, x, k, memo) for x in a]
        if op == 'add':
            r = add(d[0], d[1])
        elif op == 'sub':
            r = add(d[0], mul(-1.0, d[1]))
        elif op == 'mul':
            r = add(mul(d[0], a[1]), mul(a[0], d[1]))
        elif op == 'div':
            r = add(mul(d[0], t.apply('div', 1.0, a[1])),
                    mul(mul(-1.0, d[1]), t.apply('div', n, a[1])))
        elif op == 'neg':
            r = mul(-1.0, d[0])
        elif op == 'square':
            r = mul(mul(2.0, a[0]), d[0])
        elif op == 'sqrt':
            r = mul(d[0], t.apply('div', 0.5, n))
        elif op == 'exp':
            r = mul(d[0], n)
        elif op == 'log':
            r = mul(d[0], t.apply('div', 1.0, a[0]))
        elif op == 'sin':
            r = mul(d[0], t.apply('cos', a[0]))
        elif op == 'cos':
            r = mul(d[0], t.apply('neg', t.apply('sin', a[0])))
        elif op == 'tanh':
            r = mul(d[0], t.apply('sub', 1.0, t.apply('square', n)))
        elif op == 'pow':
            if not (isinstance(d[1], float) and d[1] == 0.0):
                raise NotImplementedError('derivative of x**y with variable y')
            r = mul(d[0], t.apply('mul', a[1], t.apply('pow', a[0], t.apply('sub', a[1], 1.0))))
        elif op in ('abs', 'max', 'min'):
            if all(isinstance(x, float) and x == 0.0 for x in d):
                r = zero
            elif op == 'abs':
                r = mul(d[0], t.apply('sign', a[0]))
            else:
                # d max(u, v) = [u >= v] du + [u < v] dv
                sel = t.apply('ge' if op == 'max' else 'le', a[0], a[1])
                r = add(mul(d[0], sel), mul(d[1], t.apply('sub', 1.0, sel)))
        else:
            raise NotImplementedError('derivative of ' + op)
    memo[key] = r
    return r


# --------------------------------------------------------------------------
# code generation
# --------------------------------------------------------------------------

C_OPS = {
    'add': 'xadd({0}, {1})', 'sub': 'xsub({0}, {1})', 'mul': 'xmul({0}, {1})',
    'div': '__ddiv_rn({0}, {1})', 'neg': '(-{0})', 'abs': 'fabs({0})',
    'sqrt': 'xsqrt_any({0})', 'square': 'xmul({0}, {0})', 'exp': 'exp({0})',
    'log': 'log({0})', 'sin': 'sin({0})', 'cos': 'cos({0})', 'tanh': 'tanh({0})',
    'log1p': 'log1p({0})', 'expm1': 'expm1({0})', 'pow': 'pow({0}, {1})',
    'max': 'xmax({0}, {1})', 'min': 'xmin({0}, {1})',
    'sign': '(({0} > 0.0) - ({0} < 0.0))', 'ge': '(double)({0} >= {1})',
    'le': '(double)({0} <= {1})', 'lt': '(double)({0} < {1})',
    'gt': '(double)({0} > {1})',
}

PRELUDE = r'''
namespace sdeb {
// numpy semantics: np.maximum/minimum propagate the first operand when it is NaN
__device__ __forceinline__ double xmax(double a, double b) { return (a >= b || a != a) ? a : b; }
__device__ __forceinline__ double xmin(double a, double b) { return (a <= b || a != a) ? a : b; }
__device__ __forceinline__ double xsqrt_any(double a) { return sqrt(a); }   // IEEE, like np.sqrt
}
'''


def hexfloat(v):
    v = float(v)
    if v != v:
        return '__longlong_as_double(0x7FF8000000000000LL)'
    if v in (float('inf'), float('-inf')):
        return ('-' if v < 0 else '') + '__longlong_as_double(0x7FF0000000000000LL)'
    return float(v).hex() if v != int(v) or abs(v) > 1e15 else repr(float(v))


class emitter:
    """Emit C statements for the nodes reachable from a set of roots."""

    def __init__(self, slot_of, var_expr):
        self.slot_of, self.var_expr = slot_of, var_expr
        self.lines, self.name = [], {}

    def ref(self, n):
        if n.id in self.name:
            return self.name[n.id]
        if n.op == 'var':
            r = self.var_expr(n.index)
        elif n.op == 'leaf':
            r = hexfloat(n.value) if getattr(n, 'literal', False) else self.slot_of(n.index)
        else:
            args = [self.ref(a) for a in n.args]
            r = 't%d' % n.id
            self.lines.append('double %s = %s;' % (r, C_OPS[n.op].format(*args)))
        self.name[n.id] = r
        return r


# --------------------------------------------------------------------------
# symbolic stand-ins for the hooks that see the whole working array
# (SDE.let, integration.py:1501-1528; info_next through itervars, 650-659)
# --------------------------------------------------------------------------

def _axis_index(key, ndim):
    """The index applied to axis -2 (the last working axis) by a key that is
    trivial on all other axes: ``[..., k, :]`` / ``[..., a:b, :]``; None for a
    key that selects everything."""
    if not isinstance(key, tuple):
        key = (key,)
    if all(k is Ellipsis or k == slice(None) for k in key):
        return None
    tail = []
    for k in reversed(key):
        if k is Ellipsis:
            break
        tail.append(k)
    if len(tail) == 2 and tail[0] == slice(None) and (Ellipsis in key or len(key) == ndim):
        return tail[1]
    raise TypeError('only [...], [..., k, :] and [..., a:b, :] index the working '
                    'state inside a traced let / info_next hook')


class _sym_state:
    """The working array ``wshape + (paths,)`` seen by a traced hook: behaves
    as variable 0 in arithmetic (single equation) and hands out one node per
    equation when indexed the way ``SDEs.unpack`` does (integration.py:
    1735-1757)."""

    __array_priority__ = 1000.

    def __init__(self, owner, variables):
        self._owner, self._vars = owner, variables

    def _var_of(self, key):
        o = self._owner
        idx = _axis_index(key, len(o.wshape) + 1)
        if idx is None:
            return None
        q = len(self._vars)
        if isinstance(idx, (int, np.integer)) and o.addaxis:
            return int(idx) % q
        if isinstance(idx, slice) and not o.addaxis and o._system:
            d = o.vshape[-1]
            a = 0 if idx.start is None else idx.start
            b = q*d if idx.stop is None else idx.stop
            if idx.step in (None, 1) and b - a == d and a % d == 0:
                return a//d
        raise TypeError('unsupported index {!r} on the working state of a traced '
                        'hook'.format(key))

    def __getitem__(self, key):
        k = self._var_of(key)
        return self if k is None else self._vars[k]

    def _only(self):
        if len(self._vars) != 1:
            raise TypeError('the working state of a system holds {} variables: '
                            'select one with self.unpack(x)'.format(len(self._vars)))
        return self._vars[0]

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        args = [a._only() if isinstance(a, _sym_state) else a for a in inputs]
        return self._only().__array_ufunc__(ufunc, method, *args, **kw)


def _delegate(name):
    def f(self, *a):
        return getattr(self._only(), name)(*a)
    return f


for _n in ('__add__', '__radd__', '__sub__', '__rsub__', '__mul__', '__rmul__',
           '__truediv__', '__rtruediv__', '__pow__', '__rpow__', '__neg__', '__abs__',
           '__lt__', '__le__', '__gt__', '__ge__'):
    setattr(_sym_state, _n, _delegate(_n))


class _sym_out:
    """Recorder standing for ``out_x`` in ``let(t, out_x, x)``: remembers what
    is assigned to the whole array or to one equation's slice."""

    def __init__(self, state, var=None, sink=None):
        self._state, self._var = state, var
        self.assigned = {} if sink is None else sink

    def __getitem__(self, key):
        k = self._state._var_of(key)
        return self if k is None else _sym_out(self._state, k, self.assigned)

    def __setitem__(self, key, value):
        k = self._state._var_of(key)
        k = self._var if k is None else k
        self.assigned['all' if k is None else k] = value


class _info_acc:
    """``self.info[key]`` inside a traced info_next: ``+=`` of a state
    expression is recorded as a per-path counter."""

    def __init__(self, key, sink):
        self._key, self._sink = key, sink

    def __iadd__(self, value):
        self._sink.append((self._key, value))
        return self

    def __getattr__(self, name):
        raise NotImplementedError(
            "only `self.info[key] += f(state)` can be compiled from info_next "
            '(attribute {!r} used)'.format(name))


class _info_recorder(dict):
    def __init__(self):
        super().__init__()
        self.recorded = []

    def __getitem__(self, key):
        return _info_acc(key, self.recorded)

    def __setitem__(self, key, value):
        if not isinstance(value, _info_acc):
            raise NotImplementedError(
                'info_next may only accumulate into entries created by info_begin')


class _sym_itervars(dict):
    def __missing__(self, key):
        raise NotImplementedError(
            "itervars[{!r}] is not available to a compiled info_next: use "
            "'last_x' (state before the step) or 'new_x' (after it)".format(key))


# --------------------------------------------------------------------------
# traced SDE mixins
# --------------------------------------------------------------------------

_compiled = {}


def _compile(source, name):
    key = hashlib.sha1(source.encode()).hexdigest()
    if key in _compiled:
        return _compiled[key]
    handle = C.c_int64()
    log = C.create_string_buffer(1 << 16)
    rc = _lib.lib.sdeb_jit_compile(source.encode(), name.encode(), C.byref(handle),
                                   log, len(log))
    if rc != 0:
        raise _lib.SdebError('NVRTC compilation of the traced SDE failed: %s\n%s'
                             % (_lib.lib.sdeb_last_error().decode('utf-8', 'replace'),
                                log.value.decode('utf-8', 'replace')))
    _compiled[key] = handle.value
    return handle.value


# general entry + lean entry (Philox, one constant-bank record; own register budget)
# The library compiles this source in stages (sdeb_jit_compile): first the
# lean entry alone (-DSDEB_JIT_NO_GENERAL), then, on demand, a general entry
# holding the single sweep variant a run needs (-DSDEB_JIT_NO_LEAN
# -DSDEB_SWEEPS=<bit>): a general kernel with all six variants costs ~6x the
# NVRTC time of one.
JIT_ENTRIES = [
    '#ifndef SDEB_JIT_NO_GENERAL',
    'extern "C" __global__ void __launch_bounds__(SDEB_THREADS, SDEB_MIN_BLOCKS)',
    'sdeb_jit_entry(const sdeb::KArgs a) { sdeb::integrate_body<sdeb::UserModel, false, 1>(a); }',
    '#endif',
    '#ifndef SDEB_JIT_NO_LEAN',
    'extern "C" __global__ void __launch_bounds__(SDEB_THREADS, 1)',
    'sdeb_jit_entry_lean(const sdeb::KArgs a) {',
    '    if (sdeb::UserModel::NPC + (sdeb::UserModel::NDW > 1 ? sdeb::UserModel::NDW *',
    '        (sdeb::UserModel::NDW + 1) / 2 : 0) <= sdeb::MAX_CBANK_PARAMS)',
    '        sdeb::integrate_body<sdeb::UserModel, true,\n            sdeb::UserModel::JUMPS ? SDEB_LEAN_JUMP_PPT : SDEB_LEAN_PPT>(a);',
    '}',
    '#endif',
    '#ifdef SDEB_JIT_STREAM',
    'extern "C" __global__ void __launch_bounds__(SDEB_THREADS, 2)',
    'sdeb_jit_entry_stream(const sdeb::KArgs a) {',
    '    sdeb::stream_body<sdeb::UserModel, (SDEB_JIT_STREAM & 2) ? sdeb::NOISE_REPLAY',
    '        : sdeb::NOISE_PHILOX, (SDEB_JIT_STREAM & 1) != 0>(a);',
    '}',
    '#endif',
    '']

PRESET_FUNCTORS = {
    _lib.MODEL_LINEAR: 'LinearSDE<%d, false, false>',
    _lib.MODEL_LINEAR_LOG: 'LinearSDE<%d, true, false>',
    _lib.MODEL_JUMPDIFF: 'LinearSDE<%d, true, true>',
    _lib.MODEL_MEANREV: 'MeanRevertingSDE<%d, false>',
    _lib.MODEL_HULL_WHITE: 'MeanRevertingSDE<%d, true>',
    _lib.MODEL_CIR: 'CoxIngersollRossSDE<%d>',
    _lib.MODEL_HESTON: 'HestonSDE<%d, false>',
    _lib.MODEL_HESTON_FULL: 'HestonSDE<%d, true>',
}


def instantiate_preset(model, ncomp, full=False):
    """NVRTC instantiation of a hand-written preset functor for a component
    count that is not pre-compiled into libsdeb.so (e.g. a 7-factor
    Hull-White), or with full-resolution draws (``draws='full'``:
    -DSDEB_DRAW_FULL=1).  Same engine, same functor source -- only the template
    argument / the draw macro differ."""
    functor = PRESET_FUNCTORS[model] % ncomp
    src = '\n'.join([
        'namespace sdeb { typedef %s UserModel; }' % functor,
        'extern "C" __constant__ int sdeb_jit_dims[6] = {',
        '    sdeb::UserModel::NW, sdeb::UserModel::NDW, sdeb::UserModel::NX,',
        '    sdeb::UserModel::NPC, sdeb::UserModel::NCNT, sdeb::UserModel::JUMPS};',
        ''] + JIT_ENTRIES)
    return _compile(engine_source(full) + src, 'preset_%d_%d%s' % (model, ncomp, '_full'*full))


def engine_source(full=False):
    """The engine header; ``full``: 96 random bits per normal pair."""
    with open(os.path.join(HERE, 'csrc', 'sde_engine.cuh')) as f:
        return ('#define SDEB_DRAW_FULL 1\n' if full else '') + f.read()


class _traced:
    """Generic lowering of an SDE whose equation is a Python ``sde`` method:
    base class of ``integration.SDE`` (the preset equations override it with
    their hand-written kernel functors)."""

    _device_schemes = ('euler', 'milstein')
    _system = False          # True for SDEs (q equations)

    @property
    def _nvars(self):
        return self.q if self._system else 1

    def _param_target(self):
        return self.vshape if self._system else self.wshape

    def _trace(self, t):
        """Call the user's sde at time t with symbolic variables.  Returns
        (tracer, roots) with roots = one list of (id, node) per equation, in
        the order the increments are summed (dict order for one equation,
        integration.py:718; sorted ids for systems, where the reference
        iterates a set, integration.py:1725-1729)."""
        tr = tracer()
        xs = [tr.var(i) for i in range(self._nvars)]
        # t enters as a NumPy scalar, never as a Python float: a Python number
        # is a structural literal baked into the generated source, whereas
        # anything derived from the time (t*x, a*(t - x), {'dt': t}) must
        # become a slot of the per-step parameter record
        A = self.sde(np.float64(t), *xs, **self._sde_args_at(t))
        self._check_sde_values(A)
        eqs = A if self._system else (A,)
        ids = sorted(set().union(*[set(a.keys()) for a in eqs]))
        roots = []
        for a in eqs:
            keys = [k for k in ids if k in a] if self._system else list(a.keys())
            roots.append([(k, tr.lift(a[k])) for k in keys])
        return tr, roots

    def _milstein_nodes(self, tr, roots):
        """((b*b')*0.5) per equation, b' = d b / d(own variable); None when the
        diffusion does not depend on it (Milstein == Euler)."""
        out, memo = [], {}
        for i, r in enumerate(roots):
            b = dict(r).get('dw')
            if b is None:
                out.append(None)
                continue
            db = diff(tr, b, i if self._system else 0, memo)
            if isinstance(db, float) and db == 0.0:
                out.append(None)
            else:
                out.append(tr.apply('mul', tr.apply('mul', b, tr.lift(db)), 0.5))
        return out

    def _traced_signature(self, tr, roots, mil):
        return tr.signature(roots) + '|mil:' + ','.join(
            'n' if m is None else str(m.id) for m in mil)

    def _build(self, t0):
        if getattr(self, '_jit', None) is not None:
            return self._jit
        tr, roots = self._trace(t0)
        milstein = self.method == 'milstein'
        mil = self._milstein_nodes(tr, roots) if milstein else [None]*len(roots)
        sig = self._traced_signature(tr, roots, mil)
        nleaf = len(tr.leaves)
        let = self._trace_let(t0)
        counters = self._trace_info()
        src = self._codegen(tr, roots, mil, let, counters)
        handle = _compile(engine_source(getattr(self, 'draws', 'fast') == 'full') + PRELUDE + src,
                          type(self).__name__)
        self._jit = dict(handle=handle, sig=sig, nleaf=nleaf, milstein=milstein,
                         source=src, let=None if let is None else let[1],
                         counters=[k for k, _, _ in counters])
        return self._jit

    # ---- hooks that are pure functions of the state: traced like `sde` ------
    def _hook_overridden(self, name):
        from .integration import SDE, SDEs
        return getattr(type(self), name) not in (getattr(SDE, name), getattr(SDEs, name))

    def _trace_let(self, t):
        """A user ``let(t, out_x, x)`` (reference integration.py:1501-1528) run
        once with a symbolic working state: what it assigns to ``out_x`` -- the
        whole array or one equation's slice at a time -- becomes the kernel's
        emit().  Returns None for the default ``out_x[...] = x``, else
        (tracer, kind, nodes): kind 'single' (one value per element, xshape =
        vshape) or 'vars' (one per equation, xshape = wshape)."""
        if not self._hook_overridden('let'):
            return None
        tr = tracer()
        xs = [tr.var(i) for i in range(self._nvars)]
        state = _sym_state(self, xs)
        out = _sym_out(state)
        self.let(np.float64(t), out, state)
        got = out.assigned
        if not got:
            raise NotImplementedError('let() did not assign to out_x')
        if 'all' in got:
            v = got['all']
            if isinstance(v, _sym_state):
                return None                           # out_x[...] = x
            return tr, 'single', [tr.lift(v)]
        if set(got) != set(range(self._nvars)):
            raise NotImplementedError(
                'let() must fill every equation\'s slice of out_x (got {})'.format(sorted(got)))
        return tr, 'vars', [tr.lift(got[k]._only() if isinstance(got[k], _sym_state) else got[k])
                            for k in range(self._nvars)]

    def _trace_info(self):
        """A user ``info_next`` (integration.py:1562-1567) of the form
        ``self.info[key] += f(itervars['last_x'] | itervars['new_x'])``: each
        accumulation becomes a per-path counter of the kernel.  Returns a list
        of (key, tracer, node-with-variables) with variables 0..q-1 = last_x,
        q..2q-1 = new_x; [] when the hook is not overridden."""
        if not self.getinfo or not self._hook_overridden('info_next'):
            return []
        tr = tracer()
        q = self._nvars
        last = _sym_state(self, [tr.var(i) for i in range(q)])
        new = _sym_state(self, [tr.var(q + i) for i in range(q)])
        real_info, real_iv = self.info, getattr(self, 'itervars', None)
        rec = _info_recorder()
        self.info, self.itervars = rec, _sym_itervars(last_x=last, new_x=new)
        try:
            self.info_next()
        finally:
            self.info = real_info
            if real_iv is None:
                del self.itervars
            else:
                self.itervars = real_iv
        out = []
        for key, v in rec.recorded:
            if isinstance(v, _sym_state):
                v = v._only()
            if not isinstance(v, node):
                raise NotImplementedError(
                    'info_next accumulates {!r}, which does not depend on the state: '
                    'do it in info_end'.format(v))
            out.append((key, tr, v))
        return out

    def _leaf_values(self, t):
        """Re-trace at time t and return the leaf values (structure checked)."""
        tr, roots = self._trace(t)
        mil = (self._milstein_nodes(tr, roots) if self._jit['milstein']
               else [None]*len(roots))
        if self._traced_signature(tr, roots, mil) != self._jit['sig']:
            raise NotImplementedError(
                'the sde function takes a different computation path at t={}: '
                'time-dependent control flow cannot be compiled'.format(t))
        return [n.value for n in tr.leaves]

    # ---- SDE lowering hooks -----------------------------------------------
    def _spec(self):
        lead, ncomp = self._lanes()
        groups = int(np.prod(lead, dtype=int))
        jit = self._build(0.)
        return _engine.problem_spec(_lib.MODEL_JIT, ncomp, groups,
                                    jit_handle=jit['handle']), lead

    def _stats_centre(self, w0l):
        if (getattr(self, '_jit', None) or {}).get('let'):
            # a user let(): the stored values are not the working state; the
            # centre is only a shift of the power sums
            lead, nw = self._lanes()
            nx = nw//self._nvars if self._jit['let'] == 'single' else nw
            return np.zeros((w0l.shape[0], nx))
        w = w0l.mean(axis=-1)
        return np.exp(w) if self.log else w

    def _records(self, spec, seg, lead, replay):
        jit = self._jit
        dw = self.sources.get('dw')
        lanes = self._param_target()
        # is anything time-dependent?  (callable parameters, explicit use of t)
        probe = [float(seg.s[0]), float(seg.s[-1])] if seg.n_steps else [0.]
        vals = [self._leaf_values(s) for s in probe]
        tdep = any(not np.array_equal(a, b) for a, b in zip(vals[0], vals[-1]))
        tdep = tdep or any(callable(z) for z in
                           self._get_args(self._sde_args_keys).values())
        corr_t = not replay and isinstance(dw, wiener_source) and callable(dw.corr)
        jumps_t = spec.jumps and not replay and self._jumps_time_dependent()
        n = seg.n_steps if (tdep or corr_t or jumps_t) and seg.n_steps else 1
        nw = spec.nw
        elems = nw//self._nvars          # elements of the last axis owned by a lane
        blocks = []
        for i in range(n):
            s = seg.s[i] if seg.n_steps else 0.
            ds = seg.ds[i] if seg.n_steps else 0.
            leaves = vals[0] if (i == 0 or not tdep) else self._leaf_values(float(s))
            # record = [elements x leaves] then, with jumps, [components x 6]
            cols = [lane_values(v, lanes, 'SDE parameter', paths=self.paths) for v in leaves]
            block = (stack_lane_columns(cols, spec.groups, elems) if cols
                     else np.zeros((spec.groups, 0)))
            for id, src in self._jump_slots():
                block = _join_blocks(block, self._jump_block(spec, id, src, s, ds, replay))
            L = None
            if spec.nchol and not replay and isinstance(dw, wiener_source):
                L = dw.chol_at(s + ds/2)
            blocks.append((block, _engine.chol_entries(L, spec.ndw)))
        return _engine.assemble_records(blocks, spec)

    def _jump_block(self, spec, id, dj, s, ds, replay):
        """[groups, 6*nw (, paths)] of ONE jump slot: per working component lam,
        (reserved), law, a, b, pa -- intensity and law sampled at the step
        midpoint (infrastructure.py:1630, 2031)."""
        nw = spec.nw
        if replay:
            return np.zeros((spec.groups, 6*nw))
        mid = s + ds/2
        if id == 'dn':                   # plain Poisson: unit jump sizes
            lam_src = dj
            kind, a, b, pa = _lib.LAW_UNIFORM, 1., 1., 0.
        else:
            lam_src = dj.dn
            kind, a, b, pa = dj.y.at(mid)
        cols = []
        for z in (lam_src.lam_at(mid), 0., float(kind), a, b, pa):
            z = np.asarray(z, dtype=float)
            if z.ndim == len(self.wshape) + 1 and z.shape[-1] == self.paths:
                full = np.broadcast_to(z, self.wshape + (self.paths,))
            else:
                if z.ndim == len(self.wshape) + 1 and z.shape[-1] == 1:
                    z = z[..., 0]
                full = np.broadcast_to(z, self.wshape)[..., np.newaxis]
            cols.append(self._to_lanes(full).reshape((spec.groups, nw, -1)))
        width = max(c.shape[-1] for c in cols)
        cols = [np.broadcast_to(c, (spec.groups, nw, width)) for c in cols]
        block = np.stack(cols, axis=2).reshape((spec.groups, 6*nw, width))
        return block[..., 0] if width == 1 else block

    # ---- code generation ----------------------------------------------------
    def _codegen(self, tr, roots, mil, let=None, counters=()):
        """UserModel functor.  A lane owns ``elems`` elements of the last axis,
        each with ``q`` variables (q = 1 for a single equation; elems = 1
        unless the Wiener components are coupled by a correlation matrix);
        variable k of element h is working component k*elems + h -- the
        reference's own stacking (integration.py:1735-1757)."""
        lead, nw = self._lanes()
        q = self._nvars
        elems = nw//q
        ids = set(k for r in roots for k, _ in r)
        unknown = ids - {'dt', 'dw', 'dj', 'dn'}
        if unknown:
            raise NotImplementedError('differentials {} have no device '
                                      'implementation'.format(unknown))
        slot_of = {id: k for k, (id, _) in enumerate(self._jump_slots())}
        jumps = len(slot_of)             # jump slots: 'dj' and / or 'dn'
        nleaf = len(tr.leaves)

        def comp(k):
            return '%d*E + h' % k if elems > 1 else '%d' % k
        em = emitter(lambda j: 'p[%d*h + %d]' % (nleaf, j),
                     lambda k: 'x[%s]' % comp(k))
        incs = []
        for k, r in enumerate(roots):
            terms = []
            for ident, nd in r:
                dz = ('ds' if ident == 'dt' else 'dw[%s]' % comp(k) if ident == 'dw' else
                      'dj[%d + %s]' % (slot_of[ident]*nw, comp(k)))
                terms.append('xmul(%s, %s)' % (em.ref(nd), dz))
            inc = terms[0] if terms else '0.0'
            for tm in terms[1:]:
                inc = 'xadd(%s, %s)' % (inc, tm)
            incs.append(inc)
        body = list(em.lines)
        em.lines = []
        for k, inc in enumerate(incs):
            body.append('double xn%d = xadd(x[%s], %s);' % (k, comp(k), inc))
        for k in range(q):
            if mil[k] is not None:
                body += _flush(em, mil[k])
                body.append('xn{0} = xadd(xn{0}, xmul({1}, xsub(xmul(dw[{2}], dw[{2}]), ds)));'
                            .format(k, em.ref(mil[k]), comp(k)))
        # info_next counters (integer-valued expressions of the state before /
        # after the step), one per accumulation and element
        base = nw*jumps
        pre, post = [], []
        for j, (key, ctr, nd) in enumerate(counters):
            uses_new = _uses_vars(nd, range(q, 2*q))
            cem = emitter(_literal_slot(ctr),
                          lambda k: ('x[%s]' % comp(k)) if k < q else 'xn%d' % (k - q))
            lines = _flush(cem, nd)
            lines.append('cnt[%d + %d*E + h] += (int)(%s);' % (base, j, cem.ref(nd)))
            (post if uses_new else pre).append('{ ' + ' '.join(lines) + ' }')
        body = pre + body + post
        for k in range(q):
            body.append('x[%s] = xn%d;' % (comp(k), k))
        npc = elems*nleaf + 6*nw*jumps
        ncnt = base + len(counters)*elems
        # user let(): emit() body
        emit, nx = None, nw
        if let is not None:
            ltr, kind, nodes = let
            lem = emitter(_literal_slot(ltr), lambda k: 'x[%s]' % comp(k))
            emit = []
            for k, nd in enumerate(nodes):
                emit += _flush(lem, nd)
                val = lem.ref(nd)
                emit.append('v[%s] = %s;' % ('h' if kind == 'single' else comp(k),
                                             'exp(%s)' % val if self.log else val))
            nx = elems if kind == 'single' else nw
        return _model_source(nw, nw, npc, jumps, 6, elems*nleaf, body, self.log, elems,
                             ncnt=ncnt, nx=nx, emit=emit)

    def _device_info(self, res, tt, segs, replay):
        """Counters compiled from info_next: added to the entries info_begin
        created (reference integration.py:1166-1173)."""
        keys = (getattr(self, '_jit', None) or {}).get('counters') or []
        if not keys or res.counter is None:
            return
        lead, nw = self._lanes()
        elems = nw//self._nvars
        base = nw*len(self._jump_slots())
        c = res.counter.reshape((-1, base + len(keys)*elems, self.paths))
        for j, key in enumerate(keys):
            part = c[:, base + j*elems:base + (j + 1)*elems, :]
            part = part.cpu().numpy().reshape(tuple(lead) + ((elems,) if elems > 1 or
                                              len(lead) < len(self.vshape) else ()) + (self.paths,))
            part = part.reshape(self.vshape + (self.paths,))
            self.info[key] = self.info[key] + part.astype(np.asarray(self.info[key]).dtype)


def _uses_vars(nd, which, seen=None):
    seen = set() if seen is None else seen
    if nd.id in seen:
        return False
    seen.add(nd.id)
    if nd.op == 'var':
        return nd.index in which
    return any(_uses_vars(a, which, seen) for a in nd.args)


def _literal_slot(tr):
    """Leaves of a traced let / info_next are baked into the source: they must
    be scalars that do not depend on time."""
    def slot(j):
        v = np.asarray(tr.leaves[j].value, dtype=float)
        if v.size != 1:
            raise NotImplementedError(
                'array-valued parameters inside let / info_next cannot be compiled')
        return hexfloat(v.reshape(()))
    return slot


def _flush(em, nd):
    """Emit the statements needed for node nd (and return them)."""
    before = len(em.lines)
    em.ref(nd)
    new = em.lines[before:]
    del em.lines[before:]
    return new


def _join_blocks(a, b):
    """Concatenate two record blocks [groups, n (, paths)] along axis 1."""
    if a.ndim != b.ndim:
        paths = (a if a.ndim == 3 else b).shape[-1]
        a, b = (np.broadcast_to(z[..., np.newaxis], z.shape + (paths,)) if z.ndim == 2 else z
                for z in (a, b))
    return np.concatenate((a, b), axis=1)


def _model_source(nw, ndw, npc, jumps, jp_stride, jp_off, body, log, elems, ncnt=None,
                  nx=None, emit=None):
    ncnt = nw*int(jumps) if ncnt is None else ncnt
    nx = nw if nx is None else nx
    lines = ['namespace sdeb {', 'struct UserModel {',
             '    enum { NW = %d, NDW = %d, NX = %d, NPC = %d, NCNT = %d, JUMPS = %d,'
             % (nw, ndw, nx, npc, ncnt, int(jumps)),
             '           JP_STRIDE = %d, JP_OFF = %d };' % (jp_stride, jp_off),
             '    static __device__ __forceinline__ void step(double (&x)[NW], const double* p,',
             '            double ds, const double* dw, const double* dj, int (&cnt)[NCNT + 1]) {']
    lines.append('        enum { E = %d };' % elems)
    lines.append('#pragma unroll')
    lines.append('        for (int h = 0; h < E; ++h) {')
    lines += ['            ' + b for b in body]
    lines.append('        }')
    lines += ['    }',
              '    static __device__ __forceinline__ void emit(const double (&x)[NW], double (&v)[NX]) {']
    if emit is None:
        lines += ['#pragma unroll',
                  '        for (int c = 0; c < NW; ++c) v[c] = %s;' % ('exp(x[c])' if log else 'x[c]')]
    else:
        lines += ['        enum { E = %d };' % elems, '#pragma unroll',
                  '        for (int h = 0; h < E; ++h) {'] + ['            ' + b for b in emit] + [
                  '        }']
    lines += ['    }', '};', '}',
              'extern "C" __constant__ int sdeb_jit_dims[6] = {%d, %d, %d, %d, %d, %d};'
              % (nw, ndw, nx, npc, ncnt, int(jumps)),
              ''] + JIT_ENTRIES
    return '\n'.join(lines)


def evaluate(nd, xs, cache=None):
    """Host interpreter of a DAG node (NumPy arithmetic, same rounding as the
    emitted CUDA for + - * / sqrt max min): used by the CPU tests to check the
    tracer and the symbolic derivative without a GPU."""
    cache = {} if cache is None else cache
    if nd.id in cache:
        return cache[nd.id]
    if nd.op == 'var':
        r = xs[nd.index]
    elif nd.op == 'leaf':
        r = nd.value
    else:
        a = [evaluate(z, xs, cache) for z in nd.args]
        f = {'add': np.add, 'sub': np.subtract, 'mul': np.multiply,
             'div': np.divide, 'neg': np.negative, 'abs': np.abs,
             'sqrt': np.sqrt, 'square': np.square, 'exp': np.exp, 'log': np.log,
             'sin': np.sin, 'cos': np.cos, 'tanh': np.tanh, 'log1p': np.log1p,
             'expm1': np.expm1, 'pow': np.power, 'max': np.maximum,
             'min': np.minimum, 'sign': np.sign,
             'ge': lambda u, v: (u >= v)*1., 'le': lambda u, v: (u <= v)*1.,
             'lt': lambda u, v: (u < v)*1., 'gt': lambda u, v: (u > v)*1.}[nd.op]
        r = f(*a)
    cache[nd.id] = r
    return r
