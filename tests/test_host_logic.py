"""CPU tests (no GPU): the C-ABI library loads and exports every symbol of
include/sdeb.h, the ctypes structs match the header, host-side lowering
(step grid, sweeps, parameter records, error conventions) is right, and
compute entry points fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import sdepy_b200 as sd
from sdepy_b200 import _engine, _lib
from oracle import sde_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
no_gpu = not torch.cuda.is_available()


def test_library_exports_header_symbols():
    header = open(os.path.join(ROOT, 'include', 'sdeb.h')).read()
    declared = set(re.findall(r'\b(sdeb_[a-z0-9_]+)\s*\(', header))
    declared -= {'sdeb_plan_t'}
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.sdeb_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    header = open(os.path.join(ROOT, 'include', 'sdeb.h')).read()
    body = header[header.index('typedef struct sdeb_problem {'):header.index('} sdeb_problem;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = re.findall(r'(\w+);', body)
    assert names == [f[0] for f in _lib.Problem._fields_]
    assert ctypes.sizeof(_lib.Problem) == 8*len(names)
    body = header[header.index('typedef struct sdeb_plan_t {'):header.index('} sdeb_plan_t;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    assert re.findall(r'(\w+);', body) == [f[0] for f in _lib.Plan._fields_]


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    assert _lib.philox((0, 0, 0, 0), (0, 0)) == (
        0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert _lib.philox((0xffffffff,)*4, (0xffffffff,)*2) == (
        0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert _lib.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
                       (0xa4093822, 0x299f31d0)) == (
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_plan_shapes():
    for model, n, want in [
            (_lib.MODEL_LINEAR, 1, (1, 1, 1, 2, 2)),
            (_lib.MODEL_LINEAR_LOG, 3, (3, 3, 3, 6, 12)),
            (_lib.MODEL_JUMPDIFF, 1, (1, 1, 1, 8, 8)),
            (_lib.MODEL_HULL_WHITE, 3, (3, 3, 1, 9, 15)),
            (_lib.MODEL_HESTON, 1, (2, 2, 1, 6, 9)),
            (_lib.MODEL_HESTON_FULL, 2, (4, 4, 4, 12, 22))]:
        s = _engine.problem_spec(model, n, 1)
        assert (s.nw, s.ndw, s.nx, s.npc, s.npt) == want
    with pytest.raises(_lib.SdebError):
        _engine.problem_spec(_lib.MODEL_HESTON, 7, 1)


def _plan(**kw):
    p = _lib.Problem()
    p.abi_version = _lib.ABI_VERSION
    for k, v in dict(ncomp=1, noise=0, n_paths=10**6, pitch=10**6, n_steps=500, n_groups=1,
                     n_rows=2, row0=0, n_psteps=1, params=0x1000, w0=0x1000, steps=0x1000,
                     store_row=0x1000).items():
        setattr(p, k, v)
    for k, v in kw.items():
        setattr(p, k, v)
    plan = _lib.Plan()
    _lib.check(_lib.lib.sdeb_plan(ctypes.byref(p), ctypes.byref(plan)))
    return plan


def test_plan_kernel_choice_and_shared_memory_budget():
    """Host-only sdeb_plan: which kernel a problem gets (0 general, 1 lean, 2 stream)
    and that the dynamic shared memory -- 32 KB alignment gap + 72 KB generator
    tables, staged records inside the gap -- leaves room for two blocks per SM
    (233472 B per SM, 1 KB reserved + < 4 KB static per block) in the hot shapes."""
    tables, gap, per_sm = 8*(2*256 + 2*256 + 2*64)*8, 32768, 233472
    rec = np.zeros(16)
    # north star: Heston, Philox, one time-invariant record, terminal statistics
    lean = _plan(model=_lib.MODEL_HESTON, params_host=rec.ctypes.data, stats=0x1000,
                 centre=0x1000, workspace=0x1000)
    assert lean.kernel == 1 and lean.stats_in_kernel == 1
    assert lean.smem_bytes == gap + tables + 2*1*8*8       # + the accumulators of 2 rows
    # C2b: three correlated factors, time-dependent records, full paths stored
    c2b = _plan(model=_lib.MODEL_HULL_WHITE, ncomp=3, n_psteps=500, n_rows=501, out=0x1000)
    assert c2b.kernel == 2 and c2b.npt == 15
    assert c2b.smem_bytes == gap + tables                  # 64 x 16 doubles live in the gap
    # replayed increments: no generator tables, records + the cp.async ring
    rep = _plan(model=_lib.MODEL_MEANREV, noise=1, dW=0x1000, n_rows=501, out=0x1000)
    assert rep.kernel == 2 and rep.smem_bytes == 8*1*256*2*8
    for plan in (lean, c2b):
        assert 2*(plan.smem_bytes + 1024 + 4096) <= per_sm
    # odd pitch: rows are not 16-byte aligned -> the general kernel
    odd = _plan(model=_lib.MODEL_MEANREV, n_paths=999, pitch=999, n_rows=501, out=0x1000)
    assert odd.kernel == 0


def test_step_grid_and_sweeps_follow_reference():
    P = sd.ornstein_uhlenbeck_process(paths=3, steps=(0.1, .2, .21, .5, .93, 1.5))
    tt = np.array((0., .37, 1.))
    target = P.pace(tt)
    target = target[(target >= 0.) & (target <= 1.)]
    grid = np.unique(np.concatenate((target, tt)))
    assert np.array_equal(grid, orc.step_grid(tt, (0.1, .2, .21, .5, .93, 1.5))[1])
    seg, = _engine.segments_of(tt, grid, 0)
    assert seg.row0 == 0 and list(seg.store_row) == [-1, -1, -1, 1, -1, -1, 2]
    assert np.array_equal(seg.ds, np.diff(grid))
    back, fwd = _engine.segments_of(tt, grid, 1)
    assert np.array_equal(back.s, grid[4:0:-1]) and (back.ds < 0).all()
    assert list(back.store_row) == [-1, -1, -1, 0] and back.row0 == 1
    assert list(fwd.store_row) == [-1, -1, 2]
    last, = _engine.segments_of(tt, grid, -1)
    assert (last.ds < 0).all() and last.row0 == 2


def test_heston_records_and_initial_state():
    P = sd.heston_process(paths=10, steps=20, x0=100., y0=.04, mu=.03, sigma=.5,
                          theta=.04, k=2., xi=.3, rho=-.7, seed=3)
    assert (P.vshape, P.xshape, P.wshape) == ((), (), (2,))
    spec, lead = P._spec()
    seg, = _engine.segments_of(np.array([0., 1.]), np.linspace(0, 1, 5), 0)
    rec = P._records(spec, seg, lead, False)
    assert rec.shape == (1, 1, 9)
    L = np.linalg.cholesky(np.array([[1, -.7], [-.7, 1]]))
    assert np.allclose(rec[0, 0], [.03, .5*.5/2, .5, .04, 2., .3, 1., L[1, 0], L[1, 1]])
    w0 = P._initial_state(0.)
    assert np.array_equal(w0[:, 0], [np.log(100.), .04])
    # time-dependent parameter => one record per step, evaluated at the left point
    Q = sd.ornstein_uhlenbeck_process(paths=4, theta=lambda t: .2 + .1*t)
    spec, lead = Q._spec()
    rec = Q._records(spec, seg, lead, False)
    assert rec.shape == (4, 1, 3) and np.allclose(rec[:, 0, 0], .2 + .1*seg.s)


def test_vshape_lanes():
    P = sd.lognorm_process(paths=5, vshape=(2, 3), x0=1.)
    assert P._lanes() == ((2, 3), 1)
    C = sd.lognorm_process(paths=5, vshape=(2, 3), corr=np.eye(3))
    assert C._lanes() == ((2,), 3)
    H = sd.hull_white_process(paths=5, vshape=(4,), factors=2)
    assert H._lanes() == ((4,), 2) and H.wshape == (4, 2) and H.xshape == (4,)
    F = sd.full_heston_process(paths=5, vshape=(2,))
    assert F.wshape == (4,) and F._lanes() == ((), 2)
    # a parameter with a real paths axis gives path-dependent records
    B = sd.lognorm_process(paths=5, mu=np.arange(5.), sigma=.5)
    spec, lead = B._spec()
    seg, = _engine.segments_of(np.array([0., 1.]), np.array([0., 1.]), 0)
    rec = B._records(spec, seg, lead, False)
    assert rec.shape == (1, 1, 2, 5)
    assert np.array_equal(rec[0, 0, 0], np.arange(5.) - .5*.5/2) and (rec[0, 0, 1] == .5).all()
    with pytest.raises(ValueError):
        C2 = sd.lognorm_process(paths=5, mu=np.arange(4.))
        C2._records(*((C2._spec()[0], seg, C2._spec()[1], False)))


def test_construction_errors():
    with pytest.raises(TypeError):
        sd.lognorm_process(paths=10, nonexistent=1.)
    with pytest.raises(ValueError):
        sd.lognorm_process(paths=10, method='nonexistent')
    with pytest.raises(ValueError):
        sd.wiener_source(paths=3, vshape=(), rho=.5)
    with pytest.raises(ValueError):
        sd.wiener_source(paths=3, vshape=(3,), rho=.5)
    with pytest.raises(TypeError):
        sd.wiener_source(paths=3, rng=1234)
    with pytest.raises(TypeError):
        class bad(sd.SDE):
            pass
        bad()
    s = sd.wiener_source(paths=3, vshape=(4,), rho=(.1, .2))
    assert s.corr.shape == (4, 4) and s.corr[0, 2] == .1 and s.corr[1, 3] == .2


def test_same_seed_same_keys():
    a = sd.wiener_source(paths=3, rng=np.random.default_rng(7))
    b = sd.wiener_source(paths=3, rng=np.random.default_rng(7))
    assert a.seed == b.seed and a.next_key() == b.next_key()
    assert a.next_key() != a.next_key()
    c = sd.wiener_source(paths=3, seed=42)
    assert c.seed == 42


def test_keys_distinct_over_seed_epoch_grid():
    """Adjacent seeds must not share streams on successive calls (the key of
    call n of seed s used to equal the key of call 0 of seed s + n)."""
    keys = set()
    for seed in range(64):
        src = sd.wiener_source(paths=2, seed=seed)
        for _ in range(64):
            keys.add(src.next_key())
    assert len(keys) == 64*64
    a, b = sd.wiener_source(paths=2, seed=0), sd.wiener_source(paths=2, seed=1)
    a.next_key()
    assert a.next_key() != b.next_key()
    P0, P1 = (sd.lognorm_process(paths=2, seed=s) for s in (0, 1))
    P0._philox_key()
    assert P0._philox_key() != P1._philox_key()


def test_explicit_time_in_traced_sde_is_a_record_slot():
    """`t` used arithmetically inside a traced function must vary from step to
    step (it used to be frozen at its value at the first trace)."""
    from sdepy_b200 import _jit

    def f(t, x, a=2.):
        return {'dt': a*(t - x), 'dw': 1 + t}

    P = sd.integrate(f)(paths=4, a=2.)
    tr0, r0 = P._trace(0.)
    tr1, r1 = P._trace(.5)
    assert tr0.signature(r0) == tr1.signature(r1)
    assert not any(getattr(l, 'literal', False) and float(l.value) == 0. for l in tr0.leaves)
    x = np.array([.1, .2, .3])
    for t, (tr, roots) in ((0., (tr0, r0)), (.5, (tr1, r1))):
        got = {k: _jit.evaluate(n, [x]) for k, n in roots[0]}
        want = f(t, x)
        assert np.array_equal(got['dt'], want['dt'])
        assert np.array_equal(np.broadcast_to(got['dw'], x.shape), np.broadcast_to(want['dw'], x.shape))
    src = P._codegen(tr0, r0, [None])
    assert 'xsub(0.0' not in src and 'xmul(0.0' not in src


def test_process_container():
    t = np.linspace(0, 1, 5)
    x = np.arange(5*2*3, dtype=float).reshape(5, 2, 3)
    p = sd.process(t, x=x)
    assert p.paths == 3 and p.vshape == (2,)
    assert np.array_equal(np.asarray(p.pmean()), x.mean(axis=-1, keepdims=True))
    assert np.array_equal(np.asarray(p.pvar(ddof=1)), x.var(axis=-1, ddof=1, keepdims=True))
    assert np.allclose(p(.125), (x[0] + x[1])/2)
    assert np.allclose(p(.25, .25), x[2] - x[1])
    with pytest.raises(ValueError):
        sd.process(t, x=x[:4])


@pytest.mark.skipif(not no_gpu, reason='only meaningful without a device')
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError):
        sd.lognorm_process(paths=10, steps=5)((0., 1.))
    with pytest.raises(RuntimeError):
        sd.montecarlo(np.arange(10.))
    out = ctypes.c_double()
    rc = _lib.lib.sdeb_fp64_peak(10, ctypes.byref(out), None)
    assert rc != 0 and _lib.lib.sdeb_last_error()


def test_tracer_and_symbolic_derivative():
    from sdepy_b200 import _jit

    def f(t, x, a=1.5, b=.3):
        return {'dt': a*(b - x)/(1 + x**2), 'dw': np.sqrt(np.maximum(x, 0.))*np.exp(-b*x) + np.sin(x)*x}

    P = sd.integrate(f)(paths=4, a=1.5, b=.3)
    tr, roots = P._trace(0.)
    x = np.array([.2, .7, 1.3, 2.])
    got = {k: _jit.evaluate(n, [x]) for k, n in roots[0]}
    want = f(0., x)
    assert np.array_equal(got['dt'], want['dt']) and np.array_equal(got['dw'], want['dw'])
    memo = {}
    db = _jit.diff(tr, dict(roots[0])['dw'], 0, memo)
    h = 1e-6
    num = (f(0., x + h)['dw'] - f(0., x - h)['dw'])/(2*h)
    assert np.allclose(_jit.evaluate(tr.lift(db), [x]), num, rtol=1e-6)
    # structure is stable in time, leaf values are not
    P2 = sd.integrate(f)(paths=4, a=lambda t: 1.5 + t, b=.3)
    t0, r0 = P2._trace(0.)
    t1, r1 = P2._trace(1.)
    assert t0.signature(r0) == t1.signature(r1)
    assert [float(l.value) for l in t0.leaves] != [float(l.value) for l in t1.leaves]


def test_integrate_decorator_infers_equations_and_sources():
    @sd.integrate
    def one(t, x, k=1.):
        return {'dt': -k*x, 'dw': 1.}
    assert one.q == 0 and one.sources == {'dt', 'dw'}

    # like the reference, a function of several variables cannot be probed by
    # f() / f(1., 1.) unless its arguments have defaults (integration.py:1851-1859)
    @sd.integrate
    def two(t=0., x=1., y=1., k=1.):
        return ({'dt': -k*x, 'dw': y}, {'dt': 0, 'dw': .1})
    assert two.q == 2 and issubclass(two, sd.SDEs)
    with pytest.raises(TypeError):
        sd.integrate(lambda t, x, y: ({'dt': x}, {'dt': y}))
    P = two(paths=3, x0=(1., 2.))
    assert P.wshape == (2,) and P.addaxis
    with pytest.raises(TypeError):
        sd.integrate(lambda: 1/0)
    with pytest.raises(TypeError):
        sd.integrate(one.sde, q=3)


def test_host_process_cdf_chf_interp_match_reference():
    from tests.cases import golden
    g, s = golden('stats_cdf_chf'), golden('stats_lognorm')
    p = sd.process(s['t'], x=s['x'])
    assert np.array_equal(p(g['tq']), g['interp'])
    assert np.array_equal(p(.25, .5), g['incr'])
    assert np.array_equal(p.cdf(g['xq']), g['cdf_tl'])
    assert np.array_equal(p.cdf(g['tq'], g['xq']), g['cdf_t'])
    assert np.allclose(p.chf(g['uq']), g['chf_tl'], rtol=1e-14, atol=1e-16)
    assert np.allclose(p.chf(g['tq'], g['uq']), g['chf_t'], rtol=1e-14, atol=1e-16)
    with pytest.raises(TypeError):
        p.cdf()


def test_stacked_system_lane_layout_round_trip():
    """SDEs with addaxis=False keep variable k of element h at k*d + h along
    the last axis (reference integration.py:1735-1757); the kernel wants the q
    variables of an element adjacent.  _to_lanes / _from_lanes transpose
    between the two and are inverse to each other."""
    import sdepy_b200 as m
    from tests.cases import user_system
    cls = m.integrate(q=2, sources={'dt', 'dw'})(user_system)
    P = cls(paths=5, vshape=(2, 3), steps=4, x0=(1., .3))
    assert P.wshape == (2, 6) and P._lanes() == ((2, 3), 2)
    w = np.arange(2*6*5, dtype=float).reshape(2, 6, 5)
    lanes = P._to_lanes(w)                       # (2, 3, q, paths)
    assert lanes.shape == (2, 3, 2, 5)
    x, y = P.unpack(w)
    assert np.array_equal(lanes[..., 0, :], x) and np.array_equal(lanes[..., 1, :], y)
    rows = np.stack((w, 2*w))
    flat = P._to_lanes(rows, 1).reshape(2, -1, 5)
    assert np.array_equal(P._from_lanes(flat, (5,)), rows)
    # presets address the reference's working layout natively
    H = m.full_heston_process(paths=5, vshape=(2,), steps=4)
    w4 = np.zeros((4, 5))
    assert H._to_lanes(w4) is w4 and H._lanes() == ((), 2)
    # addaxis=True: identity
    Q = m.integrate(q=2, sources={'dt', 'dw'}, addaxis=True)(user_system)(
        paths=5, vshape=(2, 3), steps=4, x0=(1., .3))
    w = np.zeros((2, 3, 2, 5))
    assert Q._to_lanes(w) is w and Q._lanes() == ((2, 3), 2)


def test_kfunc_parameter_management():
    """The documented kfunc protocol (reference kfun.py:262-341): parameters
    stored, variables given at evaluation, re-instantiation on new parameters."""
    import sdepy_b200 as m

    class scaled:
        def __init__(self, *, a=1., b=0.):
            self.a, self.b = a, b

        def __call__(self, t, dt=1.):
            return self.a*t + self.b*dt

    K = m.kfunc(scaled)
    assert m.iskfunc(K) and not m.iskfunc(scaled) and m.iskfunc(m.kfunc(K))
    inst = K(a=2.)
    assert m.iskfunc(inst) and isinstance(inst, scaled)
    assert inst.params == {'a': 2., 'b': 0.}
    assert inst(3.) == 6. and inst(t=3., dt=2.) == 6.
    assert inst(3., b=1.) == 7. and inst.b == 0.          # inst unaffected
    other = inst(b=5.)
    assert other.params == {'a': 2., 'b': 5.} and other is not inst
    assert K(3., 2., a=1., b=1.) == 5.                     # instantiate + evaluate
    assert K(a=1., b=1., dt=2., t=3.) == 5.                # variables by name

    @m.kfunc(nvar=1)
    def line(t, *, a=1., b=2.):
        return a*t + b
    g = line(a=3.)
    assert g(2.) == 8. and g(2., b=0.) == 6. and line(2., a=1.) == 4.
    assert g(b=5.).params == {'a': 3., 'b': 5.}
    with pytest.raises(TypeError):      # unknown parameter: raised at evaluation,
        line(c=1.)(2.)                  # as in the reference (test_kfunc.py:213)

    # the shortcuts are kfuncs over the plain classes, which stay plain
    assert m.iskfunc(m.lognorm) and not m.iskfunc(m.lognorm_process)
    P = m.heston(x0=100., rho=-.7, paths=10)
    Q = P(paths=20, steps=7)
    assert (P.paths, Q.paths, Q.params['steps']) == (10, 20, 7)
    assert isinstance(Q, m.heston_process) and Q.params['rho'] == -.7


def test_python_stepping_hooks_fail_loudly():
    """A user-written ``next`` (the reference's plug point for Python
    integration schemes, quickguide.rst:474-498) or a per-step ``info_store``
    (``let`` / ``info_next`` of traced equations ARE compiled, see
    test_user_let_and_info_next_are_traced) cannot run inside the kernel: NotImplementedError before any device work,
    never a silent Euler run."""
    import sdepy_b200 as m

    class my_integrator(m.integrator):
        def next(self):
            pass

    class my_SDE(m.SDE):
        def sde(self, t, x):
            return {'dt': 0, 'dw': x}

    class rk(my_SDE, my_integrator):
        pass

    class chatty(my_SDE, m.integrator):
        def info_store(self):
            pass

    class tuned(m.lognorm_process):       # hand-written functor: nothing can be compiled in
        def info_next(self):
            pass

    with pytest.raises(NotImplementedError, match='info_next'):
        tuned(paths=10, steps=5)((0., 1.))
    for cls, hook in ((rk, 'next'), (chatty, 'info_store')):
        with pytest.raises(NotImplementedError, match=hook):
            cls(paths=10, x0=1., steps=5)((0., 1.))


def test_sde_return_value_errors_follow_reference():
    """Run-time error conventions of the reference (tests/test_integrator.py:
    416-485): TypeError for a non-dict / non-tuple sde value, KeyError for an
    unknown differential, ValueError for a wrong number of equations,
    IndexError for a bad i0, ValueError for a bad timeline.  All raised while
    the equation is traced, before any device work."""
    import sdepy_b200 as m
    t = np.linspace(0., 1., 5)

    class f_process(m.SDE, m.integrator):
        def sde(self, t, x):
            return {'dt': 1, 'dw': 1}

    f_process.sde = lambda self, t, x: x
    with pytest.raises(TypeError):
        f_process(x0=1, paths=11, steps=30)(t)
    f_process.sde = lambda self, t, x: {'dt': 1, 'dzzz': 1}
    with pytest.raises(KeyError):
        f_process(x0=1, paths=11, steps=30)(t)

    @m.integrate(q=2, sources=('dt', 'dw'))
    def g_process(t, x, y):
        return {'dt': 1, 'dw': 1}, {'dt': 1}
    assert g_process.q == 2
    g_process.sde = lambda self, t, x, y: {'dt': 1, 'dw': 1}
    with pytest.raises(TypeError):
        g_process(x0=(1,)*2, paths=11, steps=30)(t)
    g_process.sde = lambda self, t, x, y: ({'dt': 1, 'dw': 1},)
    with pytest.raises(ValueError):
        g_process(x0=(1,)*2, paths=11, steps=30)(t)
    g_process.sde = lambda self, t, x, y: ({'dt': 1}, {'dzzz': 1})
    with pytest.raises(KeyError):
        g_process(x0=(1,)*2, paths=11, steps=30)(t)

    # declared q / sources: no test evaluation of the function
    @m.integrate(q=0, sources={'dt', 'dw'})
    def never_called(t, x):
        raise ValueError

    def bad_q():
        @m.integrate(q=1)
        def h(t, x=1., y=1.):
            return {'dt': 1}, {'dw': 1}

    def bad_sources():
        @m.integrate(sources={'dw'})
        def h(t, x=1., y=1.):
            return {'dt': 1}, {'dw': 1}

    def failing():
        @m.integrate
        def h(t, x):
            raise ValueError

    for err in (bad_q, bad_sources, failing):
        with pytest.raises(TypeError):
            err()

    P = m.lognorm_process(paths=3, i0=7)
    with pytest.raises(IndexError):
        P(t)
    for bad in (np.zeros((2, 2)), (0., 1., .5)):
        with pytest.raises(ValueError):
            m.lognorm_process(paths=3)(bad)


def test_user_let_and_info_next_are_traced():
    """`let` and `info_next` are pure functions of the state: they are traced
    like `sde` (reference plug points integration.py:1501-1528, 1562-1567) and
    become the kernel's emit() / per-path counters.  CPU check of the tracer and
    of the generated functor source (the GPU run against the reference-made
    fixture is tests/test_gpu_jit.py)."""
    import sdepy_b200 as m
    from sdepy_b200 import _jit
    from tests.cases import user_hooks_single, user_hooks_system
    A = user_hooks_single(m)(paths=5, vshape=(2,), steps=7, x0=.3)
    tr, kind, nodes = A._trace_let(0.)
    assert kind == 'vars' or kind == 'single'
    x = np.array([-.5, 2.])
    assert np.array_equal(_jit.evaluate(nodes[0], [x]), x*x + 1.)
    A.info_begin()
    cnts = A._trace_info()
    assert [k for k, _, _ in cnts] == ['neg', 'big']
    assert np.array_equal(_jit.evaluate(cnts[0][2], [x, x]), (x < 0)*1.)
    assert np.array_equal(_jit.evaluate(cnts[1][2], [x*0, x]), (x > .25)*1.)
    assert isinstance(A.info['neg'], np.ndarray)          # the real dict is back
    t, roots = A._trace(0.)
    src = A._codegen(t, roots, [None], (tr, kind, nodes), cnts)
    assert 'NCNT = 2' in src and 'cnt[0 + 0*E + h] += (int)' in src
    assert 'cnt[0 + 1*E + h] += (int)' in src and 'xn0' in src.split('cnt[0 + 1*E')[1][:80] + src
    B = user_hooks_system(m)(paths=5, vshape=(3,), steps=7, x0=(1., .8))
    assert B.xshape == (3,) and B.wshape == (6,)
    tr, kind, nodes = B._trace_let(0.)
    assert kind == 'single' and len(nodes) == 1
    assert np.array_equal(_jit.evaluate(nodes[0], [np.array(2.), np.array(.25)]), .5)
    t, roots = B._trace(0.)
    src = B._codegen(t, roots, [None, None], (tr, kind, nodes), B._trace_info())
    assert 'NX = 1' in src and 'v[h] = ' in src

    class bad(m.SDE, m.integrator):
        def sde(self, t, x):
            return {'dt': -x, 'dw': 1.}

        def info_next(self):
            self.info['n'] += self.itervars['last_dZ']['dw']

    with pytest.raises(NotImplementedError, match='last_dZ'):
        bad(paths=3)._trace_info()
