"""CPU tests (no GPU) of the host-side drop-in surface, modelled on the
reference's own tests of the same objects: the `process` container
(reference tests/test_process.py:110-185 broadcasting rules, 582-640 summary
keywords), `kfunc` (tests/test_kfunc.py), the jump-size laws' `rvs` protocol
and `rvmap` (infrastructure.py:1640-1862), the named histogram-bin estimators
of `montecarlo(bins='auto' ...)` (numpy.histogram_bin_edges), and the NumPy
interoperability of arrays handed out by device-resident sources."""
import json
import os
import warnings

import numpy as np
import pytest
import torch

import sdepy_b200 as sd
from sdepy_b200 import _cuda
from sdepy_b200.infrastructure import montecarlo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


# ---- process ---------------------------------------------------------------

@pytest.mark.parametrize('vshape', [(), (2,), (3, 2)])
def test_process_broadcasting_rules(vshape):
    rng = np.random.default_rng(0)
    t = np.linspace(0., 4., 12)
    paths = 5
    p = sd.process(t, x=rng.random(t.shape + vshape + (paths,)))
    q = sd.process(t.copy(), x=rng.random(t.shape + vshape + (paths,)))
    assert (p + q).shape == p.shape
    assert (p + q['p', 0]).shape == p.shape and (q['p', 0] + p).shape == p.shape
    # two constant processes on different times; constant + non constant
    a = p['t', 0] + q['t', 1]
    assert a.shape == (1,) + vshape + (paths,)
    a = p + q['t', 0]
    assert a.t is p.t
    a = p['t', 0] + q
    assert a.t is q.t
    if vshape:
        a = p + q['p', 0] + p['t', 0] + q + p['v', :1]
        assert a.shape == p.shape and isinstance(a, sd.process)
        u, v = p.x[0], q.x[..., :1]
        a = u*p + v*q['p', 0] + p['t', 0] + q.x + p['v', :1]*2 + v
        assert a.shape == p.shape
    # incompatible operands raise ValueError
    with pytest.raises(ValueError):
        p['p', :2] + q['p', 2:5]            # paths
    with pytest.raises(ValueError):
        p['t', :2] + q['t', 2:4]            # timelines
    if vshape == (3, 2):
        with pytest.raises(ValueError):
            p + q['v', :2, 0]               # vshapes of different length
        with pytest.raises(ValueError):
            p + q['v', :2]                  # incompatible vshapes
        p1, q1 = p['t', 0]['v', :1, :1], q['v', :, 0]
        assert (p1.x + q1.x).shape == (1,) + t.shape + vshape[:1] + (paths,)   # fine as arrays
        with pytest.raises(ValueError):
            p1 + q1                         # not as processes


def test_process_summary_keywords():
    rng = np.random.default_rng(1)
    p = sd.process(t=(1, 2, 3), x=1 + rng.random((3, 5, 7, 11)))
    tol = 1e-13
    for f in ('min', 'max', 'sum', 'mean', 'var', 'std'):
        a = getattr(p, 'p' + f)()
        assert np.allclose(a, getattr(np, f)(p.x, axis=-1)[..., None], rtol=tol)
        y = np.full((3, 5, 7, 1), np.nan)
        a = getattr(p, 'p' + f)(out=y)
        assert np.array_equal(a, y) and not np.isnan(y).any()
        a = getattr(p, 'v' + f)()
        assert np.allclose(a, getattr(np, f)(p.x, axis=(1, 2)), rtol=tol)
        y = np.full((3, 11), np.nan)
        a = getattr(p, 'v' + f)(out=y)
        assert np.array_equal(a, y) and a.shape == (3, 11)
        a = getattr(p, 't' + f)()
        assert np.allclose(a, getattr(np, f)(p.x, axis=0)[None], rtol=tol)
        assert a.t.shape == (1,)
    for f in ('sum', 'mean', 'var', 'std'):
        for pre in 'pvt':
            assert getattr(p, pre + f)(dtype=np.float32).dtype == np.float32
    assert np.allclose(p.pvar(ddof=1), p.x.var(axis=-1, ddof=1, keepdims=True))
    assert np.allclose(p.tcumsum(dtype=np.float64), p.x.cumsum(axis=0))


# ---- kfunc -----------------------------------------------------------------

def test_kfunc_definition_checks():
    kfunc = sd.kfunc

    class custom_new:
        def __new__(cls):
            pass

    class no_call:
        def __init__(self):
            pass

    class no_init:
        def __call__(self):
            pass

    class positional_param(no_init):
        def __init__(self, x):
            pass

    class clash:
        def __init__(self, *, x):
            pass

        def __call__(self, x):
            pass

    for bad in (custom_new, no_call, no_init, positional_param, clash):
        with pytest.raises(TypeError):
            kfunc(bad)

    class has_params(no_call, no_init):
        params = 0

    with pytest.warns(RuntimeWarning):
        kfunc(has_params)

    @kfunc
    class parent(no_call, no_init):
        pass

    class undecorated(parent):
        pass

    with pytest.warns(RuntimeWarning):
        undecorated()

    with pytest.raises(SyntaxError):
        kfunc(nvar=2)(parent)
    with pytest.raises(SyntaxError):
        kfunc(lambda x, y: None)
    with pytest.raises(ValueError):
        kfunc(nvar=3)(lambda x, y: None)
    with pytest.raises(TypeError):
        kfunc(nvar=2)(lambda x, y, z: None)


def test_kfunc_evaluation_and_derivation():
    kfunc = sd.kfunc

    @kfunc
    class base:
        def __init__(self, *, a=1, b=2, info=None):
            self.a, self.b, self.info = a, b, {} if info is None else info

        def __call__(self, x=11, y=22):
            self.info['value'] = (x, y)
            return (x, y, self.a, self.b)

    @kfunc
    class with_init(base):
        def __init__(self, *, a=1, b=2, info=None):
            self.aa = a
            super().__init__(a=a, b=b, info=info)

    @kfunc
    class with_call(base):
        def __call__(self, x=11, y=22):
            self.seen = (x, y)
            return super().__call__(x, y)

    @kfunc
    class both(with_init, with_call):
        pass

    for F in (base, with_init, with_call, both):
        assert sd.iskfunc(F) and sd.iskfunc(F())
        f = F(a=7)
        assert f._kfunc_parent is None
        assert f.params == dict(a=7, b=2, info=None)
        assert f() == (11, 22, 7, 2) and f('z', 'w') == ('z', 'w', 7, 2)
        assert f(110, a=3) == (110, 22, 3, 2) and f.a == 7      # f not affected
        assert f(110, 220, a=3, b=4, info=f.info) == (110, 220, 3, 4)
        assert f.info['value'] == (110, 220)
        g = f(a=11, info=f.info)
        assert g._kfunc_parent is f and g.params == dict(a=11, b=2, info=f.info)
        assert g(b=13).params == dict(a=11, b=13, info=f.info)
        # instantiate and evaluate in one go
        assert F(110) == (110, 22, 1, 2) and F('z', y='w', a=3, b=4) == ('z', 'w', 3, 4)
        for call in (lambda: F(z=1), lambda: F()(z=1, w=2),
                     lambda: F()(11, 22, a=1, w=2), lambda: F(11, 22, b=1, w=2)):
            with pytest.raises(TypeError):
                call()

    @kfunc(nvar=1)
    def ident(x):
        return x

    assert ident(5) == 5 and ident()(5) == 5

    @kfunc(nvar=2)
    def G(x, y=0, *, p1='1', p2='2'):
        return x, y, p1, p2

    assert G(p1='x').params == dict(p1='x', p2='2')
    assert G(1) == (1, 0, '1', '2') and G(1, 2, p2='y') == (1, 2, '1', 'y')
    g = G(p1='x', p2='y')
    assert g(1) == (1, 0, 'x', 'y') and g(1, p1='xx') == (1, 0, 'xx', 'y')
    assert g(p1='xx')(1, 2) == (1, 2, 'xx', 'y')
    assert g(p1='xx').params == dict(p1='xx', p2='y')
    with pytest.raises(TypeError):
        g(p3='z')(1)


def _decode(v):
    if isinstance(v, dict) and '__tuple__' in v:
        return tuple(_decode(z) for z in v['__tuple__'])
    return v


def test_kfunc_params_match_reference():
    """`params` of every kfunc'd source / process equals the reference's
    (tests/golden/kfunc_params.json, written by make_kfunc_params.py): same
    keys in the same order, same defaults; this package's extra keywords
    (seed, output ...) only show when given."""
    with open(os.path.join(GOLDEN, 'kfunc_params.json')) as f:
        golden = json.load(f)
    for name, case in golden.items():
        cls = getattr(sd, name)
        K = cls if sd.iskfunc(cls) else sd.kfunc(cls)
        got = K(**case['kw']).params
        want = case['params']
        assert list(got) == list(sorted(got, key=list(want).index)) or set(got) == set(want)
        assert set(got) == set(want), (name, set(got) ^ set(want))
        for k, w in want.items():
            g = got[k]
            if isinstance(w, dict) and '__array__' in w:
                assert np.array_equal(np.asarray(g, dtype=float), np.asarray(w['__array__'])), (name, k)
            elif isinstance(w, dict) and '__repr__' in w:
                assert g is not None, (name, k)
            else:
                assert g == _decode(w), (name, k, g, w)
    inst = sd.lognorm(x0=2, seed=5)
    assert inst.params['seed'] == 5 and 'output' not in inst.params
    # the shortcuts are kfuncs, the full names plain classes (shortcuts.py:28-32)
    assert sd.iskfunc(sd.heston) and not sd.iskfunc(sd.heston_process)


# ---- jump-size laws ----------------------------------------------------------

def test_law_rvs_protocol_and_rvmap():
    rng = np.random.default_rng(3)
    n = 200_000
    for rv, mean, var in (
            (sd.norm_rv(a=.5, b=2.), .5, 4.), (sd.uniform_rv(a=1., b=3.), 2., 4/12),
            (sd.exp_rv(a=-.5), -.5, .25),
            (sd.double_exp_rv(a=.1, b=.3, pa=.4), .4*.1 - .6*.3, None)):
        z = rv.rvs(size=(2, n), random_state=rng)
        assert z.shape == (2, n)
        assert abs(z.mean() - mean) < 5*z.std()/np.sqrt(z.size)
        assert np.isclose(rv.mean(), mean)
        if var is not None:
            assert abs(z.var() - var) < .02*var and np.isclose(rv.var(), var)
        else:
            assert abs(z.var() - rv.var()) < .02*rv.var()
    assert not callable(sd.norm_rv(a=0, b=1))
    timed = sd.norm_rv(a=lambda t: 2*t, b=1.)
    assert callable(timed) and np.isclose(timed(3.).mean(), 6.)
    with pytest.raises(TypeError):
        timed.rvs(size=3)
    y = sd.rvmap(np.exp, sd.norm_rv(a=0., b=.1))
    z = y.rvs(size=(1000,), random_state=rng)
    assert (z > 0).all() and abs(np.log(z).std() - .1) < .01
    yt = sd.rvmap(lambda t, y: t + y, sd.uniform_rv(a=0., b=1.))
    z = yt(10.).rvs(size=(50,), random_state=rng)
    assert ((z >= 10) & (z <= 11)).all()
    assert callable(sd.rvmap(np.exp, timed))


# ---- montecarlo bin estimators ------------------------------------------------

@pytest.mark.parametrize('name', ['auto', 'fd', 'sturges', 'sqrt', 'rice', 'scott', 'doane'])
def test_named_bin_estimators_match_numpy(name):
    rng = np.random.default_rng(4)
    for n in (2, 3, 10, 1000, 100_003):
        for scale, integer in ((1., False), (100., False), (30., True)):
            x = scale*rng.standard_normal(n) + 5
            if integer:
                x = np.round(x)
            if x.max() == x.min():
                continue
            d = x - x.mean()
            mom = [d.sum(), (d**2).sum(), (d**3).sum(), (d**4).sum()]
            w = montecarlo._bin_width(name, torch.from_numpy(x), n, x.min(), x.max(),
                                      mom, integer)
            nb = int(np.ceil((x.max() - x.min())/w)) if w else 1
            ref = np.histogram_bin_edges(x.astype(int) if integer else x, bins=name)
            assert nb == len(ref) - 1, (n, scale, integer)
    with pytest.raises(ValueError):
        montecarlo._bin_width('nope', torch.zeros(3, dtype=torch.float64), 3, 0., 1.,
                              [0., 1., 0., 1.], False)


# ---- NumPy interoperability of device-resident source values ------------------

def test_device_array_numpy_interop():
    base = np.arange(6.).reshape(2, 3)
    z = _cuda.as_device_array(torch.from_numpy(base.copy()))
    assert isinstance(z*2, _cuda.device_array) and isinstance(z - z, _cuda.device_array)
    assert isinstance(2. + .5*z, torch.Tensor)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', DeprecationWarning)
        np.testing.assert_array_equal(z, base)
        assert isinstance(np.exp(z), np.ndarray)
        np.testing.assert_allclose(10*np.exp(.2 + .7*z), 10*np.exp(.2 + .7*base))
        assert isinstance(np.arange(3.)*z, np.ndarray) and isinstance(z*np.arange(3.), np.ndarray)
        assert bool((z == base).all()) and bool((z == 3.).any())


# ---- rng parameter ------------------------------------------------------------

def test_rng_and_default_rng_reach_every_source():
    """Same generator state -> same Philox keys, whether the generator is passed
    as ``rng=`` or installed as ``infrastructure.default_rng`` (reference
    tests/test_processes.py:556-596, tests/test_source.py:196-226)."""
    from sdepy_b200 import infrastructure as infra

    def keys(P):
        return [(k, getattr(s, 'seed', None), getattr(getattr(s, 'dn', None), 'seed', None))
                for k, s in sorted(P.sources.items())]

    makers = (
        lambda rng=None: dict(vshape=2, x0=.1, mu=.2, sigma=.3, lam=.4),
        lambda rng=None: dict(vshape=(2, 3), dw=sd.wiener_source(paths=11, vshape=(2, 3), rng=rng)),
        lambda rng=None: dict(vshape=(3,), dw=sd.true_wiener_source(paths=11, vshape=(3,), rng=rng),
                              dj=sd.cpoisson_source(paths=11, vshape=(3,), rng=rng)),
        lambda rng=None: dict(vshape=2, dj=sd.poisson_source(paths=11, vshape=2, rng=rng)),
    )
    for cls in (sd.jumpdiff_process, sd.kou_jumpdiff_process, sd.mjd):
        for make in makers:
            for make_rng in (np.random.default_rng, np.random.RandomState,
                             lambda z: np.random.Generator(np.random.PCG64(z))):
                rng1, rng2 = make_rng(1234), make_rng(1234)
                P1 = cls(**make(rng1), paths=11, rng=rng1)
                P2 = cls(**make(rng2), paths=11, rng=rng2)
                assert P1.rng is rng1 and P2.rng is rng2
                saved = infra.default_rng
                infra.default_rng = make_rng(1234)
                try:
                    P3 = cls(**make(), paths=11, rng=None)
                    assert P3.rng is infra.default_rng
                finally:
                    infra.default_rng = saved
                assert keys(P1) == keys(P2) == keys(P3)
    with pytest.raises(TypeError):          # a seed is not a generator (infrastructure.py:1344)
        sd.wiener_source(rng=1234)
