"""GPU tests of drop-in behaviours the reference's own test-suite exercises
(its tests/test_source.py:160-310, test_processes.py:443-560 and 681-708,
test_montecarlo.py:40-95), re-stated here: source calls over arrays of times
and with a requested dtype, sources with memory returning host arrays and
sub-sources by indexing, single-point timelines with generic sources, named
bin estimators and result dtype of montecarlo."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def sd():
    import sdepy_b200
    return sdepy_b200


TIMES = [2., (5.,), (0, 1.), np.linspace(0.5, 4., 12), np.arange(30.).reshape(2, 3, 5)]


@pytest.mark.parametrize('dtype', [None, np.float32])
def test_float_sources_over_time_arrays_and_dtype(dtype):
    m = sd()
    corr = ((1., -.5), (-.5, 1.))
    cases = [
        (m.wiener_source, dict(vshape=())), (m.wiener_source, dict(vshape=(3, 2), corr=corr)),
        (m.true_wiener_source, dict(vshape=(2, 3))),
        (m.true_wiener_source, dict(vshape=(3, 2), corr=corr)),
        (m.odd_wiener_source, dict(vshape=(2,), corr=corr)),
        (m.cpoisson_source, dict(vshape=(3, 2), lam=((1.,), (2.,)), y=m.norm_rv(a=1., b=5.),
                                 ptype=np.int16)),
        (m.even_cpoisson_source, dict(vshape=(3, 2), lam=((1.,), (2.,)))),
        (m.dw, dict(vshape=())), (m.true_dw, dict(vshape=())), (m.dj, dict(vshape=())),
    ]
    for cls, kw in cases:
        src = cls(paths=10, dtype=dtype, **kw)
        assert src.paths == 10 and src.vshape == kw['vshape']
        for dt in TIMES:
            s = src(0*np.asarray(dt), dt)
            assert isinstance(s, np.ndarray)
            assert s.shape == np.asarray(dt).shape + src.vshape + (10,)
            assert s.dtype == np.dtype(float if dtype is None else dtype)


@pytest.mark.parametrize('dtype', [int, np.int16])
def test_poisson_sources_over_time_arrays_and_dtype(dtype):
    m = sd()
    for cls, kw in ((m.poisson_source, dict(vshape=(2,), lam=((1.,), (2.,)))),
                    (m.poisson_source, dict(vshape=(2, 3, 5), lam=np.arange(15.).reshape(3, 5, 1))),
                    (m.even_poisson_source, dict(vshape=())), (m.dn, dict(vshape=()))):
        src = cls(paths=10, dtype=dtype, **kw)
        for dt in TIMES:
            s = src(0*np.asarray(dt), dt)
            assert s.shape == np.asarray(dt).shape + src.vshape + (10,)
            assert s.dtype == np.dtype(dtype)
            assert (s >= 0).all()


def test_true_wiener_source_memory_indexing_and_numpy_results():
    m = sd()
    t0, z0 = 1., 3.
    t1 = t0 + .2
    src = m.true_wiener_source(vshape=(2, 3), paths=10, t0=t0, z0=z0, seed=3)
    assert (src(t0) == z0).all()
    size1 = src.size
    s = src(t1)
    assert isinstance(s, np.ndarray) and src.size >= 2*size1
    np.testing.assert_array_equal(src(t1), s)                 # memory
    for dt in (-.1, -1e-6, 0, 1e-6, .1):
        z = src(t1 + dt)
        np.testing.assert_allclose(z, src(t1) + src(t1, dt), rtol=1e-13)
        assert (np.abs(z - s) <= 5*np.sqrt(abs(dt))).all()
    dt = .1
    assert src[0, 0](t1).shape == (10,)
    assert src[:, :2](t1, dt).shape == (2, 2, 10)
    np.testing.assert_array_equal(src[:1](t1, dt), src(t1, dt)[:1])
    sub = src[:, :, np.newaxis]
    assert (sub.paths, sub.vshape) == (10, (2, 3, 1)) and sub(t0).shape == (2, 3, 1, 10)
    assert sub[0](t1).shape == (3, 1, 10)
    # the kernels get CUDA tensors that NumPy also understands
    d = src.device_call(t1, dt)
    assert isinstance(d, torch.Tensor) and d.is_cuda
    np.testing.assert_array_equal(d, src(t1, dt))
    np.testing.assert_allclose(np.exp(.5*d), np.exp(.5*src(t1, dt)))


def test_exact_wiener_and_lognorm_on_a_shared_true_source():
    """reference tests/test_processes.py:681-708, written with ndarray
    arithmetic on the source's values as a user of the reference would."""
    m = sd()
    paths, x0, mu, sigma, t0, DT = 31, 10, .2, .7, 1, 3
    dw = m.true_wiener_source(paths=paths, seed=9)
    pw = m.wiener_process(x0=x0, mu=mu, sigma=sigma, paths=paths, dw=dw)
    pl = m.lognorm_process(x0=x0, mu=mu, sigma=sigma, paths=paths, dw=dw)
    xw_exact = x0 + mu*DT + sigma*dw(t0, DT)
    xl_exact = x0*np.exp((mu - sigma*sigma/2)*DT + sigma*dw(t0, DT))
    for t in ((t0, t0 + DT/2, t0 + DT), np.linspace(t0, t0 + DT, 100)):
        np.testing.assert_allclose(xw_exact, pw(t)[-1], rtol=1e-13)
        np.testing.assert_allclose(xl_exact, pl(t)[-1], rtol=1e-13)


def test_single_point_timeline_with_generic_sources():
    """No steps to take: the run returns the initial state whatever feeds the
    differentials (reference tests/test_processes.py:458-480, tlist[0])."""
    m = sd()
    Z2 = m.wiener_process(sigma=.1, paths=11, vshape=2, seed=1)(np.linspace(0, 1, 100))

    def S2(s, ds):
        return np.random.default_rng(0).normal(size=(2, 11))*np.sqrt(abs(ds))
    S2.vshape, S2.paths = (2,), 11
    for dw in (Z2, S2, m.true_wiener_source(paths=11, vshape=2, seed=2)):
        for cls, x0 in ((m.wiener_process, 0.), (m.lognorm_process, 1.),
                        (m.ornstein_uhlenbeck_process, 0.), (m.jumpdiff_process, 1.),
                        (m.merton_jumpdiff_process, 1.)):
            for steps in (None, 5):
                for t in ((0.,), (1, 2), np.linspace(0, 1, 11)):
                    p = cls(paths=11, vshape=2, dw=dw, steps=steps)(t)
                    assert isinstance(p, m.process)
                    assert p.shape == (np.size(t), 2, 11) and p.vshape == (2,)
                    assert (np.asarray(p)[0] == x0).all()
                    np.testing.assert_array_equal(p.t, np.asarray(t).reshape(-1))


@pytest.mark.parametrize('dtype', [None, np.float32, int])
def test_montecarlo_named_bins_and_result_dtype(dtype):
    m = sd()
    rng = np.random.default_rng(5)
    shape, paths = (3, 2), 10
    a = m.montecarlo(bins='auto')
    with pytest.raises(ValueError):
        a.histogram()
    first = None
    for i in range(10):
        sample = (100*rng.normal(size=shape + (paths,))).astype(dtype)
        sample *= (1 + np.arange(6)).reshape(shape + (1,))
        first = sample if first is None else first
        a.update(sample)
    assert a.paths == 100 and a.vshape == shape
    want = np.dtype(dtype) if np.dtype(dtype).kind == 'f' else np.dtype(float)
    assert a.mean().dtype == want and a.stderr().dtype == want
    for i in np.ndindex(shape):
        counts, edges = a[i].histogram()
        ref = np.histogram_bin_edges(first[i], bins='auto')
        assert len(edges) == len(ref)
        # numpy lays float32 edges out in float32
        np.testing.assert_allclose(edges, ref, rtol=1e-12 if dtype is not np.float32 else 1e-6)
        assert counts.sum() + a[i].outpaths == 100
    for name in ('fd', 'sturges', 'sqrt', 'rice', 'scott', 'doane'):
        x = rng.normal(size=5000)
        b = m.montecarlo(x, bins=name)
        counts, edges = b.histogram()
        ref_counts, ref_edges = np.histogram(x, bins=name)
        np.testing.assert_allclose(edges, ref_edges, rtol=1e-12)
        np.testing.assert_array_equal(counts, ref_counts)


def test_many_component_heston_with_supplied_sources():
    """reference tests/test_processes.py:216-234: full_heston_process with 3, 5
    and 10 components, correlated, or driven by a supplied 2N-component source
    (NVRTC instantiation; the parameter records of time-invariant models are
    not staged in shared memory, which leaves room for the replay ring)."""
    m = sd()
    t = np.linspace(0, 1, 11)
    rng = np.random.default_rng(0)
    cm6 = np.eye(6) + .1*rng.random((6, 6))
    cm6 = (cm6 + cm6.T)/2
    cases = (
        dict(vshape=(3,), corr=cm6), dict(vshape=(3,), rho=(.1, .2, .3)),
        dict(vshape=(5,), dw=m.wiener_source(paths=11, vshape=(10,), seed=1)),
        dict(vshape=10, dw=m.true_wiener_source(paths=11, vshape=(20,), seed=2)),
    )
    for kw in cases:
        x, y = m.full_heston_process(paths=11, steps=30, **kw)(t)
        n = kw['vshape'] if isinstance(kw['vshape'], int) else kw['vshape'][0]
        assert x.shape == y.shape == (11, n, 11)
        assert np.isfinite(np.asarray(x)).all() and (np.asarray(x) > 0).all()
    # a failed launch must not poison the next call
    # (20 correlated components, time-dependent records AND a replay ring: 220 KB)
    bad = m.full_heston_process(paths=11, vshape=10, steps=30, theta=lambda t: .04 + t,
                                dw=m.true_wiener_source(paths=11, vshape=(20,), seed=3))
    with pytest.raises(Exception):
        bad(t)
    x = m.lognorm_process(paths=11, steps=5, seed=1)(t)
    assert np.isfinite(np.asarray(x)).all()
