"""GPU edge cases the reference's own tests exercise (sdepy/tests/
test_integrator.py:63-178, test_processes.py:443-596, test_source.py:164-284):
degenerate timelines, single paths, ragged path counts, shapes, sources used
stand-alone, cumulation of montecarlo updates."""
import numpy as np
import pytest
import torch

from oracle import sde_oracle as orc

pytestmark = pytest.mark.gpu


def sd():
    import sdepy_b200
    return sdepy_b200


def test_single_time_point_and_no_steps():
    m = sd()
    x = m.ornstein_uhlenbeck_process(paths=7, x0=3.)((0.5,))
    assert x.shape == (1, 7) and np.array_equal(np.asarray(x), np.full((1, 7), 3.))
    # steps=None: the timeline itself is the grid (integration.py:213-214)
    P = m.wiener_process(paths=5, x0=1., mu=2., sigma=0.)
    x = P((0., .25, 1.))
    assert np.allclose(np.asarray(x), np.array([1., 1.5, 3.])[:, None])
    assert P.info['computed_steps'] == 2 and P.info['stored_steps'] == 2
    # integer timeline is integrated in float, returned as given (529-530, 584)
    x = m.wiener_process(paths=3, x0=0., mu=1., sigma=0.)((0, 1, 2))
    assert x.t.dtype.kind == 'i' and np.allclose(np.asarray(x)[:, 0], [0., 1., 2.])


@pytest.mark.parametrize('paths', [1, 31, 32, 33, 255, 256, 257, 1000])
def test_ragged_path_counts_replay(paths):
    m = sd()
    rng = np.random.default_rng(paths)
    n = 70                      # crosses the 64-step staging block
    grid = np.linspace(0., 2., n + 1)
    dW = rng.standard_normal((n, paths))*np.sqrt(np.diff(grid))[:, None]
    par = dict(theta=.3, k=1.2, sigma=.5)
    out, _ = orc.euler_replay('ornstein_uhlenbeck', par, .1, grid, [0, 10, 64, 65, n], dW)
    x = m.ornstein_uhlenbeck_process(paths=paths, steps=grid, x0=.1,
                                     dw=m.replay_source(dW), **par)(grid[[0, 10, 64, 65, n]])
    assert np.array_equal(np.asarray(x), out)


def test_vshape_groups_and_per_component_parameters():
    m = sd()
    rng = np.random.default_rng(1)
    paths, n = 50, 12
    grid = np.linspace(0., 1., n + 1)
    vshape = (2, 3)
    dW = rng.standard_normal((n,) + vshape + (paths,))*np.sqrt(1/n)
    mu = np.arange(6.).reshape(2, 3, 1)*.01
    sigma = .1 + np.arange(6.).reshape(2, 3, 1)*.05
    x0 = 1. + np.arange(3.).reshape(3, 1)
    out, _ = orc.euler_replay('lognorm', dict(mu=mu, sigma=sigma), x0, grid, [0, n], dW)
    x = m.lognorm_process(paths=paths, vshape=vshape, steps=grid, x0=x0, mu=mu,
                          sigma=sigma, dw=m.replay_source(dW))((0., 1.))
    assert x.shape == (2, 2, 3, paths)
    assert np.abs(np.asarray(x)/out - 1).max() <= 4*np.finfo(float).eps
    # philox: every (component, path) lane gets its own stream
    y = np.asarray(m.lognorm_process(paths=2000, vshape=vshape, steps=20, x0=1., mu=0.,
                                     sigma=.2, seed=3)((0., 1.)))[-1]
    c = np.corrcoef(np.log(y).reshape(6, -1))
    assert np.abs(c - np.eye(6)).max() < 5/np.sqrt(2000)


def test_per_path_initial_condition():
    m = sd()
    rng = np.random.default_rng(2)
    paths, n = 300, 8
    grid = np.linspace(0., 1., n + 1)
    dW = rng.standard_normal((n, paths))*np.sqrt(1/n)
    x0 = 1. + rng.random(paths)
    out, _ = orc.euler_replay('ornstein_uhlenbeck', dict(theta=0., k=1., sigma=1.), x0,
                              grid, [0, n], dW)
    x = m.ornstein_uhlenbeck_process(paths=paths, steps=grid, x0=x0,
                                     dw=m.replay_source(dW))((0., 1.))
    assert np.array_equal(np.asarray(x), out)


def test_standalone_sources():
    m = sd()
    dw = m.wiener_source(paths=200_000, vshape=(2,), rho=.6, seed=1)
    z = dw(0., .25)
    assert z.shape == (2, 200_000)
    assert abs(z.var() - .25) < .25*5*np.sqrt(2/z.size)
    assert abs(np.corrcoef(z)[0, 1] - .6) < 5/np.sqrt(200_000)
    assert not np.array_equal(z, dw(0., .25))            # the stream advances
    z2 = m.wiener_source(paths=200_000, vshape=(2,), rho=.6, seed=1)(0., .25)
    assert np.array_equal(z, z2)                          # same seed, same draw
    # cpoisson with a (numerically) constant jump size: dj == dn * value
    # (sdepy/tests/test_source.py:257-269)
    dj = m.cpoisson_source(paths=100_000, lam=3., y=m.uniform_rv(a=.7, b=.7), seed=2)
    j = dj(0., 1.)
    assert np.allclose(j, dj.dn_value*.7, rtol=1e-14)
    assert abs(dj.dn_value.mean() - 3.) < 5*np.sqrt(3/100_000)
    dn = m.poisson_source(paths=100_000, lam=.5, seed=4)(0., -2.)
    assert (dn <= 0).all() and abs(dn.mean() + 1.) < 5*np.sqrt(1/100_000)


def test_outputs_agree_across_modes():
    m = sd()
    kw = dict(paths=5000, steps=30, x0=100., y0=.04, mu=.03, sigma=1., theta=.04, k=2.,
              xi=.5, rho=-.6, seed=77)
    xh, yh = m.full_heston_process(**kw)((0., .5, 1.))
    xd, yd = m.full_heston_process(output='device', **kw)((0., .5, 1.))
    assert isinstance(xd, m.device_process) and xd.shape == xh.shape
    assert np.array_equal(np.asarray(xd), np.asarray(xh))
    assert np.array_equal(np.asarray(yd.cpu()), np.asarray(yh))
    st = m.full_heston_process(output='stats', **kw)((0., .5, 1.))
    assert np.allclose(np.asarray(st.pmean())[:, 0, 0], np.asarray(xh).mean(axis=-1), rtol=1e-12)
    assert np.allclose(np.asarray(st.pmean())[:, 1, 0], np.asarray(yh).mean(axis=-1), rtol=1e-12)
    assert np.allclose(np.asarray(st.pstd())[:, 1, 0], np.asarray(yh).std(axis=-1), rtol=1e-10)
    # getinfo=False: no diagnostics, same paths
    P = m.full_heston_process(getinfo=False, **kw)
    x2, _ = P((0., .5, 1.))
    assert np.array_equal(np.asarray(x2), np.asarray(xh)) and 'negative_y_count' not in P.info


def test_reference_style_generic_source_objects():
    """A source written against the reference's protocol only (callable with
    paths/vshape) drives the kernel through host evaluation + replay."""
    m = sd()

    class lattice:
        paths, vshape = 64, ()

        def __call__(self, t, dt):
            k = np.arange(64)
            return np.sqrt(abs(dt))*np.where((k + int(round(t*8))) % 2, 1., -1.)

    x = m.wiener_process(paths=64, steps=9, dw=lattice())((0., 1.))
    grid = np.linspace(0, 1, 9)
    want = sum(lattice()(t, grid[1] - grid[0]) for t in grid[:-1])
    assert np.allclose(np.asarray(x)[-1], want, rtol=1e-14, atol=1e-15)


def test_antithetic_sources_and_montecarlo_use():
    """odd_wiener_source / even_cpoisson_source (reference infrastructure.py:
    2047-2150) and montecarlo(use='even'|'odd') (2905-2914)."""
    m = sd()
    K = 20_000
    P = m.wiener_process(paths=2*K, steps=20, x0=0., mu=0., sigma=1.,
                         dw=m.odd_wiener_source, seed=3)
    x = np.asarray(P((0., 1.)))[-1]
    assert np.array_equal(x[:K], -x[K:])                 # mirrored Brownian paths
    assert abs(x[:K].var() - 1.) < 5*np.sqrt(2/K)
    z = m.odd_wiener_source(paths=10, seed=1)(0., 1.)
    assert z.shape == (10,) and np.array_equal(z[:5], -z[5:])
    with pytest.raises(ValueError):
        m.odd_wiener_source(paths=7)
    # jump diffusion: opposite diffusion, identical jumps
    J = m.merton_jumpdiff_process(paths=2*K, steps=50, x0=1., mu=0., sigma=.2, lam=3.,
                                  a=-.1, b=.2, dw=m.odd_wiener_source,
                                  dj=m.even_cpoisson_source, seed=5)
    J._dump_increments = True
    lx = np.log(np.asarray(J((0., 1.)))[-1])
    d = J._last_run.dump[0]
    dJ, dWd = d['dJ'].cpu().numpy()[:, 0], d['dW'].cpu().numpy()[:, 0]
    assert np.array_equal(dJ[:, :K], dJ[:, K:]) and np.array_equal(dWd[:, :K], -dWd[:, K:])
    assert 'jump_count' not in J.info
    # even part of log x: drift + jumps only (the diffusion cancels exactly up to rounding)
    even = (lx[:K] + lx[K:])/2
    assert np.allclose(even, -.02 + dJ[:, :K].sum(axis=0), rtol=0, atol=1e-12)
    # montecarlo antithetic use
    a = m.montecarlo(lx, use='even')
    b = m.montecarlo(lx, use='odd')
    assert a.paths == K and b.paths == K
    assert np.allclose(a.mean(), even.mean(), rtol=1e-12)
    assert np.allclose(a.var(), even.var(), rtol=1e-10)
    odd = (lx[:K] - lx[K:])/2
    assert np.allclose(b.mean(), odd.mean(), rtol=1e-9, atol=1e-15)
    assert np.allclose(b.std(), odd.std(), rtol=1e-10)
    c, e = np.histogram(even, bins=100)
    assert np.array_equal(a.histogram()[0], c)
    with pytest.raises(ValueError):
        m.montecarlo(lx[:-1], use='even')


def test_path_dependent_and_process_valued_parameters():
    """Parameters shaped (..., paths) and multi-path `process` instances as
    parameters (reference doc/quickguide.rst:127-137, 404-412): replay parity
    with the oracle, which broadcasts them natively."""
    m = sd()
    rng = np.random.default_rng(4)
    paths, n = 333, 20
    grid = np.linspace(0., 1., n + 1)
    dW = rng.standard_normal((n, paths))*np.sqrt(1/n)
    mu, sigma = rng.normal(.05, .02, paths), .1 + .2*rng.random(paths)
    out, _ = orc.euler_replay('lognorm', dict(mu=mu, sigma=sigma), 1., grid, [0, 7, n], dW)
    x = m.lognorm_process(paths=paths, steps=grid, x0=1., mu=mu, sigma=sigma,
                          dw=m.replay_source(dW))(grid[[0, 7, n]])
    assert np.abs(np.asarray(x)/out - 1).max() <= 4*np.finfo(float).eps
    # Heston with a path-dependent vol-of-vol, y bit-exact
    dW2 = rng.standard_normal((n, 2, paths))*np.sqrt(1/n)
    xi = .2 + .6*rng.random(paths)
    par = dict(mu=.03, sigma=1., theta=.04, k=2., xi=xi)
    (ox, oy), oi = orc.euler_replay('heston', par, 100., grid, [0, n], dW2, y0=.04, full=True)
    P = m.full_heston_process(paths=paths, steps=grid, x0=100., y0=.04,
                              dw=m.replay_source(dW2), **par)
    hx, hy = P((0., 1.))
    assert np.array_equal(np.asarray(hy), oy)
    assert np.array_equal(P.info['negative_y_count'], oi['negative_y_count'])
    # time- and path-dependent: a multi-path process as the OU mean level
    tk = np.linspace(0., 1., 5)
    level = m.process(tk, x=rng.normal(.5, .1, (5, paths)).cumsum(axis=0))
    par = dict(theta=level, k=1.5, sigma=.3)
    out, _ = orc.euler_replay('ornstein_uhlenbeck', par, .2, grid, [0, n], dW)
    x = m.ornstein_uhlenbeck_process(paths=paths, steps=grid, x0=.2,
                                     dw=m.replay_source(dW), **par)((0., 1.))
    assert np.array_equal(np.asarray(x), out)

    # traced SDE with a path-dependent parameter, philox mode: per-path drift shows up
    @m.integrate
    def f(t, x, a=0., s=1.):
        return {'dt': a, 'dw': s}
    a = np.where(np.arange(20_000) % 2, 1., -1.)
    y = np.asarray(f(paths=20_000, steps=11, x0=0., a=a, s=.1, seed=2)((0., 1.)))[-1]
    assert abs(y[1::2].mean() - 1.) < .01 and abs(y[0::2].mean() + 1.) < .01


def test_true_wiener_source_memory_bridge_and_convergence():
    """Wiener source with memory on the device (reference infrastructure.py:
    2379-2499, tests/test_source.py:286-410): repeated calls agree, bridged
    values have Brownian covariances, and the same driving path integrated on
    nested grids converges strongly (doc/quickguide.rst:268-294)."""
    m = sd()
    paths = 100_000
    dw = m.true_wiener_source(paths=paths, vshape=(2,), rho=.5, seed=7)
    w1 = dw(1.)
    assert w1.shape == (2, paths)
    assert np.array_equal(dw(1.), w1)               # memory
    assert np.array_equal(dw(0.), np.zeros((2, paths)))
    w05 = dw(.5)                                    # bridge between 0 and 1
    w2 = dw(2.)                                     # extension
    w15 = dw(1.5)                                   # bridge between 1 and 2
    assert np.array_equal(dw(.5, 1.), w15 - w05)
    assert list(dw.t) == [0., .5, 1., 1.5, 2.] and dw.size == 5*2*paths
    se = 5/np.sqrt(paths)
    for a, ta in ((w05, .5), (w1, 1.), (w15, 1.5), (w2, 2.)):
        assert abs(a[0].var()/ta - 1) < se*np.sqrt(2) and abs(np.corrcoef(a)[0, 1] - .5) < se
        for b, tb in ((w05, .5), (w1, 1.), (w15, 1.5), (w2, 2.)):
            cov = np.mean(a[0]*b[0])
            assert abs(cov - min(ta, tb)) < se*np.sqrt(ta*tb + min(ta, tb)**2)
    # increments over disjoint intervals are uncorrelated
    assert abs(np.corrcoef(w05[1], (w1 - w05)[1])[0, 1]) < se
    # time-dependent correlation: matrix bridge keeps unit variances per unit time
    dwt = m.true_wiener_source(paths=paths, vshape=(2,), seed=8,
                               corr=lambda t: np.array(((1., .2 + .3*t), (.2 + .3*t, 1.))))
    e1, emid = dwt(1.), dwt(.4)
    assert abs(emid[0].var()/.4 - 1) < 2*se and abs((e1 - emid)[1].var()/.6 - 1) < 2*se
    assert abs(np.corrcoef(emid)[0, 1] - (.2 + .3*.2)) < 2*se
    # convergence study: one Brownian path, nested grids, exact lognormal solution
    tw = m.true_wiener_source(paths=20_000, seed=9)
    errs = []
    for n in (4, 16, 64):
        x = m.lognorm_process(paths=20_000, steps=n + 1, x0=1., mu=.05, sigma=.5, dw=tw)((0., 1.))
        # Euler on log x is exact for constant parameters: identical terminal values
        errs.append(np.asarray(x)[-1])
    wT = tw(1.)
    exact = np.exp((.05 - .125) + .5*wT)
    for e in errs:
        assert np.allclose(e, exact, rtol=1e-12)

    def f(t, x, mu=.05, sigma=.5):
        return {'dt': mu*x, 'dw': sigma*x}
    strong = [np.abs(np.asarray(m.integrate(f)(paths=20_000, steps=n + 1, x0=1., dw=tw)((0., 1.)))[-1]
                     - exact).mean() for n in (4, 16, 64)]
    assert strong[0] > 1.6*strong[1] > 2.5*strong[2]              # ~ sqrt(dt) strong order


def test_online_statistics_every_row_and_groups():
    """output='stats' on a full timeline and with several lane groups equals
    the reductions of the stored paths (same seed => same paths)."""
    m = sd()
    tl = np.linspace(0., 1., 41)
    kw = dict(paths=30_000, vshape=(2, 3), x0=1., mu=.02, sigma=np.array([.1, .2, .3])[:, None],
              seed=31)
    xd = m.lognorm_process(output='device', **kw)(tl)
    st = m.lognorm_process(output='stats', **kw)(tl)
    assert st.sums.shape == (41, 2, 3, 8)
    assert np.allclose(np.asarray(st.pmean())[..., 0], np.asarray(xd.pmean())[..., 0], rtol=1e-12)
    assert np.allclose(np.asarray(st.pvar())[1:, ..., 0], np.asarray(xd.pvar())[1:, ..., 0], rtol=1e-9)
    assert np.array_equal(np.asarray(st.pmin())[..., 0], np.asarray(xd.pmin())[..., 0])
    assert np.array_equal(np.asarray(st.pmax())[..., 0], np.asarray(xd.pmax())[..., 0])
    xs = np.asarray(xd)
    d = xs - 1.
    assert np.allclose(np.asarray(st.skew())[5:, ..., 0],
                       ((d - d.mean(-1, keepdims=True))**3).mean(-1)[5:]/xs.var(-1)[5:]**1.5, rtol=1e-7)
    # too many rows x components for the in-kernel accumulators: loud failure
    with pytest.raises(NotImplementedError):
        m.wiener_process(paths=10, vshape=(40,), output='stats')(np.linspace(0, 1, 400))


def test_cpoisson_with_user_jump_size_objects():
    """The reference's protocol for user jump-size laws (tests/test_source.py:
    257-284): any object with ``rvs(size, random_state)``; with a constant law
    ``dj == dn*val``.  Counts are drawn on the device, the user's Python
    object is evaluated on the host; an SDE driven by it runs through the
    generic replay path and keeps ``jump_count`` consistent."""
    import scipy.stats
    import sdepy_b200 as m
    val = 2.

    class const_rv:
        def rvs(self, size, random_state):
            return np.full(size, fill_value=val)

    class const_rv_legacy:
        def rvs(self, size):
            return np.full(size, fill_value=val)

    src = m.cpoisson_source(lam=1., paths=100, ptype=np.int16, y=const_rv(), seed=3)
    for dt in (1, 100, -2):
        s, n = src(0, dt), src.dn_value
        assert s.shape == (100,) and np.array_equal(n*val, s)
    assert (src(0, -2) <= 0).all()
    src = m.cpoisson_source(lam=1., paths=100, y=const_rv_legacy(), seed=4)
    with pytest.warns(DeprecationWarning):
        s, n = src(0, 1), src.dn_value
    assert np.array_equal(n*val, s)

    # a frozen scipy law inside a jump-diffusion: log x jumps by N(-.1, .15)
    P = m.jumpdiff_process(x0=1., mu=.05, sigma=.2, lam=20., y=scipy.stats.norm(-.1, .15),
                           paths=4000, steps=50, seed=6, rng=np.random.default_rng(1))
    x = P((0., 1.))
    assert x.shape == (2, 4000) and np.isfinite(x).all()
    assert abs(P.info['jump_count'].mean() - 20.) < 4*np.sqrt(20./4000)
    want = np.exp(.05 - .02 + 20.*(np.exp(-.1 + .15**2/2) - 1))     # E[x(1)]
    sd_log = np.sqrt(.04 + 20.*(.1**2 + .15**2))
    assert abs(np.log(x[-1]).mean() - (.05 - .02 + 20.*(-.1))) < 4*sd_log/np.sqrt(4000)
    assert want > 0


@pytest.mark.parametrize('dtype', [np.float32, np.float16])
def test_float32_float16_storage(dtype):
    """dtype= of the reference (integration.py:495-496; exercised by its
    tests/test_processes.py:122, 323 with float16): the output process has the
    requested dtype, is finite, starts from x0 rounded to it; the arithmetic
    is fp64 in registers, so the narrow run equals the float64 run of the same
    seed rounded once at the store (x0 exactly representable)."""
    m = sd()
    eps = np.finfo(dtype).eps
    tl = np.linspace(0., 1., 11)
    cases = [
        (m.wiener_process, dict(x0=.5, mu=.1, sigma=.3)),
        (m.lognorm_process, dict(x0=1., mu=.1, sigma=.3)),      # log(x0) exact in any dtype
        (m.ornstein_uhlenbeck_process, dict(x0=.25, theta=.2, k=1., sigma=.3)),
        (m.cox_ingersoll_ross_process, dict(x0=.5, theta=.4, k=1., xi=.2)),
        (m.hull_white_process, dict(factors=2, x0=((.5,), (.25,)), sigma=.1)),
        (m.heston_process, dict(x0=1., y0=.25, xi=.3, rho=(-.5, -.3))),
        (m.full_heston_process, dict(x0=1., y0=.25, xi=.3, rho=(-.5, -.3))),
        (m.merton_jumpdiff_process, dict(x0=1., lam=3., a=-.1, b=.2)),
    ]
    for cls, kw in cases:
        for i0 in (0, 3, -2):
            ps = cls(paths=33, vshape=(2,), dtype=dtype, seed=3, steps=30, i0=i0, **kw)(tl)
            ref = cls(paths=33, vshape=(2,), seed=3, steps=30, i0=i0, **kw)(tl)
            ps, ref = (z if isinstance(z, tuple) else (z,) for z in (ps, ref))
            for p, r in zip(ps, ref):
                assert isinstance(p, m.process) and p.dtype == dtype and r.dtype == np.float64
                assert p.shape == (11, 2, 33) and np.isfinite(p).all()
                assert (p[i0] == p[i0, ..., 0][..., np.newaxis]).all()
                np.testing.assert_array_equal(np.asarray(p), np.asarray(r).astype(dtype))
    # x0 that is not representable: rounded to dtype first, like the reference's
    # typed working array
    p = m.lognorm_process(paths=5, dtype=dtype, x0=.1, seed=1)(tl)
    assert np.allclose(np.asarray(p[0], dtype=float), .1, rtol=4*eps)
    # traced equations, and statistics (unaffected by the storage type)
    @m.integrate
    def f(t, x, a=.3):
        return {'dt': -a*x, 'dw': .2}
    p = f(paths=17, dtype=dtype, x0=1., seed=2, steps=20)(tl)
    r = f(paths=17, x0=1., seed=2, steps=20)(tl)
    np.testing.assert_array_equal(np.asarray(p), np.asarray(r).astype(dtype))
    with pytest.raises(NotImplementedError):
        m.wiener_process(paths=5, dtype=np.int32)(tl)
    with pytest.raises(NotImplementedError):
        m.wiener_process(paths=5, dtype=np.float32, output='device')(tl)
