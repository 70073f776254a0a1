"""GPU tests of the traced / NVRTC path: user `@integrate` functions and SDE
subclasses with a Python sde, Euler and Milstein, replay parity against the
reference's golden outputs and the oracle."""
import numpy as np
import pytest
import torch

from oracle import sde_oracle as orc
from tests.cases import golden, theta_t

pytestmark = pytest.mark.gpu
ULP4 = 4*np.finfo(float).eps


def sd():
    import sdepy_b200
    return sdepy_b200


def test_traced_ou_time_dependent_replay_bit_exact():
    m = sd()
    g = golden('replay_oruh_tdep')

    @m.integrate
    def my_ou(t, x, theta=0., k=1., sigma=1.):
        return {'dt': k*(theta - x), 'dw': sigma}

    P = my_ou(paths=g['dW'].shape[-1], steps=g['grid'], x0=.1, theta=theta_t,
              k=1., sigma=.3, dw=m.replay_source(g['dW']))
    x = P(g['tt'])
    assert isinstance(x, m.process)
    assert np.array_equal(np.asarray(x), g['out0'])


def test_traced_cir_and_lognorm_replay():
    m = sd()
    g = golden('replay_cir')

    class my_cir(m.SDE, m.integrator):
        def sde(self, t, x, theta=1., k=1., xi=1.):
            xp = np.maximum(x, 0.)
            return {'dt': k*(theta - xp), 'dw': xi*np.sqrt(xp)}

    P = my_cir(paths=g['dW'].shape[-1], steps=g['grid'], x0=.05, theta=.04,
               k=1.5, xi=.6, dw=m.replay_source(g['dW']))
    assert np.array_equal(np.asarray(P(g['tt'])), g['out0'])

    g = golden('replay_lognorm')

    @m.integrate(log=True)
    def my_lognorm(t, x, mu=0., sigma=1.):
        return {'dt': mu - sigma*sigma/2, 'dw': sigma}

    P = my_lognorm(paths=g['dW'].shape[-1], steps=g['grid'], x0=1., mu=.05,
                   sigma=.2, dw=m.replay_source(g['dW']))
    assert np.abs(np.asarray(P(g['tt']))/g['out0'] - 1).max() <= ULP4


def test_traced_jump_diffusion_replay():
    m = sd()
    g = golden('replay_merton')

    @m.integrate(q=0, sources={'dt', 'dw', 'dj'}, log=True)
    def my_jd(t, x, mu=0., sigma=1.):
        return {'dt': mu - sigma*sigma/2, 'dw': sigma, 'dj': 1}

    P = my_jd(paths=g['dW'].shape[-1], steps=g['grid'], x0=1., mu=.05, sigma=.2,
              dw=m.replay_source(g['dW']),
              dj=m.replay_source(g['dJ'], dn=g['dN']))
    assert np.abs(np.asarray(P(g['tt']))/g['out0'] - 1).max() <= ULP4


def test_traced_system_matches_heston_golden():
    """A q=2 system written by the user on (a = log x, y) -- the preset's own
    equations (reference integration.py:2425-2433) -- must reproduce the
    reference: y bit-exact, exp(a) within 4 ulp."""
    m = sd()
    g = golden('replay_heston_full')

    @m.integrate(q=2, sources={'dt', 'dw'})
    def log_heston(t, a, y, mu=0., sigma=1., theta=1., k=1., xi=1.):
        yp = np.maximum(y, 0.)
        return ({'dt': mu - sigma*sigma*yp/2, 'dw': sigma*np.sqrt(yp)},
                {'dt': k*(theta - yp), 'dw': xi*np.sqrt(yp)})

    P = log_heston(paths=g['dW'].shape[-1], steps=g['grid'],
                   x0=(np.log(100.), .04), mu=.03, sigma=1., theta=.04, k=2.,
                   xi=.3, dw=m.replay_source(g['dW']))
    a, y = P(g['tt'])
    assert np.array_equal(np.asarray(y), g['out1'])
    assert np.abs(np.exp(np.asarray(a))/g['out0'] - 1).max() <= ULP4


@pytest.mark.parametrize('name,addaxis', [('replay_system_stacked', False),
                                          ('replay_system_addaxis', True)])
@pytest.mark.parametrize('output', ['process', 'device'])
def test_user_system_both_stackings_bit_exact(name, addaxis, output):
    """q=2 system over vshape (2, 3) with per-element parameters, equations
    stacked along the last vshape axis (the reference's default) or on a new
    axis: the reference's own output, bit for bit."""
    from tests.cases import user_system
    m = sd()
    g = golden(name)
    cls = m.integrate(q=2, sources={'dt', 'dw'}, addaxis=addaxis)(user_system)
    P = cls(paths=g['dW'].shape[-1], vshape=(2, 3), steps=g['grid'],
            x0=(1., .3), mu=.05, sigma=g['p_sigma'], xi=g['p_xi'],
            dw=m.replay_source(g['dW']), output=output)
    x, y = P(g['tt'])
    if output == 'device':
        x, y = x.x.cpu().numpy(), y.x.cpu().numpy()
    assert np.array_equal(np.asarray(x), g['out0'])
    assert np.array_equal(np.asarray(y), g['out1'])


def test_user_system_stacked_statistics():
    """output='stats' of a variable-major system: same entries as the paths."""
    from tests.cases import user_system
    m = sd()
    g = golden('replay_system_stacked')
    cls = m.integrate(q=2, sources={'dt', 'dw'})(user_system)
    kw = dict(paths=g['dW'].shape[-1], vshape=(2, 3), steps=g['grid'],
              x0=(1., .3), mu=.05, sigma=g['p_sigma'], xi=g['p_xi'])
    st = cls(dw=m.replay_source(g['dW']), output='stats', **kw)(g['tt'])
    ref = np.concatenate((g['out0'], g['out1']), axis=-2)
    assert np.allclose(np.asarray(st.pmean())[..., 0], ref.mean(axis=-1), rtol=1e-13)
    assert np.array_equal(np.asarray(st.pmax())[..., 0], ref.max(axis=-1))


def test_milstein_replay_matches_oracle_and_beats_euler():
    m = sd()
    rng = np.random.default_rng(2)
    paths, n = 20_000, 50
    grid = np.linspace(0., 1., n + 1)
    dW = rng.standard_normal((n, paths))*np.sqrt(np.diff(grid))[:, None]

    def f(t, x, mu=.05, sigma=.4):
        return {'dt': mu*x, 'dw': sigma*x}

    gbm = m.integrate(f)
    par = dict(mu=.05, sigma=.4)
    xm = np.asarray(gbm(paths=paths, steps=grid, x0=1., method='milstein',
                        dw=m.replay_source(dW), **par)((0., 1.)))
    xe = np.asarray(gbm(paths=paths, steps=grid, x0=1.,
                        dw=m.replay_source(dW), **par)((0., 1.)))
    om = orc.generic_replay(f, par, 1., grid, [0, n], dW, scheme='milstein',
                            diffusion_dx=lambda t, x, mu, sigma: sigma)
    oe = orc.generic_replay(f, par, 1., grid, [0, n], dW)
    assert np.array_equal(xe, oe)
    # the oracle's Milstein is pinned to the reference machinery + plug-in fixture
    # (test_milstein_replay_matches_the_reference_machinery_fixture below)
    assert np.abs(xm/om - 1).max() < 1e-12
    exact = np.exp((.05 - .08)*1. + .4*dW.sum(axis=0))
    err_m, err_e = np.abs(xm[-1] - exact).mean(), np.abs(xe[-1] - exact).mean()
    assert err_m < err_e/5


def test_config5_milstein_philox_montecarlo():
    """BASELINE config 5 (scaled): custom @integrate SDE, Milstein, in-kernel
    Philox draws, montecarlo moments + histogram of the terminal value."""
    m = sd()

    @m.integrate
    def f(t, x, mu=.05, sigma=.2):
        return {'dt': mu*x, 'dw': sigma*x}

    paths = 400_000
    x = f(paths=paths, steps=201, x0=1., method='milstein', seed=5,
          output='device')((0., 1.))
    a = m.montecarlo(x.x[-1], bins=100)
    assert abs(float(a.mean()) - np.exp(.05)) < 4*float(a.stderr())
    sdv = np.exp(.05)*np.sqrt(np.exp(.04) - 1)
    assert abs(float(a.std()) - sdv) < 6*sdv*np.sqrt(1/paths)
    counts, edges = a.histogram()
    xT = x.x[-1].cpu().numpy()
    c, e = np.histogram(xT, bins=100)
    assert np.array_equal(edges, e) and np.array_equal(counts, c)
    assert counts.sum() == paths and a.outpaths == 0


def test_untraceable_functions_fail_loudly():
    m = sd()

    @m.integrate(q=0, sources={'dt', 'dw'})
    def branchy(t, x, k=1.):
        return {'dt': k if x > 0 else -k, 'dw': 1.}

    with pytest.raises(TypeError):
        branchy(paths=10, steps=5)((0., 1.))


def test_preset_component_counts_beyond_the_precompiled_ones():
    """7-factor Hull-White with a correlation matrix: not in libsdeb.so, the
    same functor is instantiated by NVRTC.  Replay vs the oracle, bit-exact."""
    m = sd()
    rng = np.random.default_rng(3)
    F, paths, n = 7, 120, 15
    grid = np.linspace(0., 1., n + 1)
    dW = rng.standard_normal((n, F, paths))*np.sqrt(1/n)
    k = (.1 + .2*np.arange(F)).reshape(F, 1)
    sigma = (.01 + .002*np.arange(F)).reshape(F, 1)
    x0 = (.01*np.arange(F)).reshape(F, 1)
    out, _ = orc.euler_replay('hull_white', dict(theta=.02, k=k, sigma=sigma), x0, grid,
                              [0, 5, n], dW)
    x = m.hull_white_process(paths=paths, factors=F, steps=grid, x0=x0, theta=.02, k=k,
                             sigma=sigma, dw=m.replay_source(dW))(grid[[0, 5, n]])
    assert np.array_equal(np.asarray(x), out)
    # philox with a 7x7 correlation
    c = .3*np.ones((F, F)) + .7*np.eye(F)
    P = m.hull_white_process(paths=50_000, factors=F, steps=5, x0=x0, theta=.02, k=k,
                             sigma=sigma, corr=c, seed=1)
    P._dump_increments = True
    P((0., 1.))
    z = P._last_run.dump[0]['dW'].cpu().numpy()[0]
    assert np.abs(np.corrcoef(z) - c).max() < 5/np.sqrt(50_000)


def test_traced_sde_with_plain_poisson_differential():
    """'dn' differentials in user SDEs (reference tests/test_integrator.py:
    dn/dj sources): Philox counts have the right law; replay is exact."""
    m = sd()

    @m.integrate(q=0, sources={'dt', 'dn'})
    def counter(t, x, c=1.):
        return {'dt': 0., 'dn': c}

    paths = 200_000
    x = np.asarray(counter(paths=paths, steps=51, x0=0., c=2., lam=3., seed=1)((0., 1.)))[-1]
    assert np.array_equal(x, np.round(x)) and (x % 2 == 0).all()      # multiples of c
    assert abs(x.mean()/2 - 3.) < 5*np.sqrt(3/paths)
    assert abs(x.var()/4 - 3.) < 8*3*np.sqrt(2/paths)
    # replay: a recorded table of counts drives the same equation
    rng = np.random.default_rng(0)
    dn = rng.poisson(.06, size=(50, 300))
    y = np.asarray(counter(paths=300, steps=51, x0=0., c=2.,
                           dn=m.replay_source(dn))((0., 1.)))[-1]
    assert np.array_equal(y, 2.*dn.sum(axis=0))


def test_user_system_with_jumps_replay_bit_exact():
    """Traced 2-equation system with Poisson ('dn') and Wiener terms and a
    time-dependent coefficient: the reference's own output from its recorded
    increments, bit for bit (golden fixture replay_system_jumps)."""
    from tests.cases import jump_system, k_of_t
    m = sd()
    g = golden('replay_system_jumps')
    cls = m.integrate(q=2, sources={'dt', 'dn', 'dw'})(jump_system)
    P = cls(paths=g['dW'].shape[-1], steps=30, x0=(1., .5), k=k_of_t,
            dw=m.replay_source(g['dW']), dn=m.replay_source(g['dN'].astype(float)))
    x, y = P(g['tt'])
    assert np.array_equal(np.asarray(x), g['out0'])
    assert np.array_equal(np.asarray(y), g['out1'])


def test_reference_test_SDE_cases():
    """The calls of the reference's tests/test_integrator.py::test_SDE
    (:185-404) with its shape assertions, plus moments where they are known:
    systems with time-dependent rho / corr over a stacked axis, Poisson and
    compound Poisson terms, four equations without 'dt', a shared
    true_wiener_source, partially omitted terms."""
    m = sd()
    t = np.linspace(0., 1., 7)

    def f(t, x=0, y=0, z=0, sigma=1., a=2.):
        return ({'dt': a*x, 'dw': sigma*y}, {'dt': -y, 'dw': sigma}, {'dt': z, 'dw': 0.1*x})
    args = dict(x0=(1, 2, 3), sigma=((.1,), (.2,)), a=lambda t: ((1 + t,), (2*t,)),
                vshape=2, paths=11, steps=30)
    for kw in (dict(rho=lambda t: (.01*t, .2, .3)),
               dict(corr=lambda t: ((1, .01*t, .2, .3, .4, .5), (.01*t, 1, -.1, -.2, -.3, -.4),
                                    (.2, -.1, 1, 0, .1, 0), (.3, -.2, 0, 1, .1, .2),
                                    (.4, -.3, .1, .1, 1, 0), (.5, -.4, 0, .2, 0, 1)))):
        for u in m.integrate(f)(**args, **kw)(t):
            assert u.shape == (7, 2, 11) and np.isfinite(u).all()

    # Poisson jumps, alone and with Wiener increments
    def f(t, x=0, y=0, k=1):
        return ({'dt': x, 'dn': k}, {'dt': 1, 'dn': k*y})
    for u in m.integrate(f)(x0=(0., 0.), k=lambda t: .01*t, lam=lambda t: .2*t)(t):
        assert u.shape == (7, 1)

    @m.integrate(q=2, sources={'dt', 'dn'})
    def f_process(t, x, y, k):
        return f(t, x, y, k)
    for u in f_process(x0=(0., 0.), k=lambda t: .01*t, lam=lambda t: .2*t, steps=30)(t):
        assert u.shape == (7, 1)
    # E[x(1)] of dx = x dt + k dn, x0 = 0, constant k and lam: k lam (e - 1)
    x, y = f_process(x0=(0., 0.), k=.5, lam=3., steps=200, paths=200000, seed=5)(t)
    want = .5*3.*(np.e - 1)
    assert abs(x[-1].mean()/want - 1) < .02

    def f(t, x=0, y=0, k=1):
        return ({'dt': x, 'dn': k, 'dw': y}, {'dt': 1, 'dn': k*y, 'dw': x - y})
    for u in m.integrate(f)(x0=(0., 0.), k=lambda t: .01*t, lam=lambda t: .2*t, rho=-.5)(t):
        assert u.shape == (7, 1)

    @m.integrate(q=2, sources={'dt', 'dn', 'dw'})
    def f_process(t, x, y, k):
        return f(t, x, y, k)
    for u in f_process(x0=(0., 0.), k=lambda t: .01*t, lam=lambda t: .2*t,
                       rho=lambda t: -.01*t, steps=30)(t):
        assert u.shape == (7, 1)

    # compound Poisson jumps with every jump law
    def f(t, x=0, y=0, k=1):
        return ({'dt': x, 'dj': k}, {'dt': 1, 'dj': k*y})
    for u in m.integrate(f)(x0=(0., 0.), k=lambda t: .01*t, lam=lambda t: .2*t,
                            y=m.uniform_rv(a=-1, b=lambda t: t))(t):
        assert u.shape == (7, 1)

    @m.integrate(q=2, sources={'dt', 'dj'})
    def f_process(t, x, y, k):
        return f(t, x, y, k)
    for u in f_process(x0=(0., 0.), k=lambda t: .01*t, lam=lambda t: .2*t,
                       y=m.exp_rv(a=-1), steps=30)(t):
        assert u.shape == (7, 1)

    def f(t, x=0, y=0, k=1):
        return ({'dt': x, 'dj': k, 'dw': y}, {'dt': 1, 'dj': k*y, 'dw': x - y})
    for u in m.integrate(f)(x0=(0., 0.), k=lambda t: .01*t, lam=lambda t: .2*t, rho=-.5,
                            y=m.norm_rv(a=1, b=lambda t: 2 + t))(t):
        assert u.shape == (7, 1)

    @m.integrate(q=2, sources={'dt', 'dj', 'dw'})
    def f_process(t, x, y, k):
        return f(t, x, y, k)
    for u in f_process(x0=(0., 0.), k=lambda t: .01*t, lam=lambda t: .2*t,
                       rho=lambda t: -.01*t, y=m.double_exp_rv(a=1, b=2, pa=.1), steps=30)(t):
        assert u.shape == (7, 1)

    # four equations with no 'dt' term, correlated; then a shared true_wiener_source
    def f(t, x=0, y=0, z=0, w=0):
        return ({'dw': x}, {'dw': y}, {'dw': z}, {'dw': w})
    rm4 = np.random.default_rng(1).random((4, 4))
    corr4 = np.eye(4) + 0.1*(rm4 + rm4.T)
    for u in m.integrate(f)(x0=(1,)*4, corr=corr4, paths=11, steps=30)(t):
        assert u.shape == (7, 11)
    tw = m.true_wiener_source(vshape=4, paths=11, corr=corr4)
    xs1 = m.integrate(f)(x0=(1.,)*4, dw=tw, paths=11, steps=30)(t)
    xs2 = m.integrate(f)(x0=(1.,)*4, dw=tw, paths=11, steps=30)(t)
    for u, v in zip(xs1, xs2):
        assert u.shape == (7, 11)
        assert np.allclose(np.asarray(u), np.asarray(v), rtol=16*np.finfo(float).eps)

    @m.integrate
    def f_process(t, x=0, y=0, z=0, w=0):
        return ({'dt': 1, 'dw': 1}, {'dt': 1}, {'dw': 1}, {})
    xs = f_process(x0=(1,)*4, paths=11, steps=30)(t)
    assert np.allclose(np.asarray(xs[1])[-1], 2.) and np.array_equal(np.asarray(xs[3])[-1], np.ones(11))


def test_milstein_replay_matches_the_reference_machinery_fixture():
    """tests/golden/replay_milstein_plugin.npz = the unmodified reference
    integrator running a Milstein scheme plugged in through its ``method=``
    hook (make_milstein_plugin.py; the oracle reproduces it bit for bit,
    tests/test_oracle_golden.py).  Kernel within the north-star 1e-12."""
    from tests.cases import golden
    m = sd()
    g = golden('replay_milstein_plugin')

    def gbm(t, x, mu=.05, sigma=.4):
        return {'dt': mu*x, 'dw': sigma*x}

    def cev(t, x, mu=0., sigma=1.):
        return {'dt': mu*x, 'dw': sigma*x**1.5}

    cases = (('gbm', gbm, dict(mu=.05, sigma=.4), 1.),
             ('cev', cev, dict(mu=lambda t: .02 + .03*t, sigma=lambda t: .3 - .1*t), .8))
    for name, f, par, x0 in cases:
        grid, tt, dW, want = (g[name + '_' + k] for k in ('grid', 'tt', 'dW', 'x'))
        P = m.integrate(f)(paths=dW.shape[-1], steps=grid, x0=x0, method='milstein',
                           dw=m.replay_source(dW), **par)
        x = np.asarray(P(tt))
        assert x.shape == want.shape
        assert np.abs(x/want - 1).max() < 1e-12, name


def test_explicit_time_in_traced_sde_replay_bit_exact():
    """A traced function that uses `t` arithmetically with the state and as a
    coefficient of its own: every step must see its own time (the value used
    to be frozen at the first trace)."""
    m = sd()

    def f(t, x, a=2.):
        return {'dt': a*(t - x), 'dw': 1 + t}

    n, paths = 40, 512
    grid = np.cumsum(np.concatenate(([0.], np.random.default_rng(5).uniform(.01, .09, n))))
    rng = np.random.default_rng(3)
    dW = rng.standard_normal((n, paths))*np.sqrt(np.diff(grid))[:, None]
    where = [0, n//2, n]
    P = m.integrate(f)(paths=paths, steps=grid, x0=.5, a=2., dw=m.replay_source(dW))
    x = np.asarray(P(grid[where]))
    o = orc.generic_replay(f, dict(a=2.), .5, grid, where, dW)
    assert np.array_equal(x, o)
    # the frozen-t result would be a plain OU pulled to 0 with unit diffusion
    frozen = orc.generic_replay(lambda t, x, a: {'dt': a*(0. - x), 'dw': 1.}, dict(a=2.),
                                .5, grid, where, dW)
    assert np.abs(x[-1] - frozen[-1]).max() > .1


def test_user_let_and_info_hooks_bit_exact_vs_reference():
    """User classes with `let`, `info_begin`, `info_next`, `info_end` overrides
    (tests/cases.py:user_hooks_*): the reference ran them on its own recorded
    increments (tests/golden/make_user_hooks.py); the kernel replays the same
    increments with the hooks compiled in -- stored values and counters bit for
    bit (the let bodies only multiply and add)."""
    from tests.cases import golden, user_hooks_single, user_hooks_system
    import sdepy_b200 as m
    g = golden('replay_user_hooks')
    A = user_hooks_single(m)
    kw = dict(paths=37, vshape=(2,), steps=23, x0=.3, k=g['a_k'])
    P = A(dw=m.replay_source(g['a_dW']), **kw)
    x = P(g['a_tt'])
    assert isinstance(x, m.process) and x.shape == g['a_out'].shape
    assert np.array_equal(np.asarray(x), g['a_out'])
    assert np.array_equal(P.info['neg'], g['a_neg'])
    assert np.array_equal(P.info['big'], g['a_big'])
    assert P.info['total_neg'] == int(g['a_total_neg'])
    # Philox mode, device and stats outputs go through the same compiled hooks
    Q = A(seed=3, output='device', **kw)
    xd = Q(g['a_tt'])
    S = A(seed=3, output='stats', **kw)
    st = S(g['a_tt'])
    assert np.allclose(np.asarray(st.pmean())[..., 0], np.asarray(xd.pmean())[..., 0], rtol=1e-12)
    assert (xd.x >= 1.).all()                       # x*x + 1
    assert np.array_equal(S.info['neg'], Q.info['neg']) and Q.info['neg'].sum() > 0

    B = user_hooks_system(m)
    kw = dict(paths=29, vshape=(3,), steps=19, x0=(1., .8))
    P = B(dw=m.replay_source(g['b_dW']), **kw)
    x = P(g['b_tt'])
    assert isinstance(x, m.process) and x.shape == g['b_out'].shape == (5, 3, 29)
    assert np.array_equal(np.asarray(x), g['b_out'])
    assert np.array_equal(P.info['ylow'], g['b_ylow'])
    # getinfo=False: no counters, same paths
    P = B(dw=m.replay_source(g['b_dW']), getinfo=False, **kw)
    assert np.array_equal(np.asarray(P(g['b_tt'])), g['b_out']) and 'ylow' not in P.info


def test_sde_with_both_dn_and_dj_terms():
    """An equation with a Poisson ('dn') AND a compound Poisson ('dj') term (two
    jump slots in the kernel).  Replay of the reference's own recorded increments
    (tests/golden/make_two_jump_terms.py) bit for bit; in Philox mode the two
    sources are independent streams and the mean follows
    m' = (c lam - a) m + lam E[y]."""
    from tests.cases import golden, two_jump_terms
    import sdepy_b200 as m
    g = golden('replay_two_jump_terms')
    cls = m.integrate(two_jump_terms)
    assert set(cls.sources) == {'dt', 'dw', 'dn', 'dj'}
    kw = dict(paths=31, vshape=(2,), steps=25, x0=1., a=g['p_a'])
    P = cls(dw=m.replay_source(g['dW']), dn=m.replay_source(g['dN'].astype(float)),
            dj=m.replay_source(g['dJ']), **kw)
    x = P(g['tt'])
    assert np.array_equal(np.asarray(x), g['out'])
    # Philox mode
    a, c, lam, ya = .5, .2, 3., -.1
    paths = 400_000
    Q = cls(paths=paths, steps=401, x0=1., a=a, c=c, lam=lam, y=m.norm_rv(a=ya, b=.2), seed=21,
            output='stats')
    st = Q((0., 1.))
    r = c*lam - a
    want = np.exp(r) + lam*ya*(np.exp(r) - 1)/r
    got, err = float(np.asarray(st.pmean())[-1, 0]), float(np.asarray(st.stderr())[-1, 0])
    assert abs(got - want) < 4*err + 2e-3, (got, want, err)
    # the two jump sources are distinct streams: dumped increments
    D = cls(paths=20_000, steps=51, x0=1., lam=lam, y=m.norm_rv(a=ya, b=.2), seed=22)
    D._dump_increments = True
    D((0., 1.))
    d = D._last_run.dump[0]
    dj, dn = d['dJ'].cpu().numpy().reshape(50, 2, -1)[:, 0], d['dJ'].cpu().numpy().reshape(50, 2, -1)[:, 1]
    cn = d['dN'].cpu().numpy().reshape(50, 2, -1)
    assert np.array_equal(dn, cn[:, 1].astype(float))            # 'dn' slot: unit jumps
    assert abs(cn[:, 0].mean() - lam/50) < 5*np.sqrt(lam/50/cn[:, 0].size)
    assert abs(cn[:, 1].mean() - lam/50) < 5*np.sqrt(lam/50/cn[:, 1].size)
    assert not np.array_equal(cn[:, 0], cn[:, 1])
    assert abs(dj.sum()/max(cn[:, 0].sum(), 1) - ya) < 5*.2/np.sqrt(max(cn[:, 0].sum(), 1))
