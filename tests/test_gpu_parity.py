"""GPU parity tests (-m gpu): the CUDA path, called through the sdepy-style
classes -> C ABI, against (a) the golden outputs of the unmodified reference
and (b) the CPU oracle on seeded inputs.

Tolerances
* replay mode, non-exponentiated components: BIT-EXACT (np.array_equal);
* replay mode, components that go through a final exp(): <= 4 ulp relative
  (CUDA exp vs the host libm exp; north-star bound 1e-12);
* integer diagnostics (negative_y_count, jump_count): bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import sde_oracle as orc
from tests.cases import golden, REPLAY, HW, hw_corr, hw_theta, theta_t, HESTON

pytestmark = pytest.mark.gpu

ULP4 = 4*np.finfo(float).eps


def sd():
    import sdepy_b200
    return sdepy_b200


def classes():
    m = sd()
    return {
        'replay_wiener': (m.wiener_process, dict(x0=.5, mu=.1, sigma=.7), {}),
        'replay_lognorm': (m.lognorm_process, dict(x0=1., mu=.05, sigma=.2), {}),
        'replay_lognorm_v3': (m.lognorm_process, dict(
            x0=((1.,), (2.,), (3.,)), mu=((.05,), (.0,), (-.1,)),
            sigma=((.2,), (.3,), (.1,))), dict(vshape=(3,))),
        'replay_oruh_tdep': (m.ornstein_uhlenbeck_process,
                             dict(x0=.1, theta=theta_t, k=1., sigma=.3), {}),
        'replay_hw3_tdep': (m.hull_white_process, dict(
            factors=3, x0=HW['x0'], theta=hw_theta, k=HW['k'],
            sigma=HW['sigma']), {}),
        'replay_cir': (m.cox_ingersoll_ross_process,
                       dict(x0=.05, theta=.04, k=1.5, xi=.6), {}),
        'replay_heston': (m.heston_process, dict(
            x0=100., mu=.03, sigma=1., y0=.04, theta=.04, k=2., xi=.9), {}),
        'replay_heston_full': (m.full_heston_process, dict(
            x0=100., mu=.03, sigma=1., y0=.04, theta=.04, k=2., xi=.3), {}),
        'replay_heston_v2': (m.full_heston_process, dict(
            x0=((100.,), (50.,)), mu=.03, sigma=1., y0=.04, theta=.04, k=2.,
            xi=((.3,), (1.1,))), dict(vshape=(2,))),
        'replay_merton': (m.merton_jumpdiff_process,
                          dict(x0=1., mu=.05, sigma=.2), {}),
        'replay_kou': (m.kou_jumpdiff_process,
                       dict(x0=1., mu=.05, sigma=.2), {}),
        'replay_oruh_ragged': (m.ornstein_uhlenbeck_process,
                               dict(x0=1., theta=.5, k=2., sigma=.4), {}),
    }


EXACT = {'replay_wiener', 'replay_oruh_tdep', 'replay_hw3_tdep', 'replay_cir',
         'replay_oruh_ragged'}


@pytest.mark.parametrize('name', sorted(REPLAY))
@pytest.mark.parametrize('table_on', ['host', 'device'])
def test_replay_matches_reference(name, table_on):
    m = sd()
    g = golden(name)
    cls, params, extra = classes()[name]
    dW = g['dW']
    paths = dW.shape[-1]

    def tab(a):
        return torch.from_numpy(np.ascontiguousarray(a)).cuda() if table_on == 'device' else a
    kw = dict(dw=m.replay_source(tab(dW)))
    if 'dJ' in g:
        kw['dj'] = m.replay_source(tab(g['dJ']), dn=tab(g['dN']))
    # the explicit grid points reproduce the reference's step grid exactly
    P = cls(paths=paths, steps=g['grid'], **params, **extra, **kw)
    out = P(g['tt'])
    outs = out if isinstance(out, tuple) else (out,)
    for i, o in enumerate(outs):
        ref = g['out%d' % i]
        assert isinstance(o, m.process) and o.shape == ref.shape
        assert np.array_equal(o.t, g['tt'])
        exact = name in EXACT or (name.startswith('replay_heston') and i == 1)
        if exact:
            assert np.array_equal(np.asarray(o), ref), name
        else:
            err = np.abs(np.asarray(o)/ref - 1).max()
            assert err <= ULP4, (name, err)
    assert P.info['computed_steps'] == int(g['computed_steps'])
    assert P.info['stored_steps'] == int(g['stored_steps'])
    for key in ('negative_y_count', 'jump_count'):
        if key in g:
            assert np.array_equal(P.info[key], g[key]), key
    if 'jump_rate' in g:
        assert np.allclose(P.info['jump_rate'], g['jump_rate'], rtol=1e-14, atol=0)


def test_replay_generic_source_protocol():
    """Any object obeying the reference's source protocol can drive the
    kernel: here a host `process` holding a Brownian path (its __call__
    returns increments, reference infrastructure.py:615-633)."""
    m = sd()
    g = golden('known_lognorm_exact')
    t, w = g['t'], g['w']
    wp = m.process(t=t, x=w)
    x = m.lognorm_process(paths=w.shape[-1], x0=1., mu=.05, sigma=.2, dw=wp)(t)
    exact = np.exp((.05 - .2*.2/2)*t[:, None] + .2*w)
    assert np.allclose(np.asarray(x), exact, rtol=16*np.finfo(float).resolution)
    assert np.allclose(np.asarray(x), g['x'], rtol=1e-13)


def test_seeded_oracle_large_heston():
    """CUDA vs oracle on fresh seeded increments at a size the oracle
    finishes in seconds (2e5 paths x 252 steps)."""
    m = sd()
    rng = np.random.default_rng(5)
    paths, n = 200_000, 252
    grid = np.linspace(0., 1., n + 1)
    dW = rng.standard_normal((n, 2, paths))*np.sqrt(np.diff(grid))[:, None, None]
    par = dict(mu=.03, sigma=1., theta=.04, k=2., xi=.3)
    (ox, oy), oinfo = orc.euler_replay('heston', par, 100., grid, [0, n], dW,
                                       y0=.04, full=True)
    P = m.full_heston_process(paths=paths, steps=grid, x0=100., y0=.04,
                              dw=m.replay_source(dW), **par)
    x, y = P((0., 1.))
    assert np.array_equal(np.asarray(y), oy)
    assert np.abs(np.asarray(x)/ox - 1).max() <= ULP4
    assert np.array_equal(P.info['negative_y_count'], oinfo['negative_y_count'])


def test_backward_and_forward_from_inner_point():
    m = sd()
    rng = np.random.default_rng(9)
    tt = np.array([0., .25, 1.])
    grid = np.linspace(0, 1, 9)
    paths = 77
    # i0 = 1: backward sweep .25 -> 0 (2 steps), forward .25 -> 1 (6 steps)
    dWb = rng.standard_normal((2, paths))*np.sqrt(.125)
    dWf = rng.standard_normal((6, paths))*np.sqrt(.125)
    par = dict(theta=.5, k=2., sigma=.4)
    P = m.ornstein_uhlenbeck_process(paths=paths, steps=grid, i0=1, x0=1.,
                                     dw=m.replay_source(np.concatenate((dWb, dWf))), **par)
    x = np.asarray(P(tt))
    xb, _ = orc.euler_replay('ornstein_uhlenbeck', par, 1., grid[2::-1], [0, 2], dWb)
    xf, _ = orc.euler_replay('ornstein_uhlenbeck', par, 1., grid[2:], [0, 6], dWf)
    assert np.array_equal(x[1], xb[0]) and np.array_equal(x[0], xb[1])
    assert np.array_equal(x[2], xf[1])


def test_error_conventions():
    m = sd()
    P = m.lognorm_process(paths=10, steps=5)
    with pytest.raises(ValueError):
        P(((0., 1.), (2., 3.)))
    with pytest.raises(ValueError):
        P((0., 2., 1.))
    with pytest.raises(IndexError):
        m.lognorm_process(paths=10, i0=5)((0., 1.))
    with pytest.raises(TypeError):
        m.lognorm_process(paths=10, nonexistent=1.)
    with pytest.raises(ValueError):
        m.lognorm_process(paths=10, method='nonexistent')
    with pytest.raises(ValueError):
        m.lognorm_process(paths=10, dw=m.replay_source(np.zeros((3, 11))))
    with pytest.raises(TypeError):
        m.wiener_source(paths=3, rng=1234)
