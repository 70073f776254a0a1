"""The full-path STREAM kernel (csrc/sde_engine.cuh:stream_body) against the
general kernel: same seed / same replayed increments -> bit-identical paths and
diagnostics, whichever kernel the launcher picks (SDEB_NO_STREAM=1 forces the
general one).  Parity of both with the oracle is tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KERNEL_GENERAL, KERNEL_LEAN, KERNEL_STREAM = 0, 1, 2


def sd():
    import sdepy_b200
    return sdepy_b200


def both(make, timeline, expect_stream=True):
    """Run make() twice -- stream kernel allowed / forbidden -- and return the
    two device results after checking which kernels ran."""
    outs = []
    for no_stream in ('0', '1'):
        os.environ['SDEB_NO_STREAM'] = no_stream
        try:
            P = make()
            P._trace_kernels = True
            x = P(timeline)
            kern = list(P._last_run.kernels)
        finally:
            os.environ.pop('SDEB_NO_STREAM', None)
        if no_stream == '1':
            assert KERNEL_STREAM not in kern
        elif expect_stream:
            assert kern and all(k == KERNEL_STREAM for k in kern), kern
        outs.append((x, P))
    return outs


def same(a, b):
    a, b = (z if isinstance(z, tuple) else (z,) for z in (a, b))
    assert len(a) == len(b)
    for u, v in zip(a, b):
        assert u.x.shape == v.x.shape
        assert torch.equal(u.x, v.x)


def hw_theta(t):
    return np.array(((.02 + .001*t,), (0.,), (0.,)))


def hw_corr(t):
    c01, c02, c12 = .3*np.cos(t), -.2 + .05*t, .1
    return np.array(((1, c01, c02), (c01, 1, c12), (c02, c12, 1)))


@pytest.mark.parametrize('paths', [2, 510, 5000])
def test_stream_equals_general_philox(paths):
    m = sd()
    tl = np.linspace(0., 2., 71)          # 70 steps: a whole step block + a tail, odd periods
    cases = [
        lambda: m.ornstein_uhlenbeck_process(x0=.1, theta=lambda s: .2 + .1*s, k=1., sigma=.3,
                                             paths=paths, seed=2, output='device'),
        lambda: m.hull_white_process(factors=3, x0=((.01,), (0.,), (0.,)), theta=hw_theta,
                                     k=((.1,), (.5,), (1.,)), sigma=((.01,), (.008,), (.005,)),
                                     corr=hw_corr, paths=paths, seed=3, output='device'),
        lambda: m.full_heston_process(x0=1., y0=.04, mu=lambda t: .01*t, sigma=1., theta=.04, k=2.,
                                      xi=.6, rho=-.7, paths=paths, seed=4, output='device'),
        lambda: m.cox_ingersoll_ross_process(x0=.04, theta=lambda t: .04 + .01*t, k=2., xi=.5,
                                             paths=paths, vshape=(3,), seed=5, output='device'),
        # time-invariant records with several lane groups (the lean kernel needs one group)
        lambda: m.lognorm_process(x0=1., mu=((.01,), (.02,)), sigma=.2, vshape=(2,), paths=paths,
                                  seed=6, output='device'),
        lambda: m.wiener_process(x0=0., vshape=(2,), corr=((1, .5), (.5, 1)), paths=paths, seed=7,
                                 output='device', steps=141),
    ]
    for k, make in enumerate(cases):
        # the last case is one lane group with a time-invariant record: that is
        # the lean kernel's configuration (also bit-identical to the general one)
        (xs, Ps), (xg, Pg) = both(make, tl, expect_stream=k < len(cases) - 1)
        same(xs, xg)
        if 'negative_y_count' in Ps.info:
            assert torch.equal(Ps.info['negative_y_count'], Pg.info['negative_y_count'])


def test_stream_equals_general_replay_and_oracle():
    from oracle import sde_oracle as orc
    m = sd()
    paths, n = 1000, 90
    grid = np.linspace(0., 1., n + 1)
    rng = np.random.default_rng(3)
    dW = rng.standard_normal((n, paths))*np.sqrt(np.diff(grid))[:, None]
    par = dict(theta=.2, k=1.5, sigma=.3)
    (xs, _), (xg, _) = both(lambda: m.ornstein_uhlenbeck_process(
        x0=.1, paths=paths, dw=m.replay_source(dW), output='device', **par), grid)
    same(xs, xg)
    want, _ = orc.euler_replay('ornstein_uhlenbeck', par, .1, grid, range(n + 1), dW)
    assert np.array_equal(xs.x.cpu().numpy(), want)
    # Heston, both components, every 3rd step stored, time-dependent parameter
    dW2 = rng.standard_normal((n, 2, paths))*np.sqrt(np.diff(grid))[:, None, None]
    hp = dict(mu=lambda t: .03 + .01*t, sigma=1., theta=.04, k=2., xi=.6)
    where = list(range(0, n + 1, 3))
    (xs, Ps), (xg, Pg) = both(lambda: m.full_heston_process(
        x0=100., y0=.04, paths=paths, steps=grid, dw=m.replay_source(dW2), output='device',
        **hp), grid[where])
    same(xs, xg)
    (ox, oy), oinfo = orc.euler_replay('heston', hp, 100., grid, where, dW2, y0=.04, full=True)
    assert np.array_equal(xs[1].x.cpu().numpy(), oy)
    assert np.abs(xs[0].x.cpu().numpy()/ox - 1).max() <= 4*np.finfo(float).eps
    assert np.array_equal(Ps.info['negative_y_count'].cpu().numpy(), oinfo['negative_y_count'])


def test_stream_backward_sweep_and_per_path_x0():
    m = sd()
    paths = 3000
    tl = np.linspace(0., 1., 31)
    x0 = np.linspace(.5, 1.5, paths)
    for i0 in (0, 11, -1):
        (xs, _), (xg, _) = both(lambda: m.ornstein_uhlenbeck_process(
            x0=x0, theta=lambda s: .2*s, k=1., sigma=.3, paths=paths, seed=9, i0=i0,
            steps=91, output='device'), tl)
        same(xs, xg)
        assert torch.equal(xs.x[i0].cpu(), torch.from_numpy(x0))


def test_stream_dense_and_sparse_step_blocks():
    """Step blocks (64 steps) whose steps all store into consecutive rows take the
    advancing-cursor loop, the others look each row up: both in one run, either order."""
    m = sd()
    paths = 1030
    dense, sparse = np.linspace(0., 1., 65), np.linspace(1., 2., 65)
    for first_dense in (True, False):
        if first_dense:
            grid = np.concatenate((dense, sparse[1:]))
            tl = np.concatenate((dense, sparse[[20, 64]]))
        else:
            grid = np.concatenate((dense, sparse[1:]))
            tl = np.concatenate((dense[[0, 7, 64]], sparse[1:]))
        (xs, _), (xg, _) = both(lambda: m.ornstein_uhlenbeck_process(
            x0=.1, theta=lambda s: .2 + .1*s, k=1., sigma=.3, paths=paths, seed=12,
            steps=grid, output='device'), tl)
        same(xs, xg)
        assert xs.x.shape[0] == tl.size and bool(torch.isfinite(xs.x).all())


def test_records_too_long_for_the_alignment_gap():
    """Five correlated Heston pairs with time-dependent theta: 85 doubles per staged
    record, 64 x 86 x 8 B = 44 KB per step block -- more than the 28 KB of the gap
    in front of the generator tables, so both kernels place the records BEHIND
    the tables (csrc/sde_engine.cuh: `front` / FRONT false; sdeb.cu adds the bytes)."""
    m = sd()
    rng = np.random.default_rng(5)
    c = np.eye(10) + .1*rng.random((10, 10))
    c = (c + c.T)/2
    tl = np.linspace(0., 1., 81)                 # two step blocks
    (xs, Ps), (xg, Pg) = both(lambda: m.full_heston_process(
        vshape=(5,), x0=1., y0=.04, mu=.02, sigma=1., theta=lambda t: .04 + .01*t, k=2., xi=.4,
        corr=c, paths=700, seed=8, output='device'), tl)
    same(xs, xg)
    assert torch.equal(Ps.info['negative_y_count'], Pg.info['negative_y_count'])
    x, y = (np.asarray(z.x.cpu()) for z in xs)
    assert np.isfinite(x).all() and (x > 0).all()
    assert abs(y[-1].mean() - .0457) < .004      # y relaxes towards theta(t) from .04
    zc = np.corrcoef(np.log(x[-1, 0]/x[-2, 0]), np.log(x[-1, 1]/x[-2, 1]))[0, 1]
    assert abs(zc - c[0, 1]) < .12               # last-step log-returns carry corr


def test_stream_traced_sde():
    m = sd()

    @m.integrate
    def f(t, x, a=.3, b=.2):
        return {'dt': -a*x*t, 'dw': b*np.sqrt(1 + x*x)}

    tl = np.linspace(0., 1., 41)
    for method in ('euler', 'milstein'):
        (xs, _), (xg, _) = both(lambda: f(paths=4000, x0=1., seed=5, method=method,
                                          output='device'), tl)
        same(xs, xg)


def test_jump_models_and_odd_pitch_keep_the_general_kernel():
    m = sd()
    tl = np.linspace(0., 1., 11)
    for make in (lambda: m.merton_jumpdiff_process(paths=1000, lam=2., seed=1, output='device'),
                 lambda: m.ornstein_uhlenbeck_process(paths=1001, theta=lambda t: t, seed=1,
                                                      output='device')):
        P = make()
        P._trace_kernels = True
        P(tl)
        assert KERNEL_STREAM not in P._last_run.kernels


@pytest.mark.parametrize('kind', ['merton', 'kou'])
def test_stream_replays_jump_models(kind):
    """Replayed jump-diffusions take the stream kernel too (dJ and the dN counts
    ride the cp.async ring next to dW): same paths, jump counts and jump rates
    as the general kernel, and the oracle's paths within 4 ulp."""
    from oracle import sde_oracle as orc
    m = sd()
    paths, n = 2000, 150
    grid = np.linspace(0., 1., n + 1)
    par = dict(mu=.05, sigma=.2)
    if kind == 'merton':
        law, cls, lkw = orc.jump_law('norm', a=-.1, b=.15), m.merton_jumpdiff_process, dict(a=-.1, b=.15)
    else:
        law, cls = orc.jump_law('double_exp', a=.1, b=.15, pa=.4), m.kou_jumpdiff_process
        lkw = dict(a=.1, b=.15, pa=.4)
    rng = np.random.default_rng(17)
    dW = np.empty((n, paths)); dJ = np.empty((n, paths)); dN = np.empty((n, paths), dtype=np.int64)
    for i in range(n):
        s, ds = grid[i], grid[i + 1] - grid[i]
        dJ[i], dN[i] = orc.draw_cpoisson(rng, s, ds, (), paths, 4., law)
        dW[i] = orc.draw_wiener(rng, s, ds, (), paths)
    where = list(range(0, n + 1, 10))
    (xs, Ps), (xg, Pg) = both(lambda: cls(
        paths=paths, steps=grid, x0=1., lam=4., dw=m.replay_source(dW),
        dj=m.replay_source(dJ, dn=dN), output='device', **par, **lkw), grid[where])
    same(xs, xg)
    assert torch.equal(Ps.info['jump_count'], Pg.info['jump_count'])
    assert np.array_equal(Ps.info['jump_rate'], Pg.info['jump_rate'])
    want, winfo = orc.euler_replay('jumpdiff', par, 1., grid, where, dW, dJ=dJ, dN=dN)
    assert np.abs(xs.x.cpu().numpy()/want - 1).max() <= 4*np.finfo(float).eps
    assert np.array_equal(Ps.info['jump_count'].cpu().numpy(), winfo['jump_count'])
