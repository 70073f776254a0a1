"""BASELINE.json configurations at FULL size on the GPU, checked through
size-independent properties (the oracle cannot run these sizes in seconds):
additivity of shard statistics, discrete-Euler moment recursions, closed forms
within standard errors, dump -> replay idempotence, histogram conservation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HESTON = dict(x0=100., mu=.03, sigma=1., y0=.04, theta=.04, k=2., xi=.3, rho=-.7)


def sd():
    import sdepy_b200
    return sdepy_b200


def test_config3_heston_1e8_terminal_stats_and_shard_additivity():
    m = sd()
    from sdepy_b200 import _lib
    paths = 100_000_000
    grid = np.linspace(0., 1., 253)
    pay = ('call', 100., float(np.exp(-.03)))
    kw = dict(steps=grid, seed=2024, output='stats', payoff=pay, getinfo=False, **HESTON)
    full = m.heston_process(paths=paths, **kw)((0., 1.))
    price = float(np.asarray(full.payoff_mean())[-1, 0])
    err = float(np.asarray(full.payoff_stderr())[-1, 0])
    # closed form 9.2425 (Gil-Pelaez on sdepy.analytical.heston_log_chf); the
    # Euler/full-truncation bias at dt = 1/252 is ~3e-3 (reference at 1e6
    # paths: 9.2335 +/- 0.0116)
    assert err < 2e-3 and abs(price - 9.2425) < 4*err + 1e-2
    mean = float(np.asarray(full.pmean())[-1, 0])
    assert abs(mean - 100*np.exp(.03)) < 5*float(np.asarray(full.stderr())[-1, 0])
    # the same global paths integrated as 3 uneven shards: sums add up
    parts, off = [], 0
    for n in (40_000_000, 35_000_001, 24_999_999):
        parts.append(m.heston_process(paths=n, path_offset=off, **kw)((0., 1.)))
        off += n
    s = sum(p.sums for p in parts)
    for k in (0, 1, 2, 3, 6, 7):
        assert np.allclose(s[..., k], full.sums[..., k], rtol=1e-10, atol=1e-6)
    assert np.array_equal(np.min([p.sums[..., 4] for p in parts], axis=0), full.sums[..., 4])
    assert np.array_equal(np.max([p.sums[..., 5] for p in parts], axis=0), full.sums[..., 5])


def test_config2_ou_hw_1e6_full_path():
    m = sd()
    paths, n = 1_000_000, 500
    tl = np.linspace(0., 5., n + 1)
    x = m.ornstein_uhlenbeck_process(x0=.1, theta=lambda t: .2 + .1*t, k=1., sigma=.3,
                                     paths=paths, seed=9, output='device')(tl)
    assert x.shape == (n + 1, paths)
    mean = np.asarray(x.pmean())[:, 0]
    var = np.asarray(x.pvar())[:, 0]
    mm, vv, want_m, want_v = .1, 0., [.1], [0.]
    for i in range(n):          # exact moments of the discrete Euler recursion
        dt = tl[i + 1] - tl[i]
        mm, vv = mm + 1.*((.2 + .1*tl[i]) - mm)*dt, (1 - dt)**2*vv + .09*dt
        want_m.append(mm); want_v.append(vv)
    want_m, want_v = np.array(want_m), np.array(want_v)
    assert np.abs(mean[1:] - want_m[1:]).max() < 5*np.sqrt(want_v.max()/paths)
    assert np.abs(var[1:]/want_v[1:] - 1).max() < 6*np.sqrt(2/paths)
    assert torch.isfinite(x.x).all()
    from tests.cases import HW
    h = m.hull_white_process(paths=paths, seed=10, output='device', **HW)(tl)
    assert h.shape == (n + 1, paths) and torch.isfinite(h.x).all()
    hm = np.asarray(h.pmean())[:, 0]
    # factor means follow m <- m + k(theta - m)dt; only factor 0 has a non-zero mean
    f0, want = .01, [.01]
    for i in range(n):
        f0 = f0 + .1*((.02 + .001*tl[i]) - f0)*(tl[i + 1] - tl[i])
        want.append(f0)
    hv = float(np.asarray(h.pvar())[-1, 0])
    assert np.abs(hm - np.array(want)).max() < 5*np.sqrt(hv/paths)


def test_config4_jumps_replay_idempotence_chunk():
    """One 1e6-path chunk of the 1e7 x 1000 Merton config: dump the Philox
    increments, replay them, compare paths and jump counts bit for bit."""
    m = sd()
    paths, n = 1_000_000, 1000
    kw = dict(x0=1., mu=.05, sigma=.2)
    P = m.merton_jumpdiff_process(paths=paths, steps=n + 1, lam=2., a=-.1, b=.15, seed=11,
                                  output='device', **kw)
    P._dump_increments = True
    x = P((0., 1.))
    d = P._last_run.dump[0]
    jc = P.info['jump_count']
    assert abs(float(jc.double().mean()) - 2.) < 5*np.sqrt(2/paths)
    R = m.merton_jumpdiff_process(
        paths=paths, steps=n + 1, output='device',
        dw=m.replay_source(d['dW'].reshape(n, paths)),
        dj=m.replay_source(d['dJ'].reshape(n, paths), dn=d['dN'].reshape(n, paths)), **kw)
    xr = R((0., 1.))
    assert torch.equal(xr.x, x.x)
    assert torch.equal(R.info['jump_count'], jc)
    assert torch.equal(d['dN'].sum(dim=0).reshape(-1), jc.reshape(-1))
    assert np.allclose(R.info['jump_rate'], P.info['jump_rate'], rtol=1e-13)


@pytest.mark.parametrize('kind', ['merton', 'kou'])
def test_config4_chunks_against_the_oracle(kind):
    """Config 4 at its own step count against the ORACLE (not against the
    kernel's own draws): chunks of 1e5 paths x 1000 steps whose increments are
    drawn by the oracle's restatement of the reference sources
    (wiener_source / cpoisson_source, infrastructure.py:1503-1560, 2017-2040),
    integrated by the kernel in replay mode and by orc.euler_replay('jumpdiff'):
    paths within 4 ulp (the final exp), jump counts bit-equal.  Chunks are
    looped until 1e7 paths or a time budget is spent."""
    import time
    from oracle import sde_oracle as orc
    m = sd()
    chunk, n, budget = 100_000, 1000, 25.
    grid = np.linspace(0., 1., n + 1)
    par = dict(mu=.05, sigma=.2)
    if kind == 'merton':
        law, cls, lkw = orc.jump_law('norm', a=-.1, b=.15), m.merton_jumpdiff_process, dict(a=-.1, b=.15)
    else:
        law, cls = orc.jump_law('double_exp', a=.1, b=.15, pa=.4), m.kou_jumpdiff_process
        lkw = dict(a=.1, b=.15, pa=.4)
    rng = np.random.default_rng(404)
    done, t0, eps = 0, time.perf_counter(), np.finfo(float).eps
    while done < 10_000_000 and (done == 0 or time.perf_counter() - t0 < budget):
        dW = np.empty((n, chunk)); dJ = np.empty((n, chunk)); dN = np.empty((n, chunk), dtype=np.int64)
        for i in range(n):      # sorted-id order of the reference: dj before dw
            s, ds = grid[i], grid[i + 1] - grid[i]
            dJ[i], dN[i] = orc.draw_cpoisson(rng, s, ds, (), chunk, 2., law)
            dW[i] = orc.draw_wiener(rng, s, ds, (), chunk)
        want, winfo = orc.euler_replay('jumpdiff', par, 1., grid, [0, n], dW, dJ=dJ, dN=dN)
        P = cls(paths=chunk, steps=grid, x0=1., lam=2., dw=m.replay_source(dW),
                dj=m.replay_source(dJ, dn=dN), **par, **lkw)
        got = np.asarray(P((0., 1.)))
        assert got.shape == want.shape == (2, chunk)
        assert np.abs(got/want - 1).max() <= 4*eps
        assert np.array_equal(P.info['jump_count'], winfo['jump_count'])
        done += chunk
    assert done >= chunk


def test_poisson_counts_at_large_intensity():
    """lam*|dt| far beyond the inversion range (exp(-lam|dt|) underflows past
    ~745): counts are drawn as a sum of independent Poisson(lam|dt|/m) variates
    -- mean and variance of Poisson(L) within 5 standard errors, for L on both
    sides of the switch at 30 and far beyond it."""
    m = sd()
    paths = 400_000
    for L in (29., 31., 500., 2000.):
        dn = np.asarray(m.poisson_source(paths=paths, lam=L, seed=int(L))(0., 1.)).astype(float)
        assert abs(dn.mean() - L) < 5*np.sqrt(L/paths)
        assert abs(dn.var() - L) < 5*L*np.sqrt(2/paths) + 5*np.sqrt(L/paths)
        third = ((dn - L)**3).mean()                  # third central moment = L
        assert abs(third - L) < 6*np.sqrt(15*L**3/paths)


def test_config5_milstein_1e8_montecarlo():
    m = sd()

    @m.integrate
    def f(t, x, mu=.05, sigma=.2):
        return {'dt': mu*x, 'dw': sigma*x}

    paths = 100_000_000
    x = f(paths=paths, steps=2001, x0=1., method='milstein', seed=12, output='device',
          getinfo=False)((0., 1.))
    a = m.montecarlo(bins=100)
    for k in range(4):                      # cumulate in 4 chunks
        a.update(x.x[-1, k*paths//4:(k + 1)*paths//4])
    assert a.paths == paths
    counts, edges = a.histogram()
    assert counts.sum() + a.outpaths == paths
    assert abs(float(a.mean()) - np.exp(.05)) < 4*float(a.stderr())
    sdv = np.exp(.05)*np.sqrt(np.exp(.04) - 1)
    assert abs(float(a.std())/sdv - 1) < 1e-3
    one = m.montecarlo(x.x[-1], bins=edges)
    assert np.array_equal(one.histogram()[0], counts)
    assert np.allclose(one.mean(), a.mean(), rtol=1e-12)


def test_long_timeline_few_paths_is_exact():
    """The other extreme the reference talks about ("one million time steps
    across 100 paths", sdepy/__init__.py:87-91): 300 000 steps of 100 paths,
    every step stored.  The Wiener Euler scheme is a running sum, so the
    dumped increments replayed on the host (same operation order) must give
    the stored paths bit for bit -- step staging, store rows and the Philox
    period addressing over a long run."""
    m = sd()
    n, paths = 300_000, 100
    tt = np.linspace(0., 3., n + 1)
    P = m.wiener_process(x0=1., mu=.3, sigma=.7, paths=paths, seed=5)
    P._dump_increments = True
    x = P(tt)
    assert x.shape == (n + 1, paths) and np.isfinite(x).all()
    dW = P._last_run.dump[0]['dW'].cpu().numpy()[:, 0]          # [n, paths]
    ds = np.diff(tt)
    want = np.empty((n + 1, paths))
    want[0] = 1.
    acc = want[0].copy()
    for k in range(n):
        acc = acc + (.3*ds[k] + .7*dW[k])
        want[k + 1] = acc
    assert np.array_equal(np.asarray(x), want)
    # and the increments are N(0, dt) along the whole run
    z = dW/np.sqrt(ds)[:, None]
    assert abs(z.mean()) < 4/np.sqrt(z.size) and abs(z.var() - 1) < 4*np.sqrt(2/z.size)
    assert abs(np.mean(z[:-1]*z[1:])) < 4/np.sqrt(z.size)       # no lag-1 correlation


def test_no_weak_bias_in_the_draws():
    """E[x_N] of Euler GBM is (1 + mu dt)^N exactly when E[dw] = 0 and
    E[dw^2] = dt; a bias of the mean of the normals is amplified by
    N sigma sqrt(dt) = 2.8.  3e8 traced paths (a quarter of a Philox block per
    step) resolve 1.2e-5; the preset lognormal adds the variance at 2e-4."""
    m = sd()

    @m.integrate
    def gbm(t, x, mu=.05, sigma=.2):
        return {'dt': mu*x, 'dw': sigma*x}

    want = (1 + .05/200)**200
    means, var = [], 0.
    for seed in (1, 2, 3):
        st = gbm(paths=100_000_000, steps=201, x0=1., seed=seed, output='stats',
                 getinfo=False)((0., 1.))
        means.append(float(np.asarray(st.pmean())[-1, 0]))
        var += float(np.asarray(st.stderr())[-1, 0])**2
    assert abs(np.mean(means) - want) < 4*np.sqrt(var)/3
    st = m.lognorm_process(paths=200_000_000, steps=201, x0=1., mu=.05, sigma=.2, seed=9,
                           output='stats', getinfo=False)((0., 1.))
    assert abs(float(np.asarray(st.pmean())[-1, 0]) - np.exp(.05)) < \
        4*float(np.asarray(st.stderr())[-1, 0])
    v, wantv = float(np.asarray(st.pvar())[-1, 0]), np.exp(.1)*(np.exp(.04) - 1)
    # var of the sample variance of a lognormal: (kurtosis - 1)/n, kurtosis ~ 3.7
    assert abs(v/wantv - 1) < 4*np.sqrt(2.7/2e8)
