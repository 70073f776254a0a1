"""GPU tests of the in-kernel Philox draws: generator accuracy, distribution,
cross-mode consistency (philox == replay of its own dumped increments) and
moments against closed forms / the oracle within 3-4 standard errors."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import sde_oracle as orc

pytestmark = pytest.mark.gpu


def sd():
    import sdepy_b200
    return sdepy_b200


def _test_normals(n, seed=1234):
    from sdepy_b200 import _lib, _cuda
    zf = torch.empty(2*n, dtype=torch.float64, device='cuda')
    zl = torch.empty(2*n, dtype=torch.float64, device='cuda')
    _lib.check(_lib.lib.sdeb_test_normals(seed, n, _cuda.ptr(zf), _cuda.ptr(zl),
                                          _cuda.stream_ptr(zf.device)))
    return zf, zl


def test_normal_pair_matches_libdevice():
    """First half of the sample: the default 64-bit map (32-bit radius uniform,
    32-bit direction); second half: the 96-bit full-resolution map.  The first
    128 pairs of each half force the corners of the bit space."""
    import scipy.stats
    n = 1 << 21
    zf, zl = (z.cpu().numpy() for z in _test_normals(n))
    assert np.isfinite(zf).all()
    # hand-rolled log / sqrt / sincos vs libdevice on identical bits.  Both
    # maps carry ~2e-16 ABSOLUTE error in s2 = -2 ln u, i.e. ~1e-16/r in
    # z = r (cos, sin) -- only visible in the forced corner u -> 1 (r -> 0)
    r = np.maximum(np.hypot(zl[0::2], zl[1::2]), 1e-300).repeat(2)
    assert (np.abs(zf - zl) < 4e-15 + 3e-16/r).all()
    for half, rmax in ((0, np.sqrt(2*33*np.log(2.))), (1, np.sqrt(2*53*np.log(2.)))):
        z = zf[half*n:(half + 1)*n]
        zc, z = z[:256], z[256:]
        assert np.abs(z - zl[half*n:(half + 1)*n][256:]).max() < 1e-13
        # the forced smallest u reaches the advertised largest radius
        assert np.isclose(np.hypot(zc[2], zc[3]), rmax, rtol=1e-12)
        # radius tail: P(u < 2^-k) = 2^-k (the exponent comes from the uniform's own
        # exponent field: no leading-zero count, no tail draw)
        r2 = z[0::2]**2 + z[1::2]**2
        for k in (8, 12, 14, 16):
            want = r2.size*2.**-k
            got = (r2 > 2*k*np.log(2.)).sum()
            assert abs(got - want) < 5*np.sqrt(want), (k, got, want)
        assert abs(z.mean()) < 4/np.sqrt(z.size)
        assert abs(z.var() - 1) < 4*np.sqrt(2/z.size)
        assert abs((z**4).mean() - 3) < 4*np.sqrt(96/z.size)
        assert scipy.stats.kstest(z[:200000], 'norm').pvalue > 1e-4
        # pairs are uncorrelated
        assert abs(np.mean(z[0::2]*z[1::2])) < 4/np.sqrt(z.size/2)


def test_normal_draws_fine_grid_and_deep_tails():
    """2^30 default-map normals, reduced on the device: (i) chi-square of
    Phi(z) over 2^16 equiprobable cells (fine-grid uniformity), (ii) two-sided
    tail counts down to 2^-24, each within 5 sigma of its binomial expectation."""
    cells = 1 << 16
    hist = torch.zeros(cells, dtype=torch.float64, device='cuda')
    ks = (12, 16, 20, 24)
    import scipy.stats
    thr = [float(scipy.stats.norm.isf(2.**-(k + 1))) for k in ks]      # P(|z| > thr) = 2^-k
    tails = np.zeros(len(ks))
    total = 0
    n = 1 << 26                       # pairs per call: 2^27 doubles per array
    for it in range(8):
        zf, _ = _test_normals(n, seed=77 + it)
        z = zf[:n][256:]              # default map, corners dropped
        p = torch.special.ndtr(z)
        idx = torch.clamp((p*cells).to(torch.int64), 0, cells - 1)
        hist += torch.bincount(idx, minlength=cells).to(torch.float64)
        az = z.abs()
        tails += np.array([float((az > t).sum()) for t in thr])
        total += z.numel()
        del zf, z, p, idx, az
    h = hist.cpu().numpy()
    assert h.sum() == total
    exp = total/cells
    chi2 = ((h - exp)**2/exp).sum()
    # chi-square with cells-1 degrees of freedom: mean cells-1, sd sqrt(2(cells-1))
    assert abs(chi2 - (cells - 1)) < 5*np.sqrt(2*(cells - 1)), chi2
    for k, got in zip(ks, tails):
        want = total*2.**-k
        assert abs(got - want) < 5*np.sqrt(want) + 1, (k, got, want)


def test_philox_mode_equals_replay_of_its_own_increments():
    """Same step arithmetic in both noise modes: dump the Philox increments,
    feed them to the CPU oracle and to the kernel in replay mode."""
    m = sd()
    paths, n = 5000, 40
    grid = np.linspace(0., 1., n + 1)
    par = dict(mu=.03, sigma=1., theta=.04, k=2., xi=.9)
    P = m.full_heston_process(paths=paths, steps=grid, x0=100., y0=.04, rho=-.7,
                              seed=11, **par)
    P._dump_increments = True
    x, y = P((0., .5, 1.))
    dW = P._last_run.dump[0]['dW'].cpu().numpy()
    (ox, oy), oinfo = orc.euler_replay('heston', par, 100., grid, [0, 20, 40],
                                       dW, y0=.04, full=True)
    assert np.array_equal(np.asarray(y), oy)
    assert np.abs(np.asarray(x)/ox - 1).max() <= 4*np.finfo(float).eps
    assert np.array_equal(P.info['negative_y_count'], oinfo['negative_y_count'])
    # distribution of the dumped increments
    z = dW/np.sqrt(np.diff(grid))[:, None, None]
    assert abs(z.mean()) < 4/np.sqrt(z.size)
    c = np.corrcoef(z[:, 0].ravel(), z[:, 1].ravel())[0, 1]
    assert abs(c + .7) < 4/np.sqrt(z.size/2)
    # reproducible: same seed, same stream; different call, new stream
    Q = m.full_heston_process(paths=paths, steps=grid, x0=100., y0=.04, rho=-.7,
                              seed=11, **par)
    x2, y2 = Q((0., .5, 1.))
    # (this run takes the lean kernel, whose contracted arithmetic agrees with
    # the reference-exact kernel of the dump run to rounding level)
    assert np.allclose(np.asarray(y2), np.asarray(y), rtol=1e-11, atol=1e-13)
    assert np.allclose(np.asarray(x2), np.asarray(x), rtol=1e-11)
    R = m.full_heston_process(paths=paths, steps=grid, x0=100., y0=.04, rho=-.7,
                              seed=11, **par)
    x4, y4 = R((0., .5, 1.))
    assert np.array_equal(np.asarray(y4), np.asarray(y2))
    assert np.array_equal(np.asarray(x4), np.asarray(x2))
    x3, y3 = Q((0., .5, 1.))
    assert not np.allclose(np.asarray(y3), np.asarray(y))


def test_sharding_is_invisible():
    """Counter = global path index: two shards reproduce the single run."""
    m = sd()
    kw = dict(steps=30, x0=1., mu=.05, sigma=.2, seed=5)
    full = np.asarray(m.lognorm_process(paths=1000, **kw)((0., 1.)))
    a = np.asarray(m.lognorm_process(paths=600, path_offset=0, **kw)((0., 1.)))
    b = np.asarray(m.lognorm_process(paths=400, path_offset=600, **kw)((0., 1.)))
    assert np.array_equal(full, np.concatenate((a, b), axis=-1))


def test_config1_lognorm_terminal_moments():
    """BASELINE config 1: lognorm 1e5 paths x 250 steps, pmean/pstd vs the
    closed forms of sdepy.analytical.lognorm_mean/std (analytical.py:133,165)."""
    m = sd()
    paths = 100_000
    x = m.lognorm_process(x0=1., mu=.05, sigma=.2, paths=paths, seed=1234,
                          output='device')(np.linspace(0, 1, 251))
    mean, std = float(np.asarray(x.pmean())[-1, 0]), float(np.asarray(x.pstd())[-1, 0])
    em = np.exp(.05)
    es = em*np.sqrt(np.exp(.2**2) - 1)
    assert abs(mean - em) < 3.5*es/np.sqrt(paths)
    assert abs(std - es) < 3.5*es*np.sqrt(1.5/paths)   # lognormal kurtosis margin
    xs = np.asarray(x)
    assert np.allclose(np.asarray(x.pmean())[..., 0], xs.mean(axis=-1), rtol=1e-13)
    assert np.allclose(np.asarray(x.pvar())[..., 0], xs.var(axis=-1), rtol=1e-11)


def test_ou_hw_discrete_euler_moments():
    """Linear SDEs: compare with the exact moments of the DISCRETE Euler
    recursion m <- m + k(theta-m)dt, v <- (1-k dt)^2 v + sigma^2 dt."""
    m = sd()
    paths, n = 400_000, 100
    k, theta, sigma, T = 1., .2, .3, 2.
    x = m.ornstein_uhlenbeck_process(paths=paths, steps=n + 1, x0=.1, theta=theta,
                                     k=k, sigma=sigma, seed=3, output='device')((0., T))
    dt = T/n
    mm, vv = .1, 0.
    for _ in range(n):
        mm, vv = mm + k*(theta - mm)*dt, (1 - k*dt)**2*vv + sigma**2*dt
    mean = float(np.asarray(x.pmean())[-1, 0]); var = float(np.asarray(x.pvar())[-1, 0])
    assert abs(mean - mm) < 4*np.sqrt(vv/paths)
    assert abs(var - vv) < 4*vv*np.sqrt(2/paths)


def test_heston_stats_vs_oracle_sample():
    """Philox Heston vs the oracle's own numpy-driven sample (the reference's
    draws): log-mean and log-variance within combined standard errors; fused
    'stats' output equals the statistics of the stored paths."""
    m = sd()
    paths = 400_000
    par = dict(mu=.03, sigma=1., theta=.04, k=2., xi=.3)
    grid = np.linspace(0., 1., 65)
    kw = dict(paths=paths, steps=grid, x0=100., y0=.04, rho=-.7, seed=99, **par)
    xd = m.heston_process(output='device', **kw)((0., 1.))
    xT = xd.x[-1].cpu().numpy()
    oT, _ = orc.heston_stream(par, 100., .04, -.7, grid, 200_000,
                              np.random.default_rng(1))
    la, lb = np.log(xT), np.log(oT)
    se = np.sqrt(la.var()/la.size + lb.var()/lb.size)
    assert abs(la.mean() - lb.mean()) < 4*se
    assert abs(la.var() - lb.var()) < 5*la.var()*np.sqrt(2/la.size + 2/lb.size)*1.5
    # fused statistics (same seed => same paths)
    st = m.heston_process(output='stats', payoff=('call', 100., np.exp(-.03)), **kw)((0., 1.))
    assert np.allclose(np.asarray(st.pmean())[-1, 0], xT.mean(), rtol=1e-12)
    assert np.allclose(np.asarray(st.pvar())[-1, 0], xT.var(), rtol=1e-10)
    pay = np.maximum(xT - 100., 0.)*np.exp(-.03)
    assert np.allclose(np.asarray(st.payoff_mean())[-1, 0], pay.mean(), rtol=1e-12)
    assert np.asarray(st.pmin())[-1, 0] == xT.min()
    assert np.asarray(st.pmax())[-1, 0] == xT.max()
    # Gil-Pelaez price from sdepy.analytical.heston_log_chf:
    # 9.2425; Euler/full-truncation bias is below the 4-se band at this size
    price, err = np.asarray(st.payoff_mean())[-1, 0], np.asarray(st.payoff_stderr())[-1, 0]
    assert abs(price - 9.2425) < 4*err + 0.03


def test_merton_kou_moments_and_jump_counts():
    m = sd()
    paths = 200_000
    lam, a, b = 2., -.1, .15
    P = m.merton_jumpdiff_process(paths=paths, steps=101, x0=1., mu=.05, sigma=.2,
                                  lam=lam, a=a, b=b, seed=4)
    P._dump_increments = True
    x = np.asarray(P((0., 1.)))
    lx = np.log(x[-1])
    mean = (.05 - .02) + lam*a
    var = .04 + lam*(a*a + b*b)
    assert abs(lx.mean() - mean) < 4*np.sqrt(var/paths)
    assert abs(lx.var() - var) < 6*var*np.sqrt(2/paths)
    jc = P.info['jump_count']
    assert abs(jc.mean() - lam) < 4*np.sqrt(lam/paths)
    assert abs(jc.var() - lam) < 6*lam*np.sqrt(2/paths)
    d = P._last_run.dump[0]
    dN = d['dN'].cpu().numpy()
    assert np.array_equal(dN.sum(axis=0).reshape(jc.shape), jc)
    assert np.allclose(P.info['jump_rate'][0], dN.sum()/paths, rtol=1e-12)
    # replaying the dumped increments reproduces the run bit for bit
    R = m.merton_jumpdiff_process(
        paths=paths, steps=101, x0=1., mu=.05, sigma=.2,
        dw=m.replay_source(d['dW'].reshape(100, paths)),
        dj=m.replay_source(d['dJ'].reshape(100, paths), dn=d['dN'].reshape(100, paths)))
    xr = np.asarray(R((0., 1.)))
    assert np.array_equal(xr, x)
    assert np.array_equal(R.info['jump_count'], jc)
    # Kou: mean of the double exponential = pa*a - (1-pa)*b
    K = m.kou_jumpdiff_process(paths=paths, steps=101, x0=1., mu=.05, sigma=.2,
                               lam=lam, a=.1, b=.15, pa=.4, seed=6)
    lk = np.log(np.asarray(K((0., 1.)))[-1])
    ym = .4*.1 - .6*.15
    yv = .4*.6*(.25)**2 + (.4*.01 + .6*.0225)
    assert abs(lk.mean() - (.03 + lam*ym)) < 4*np.sqrt((.04 + lam*(yv + ym*ym))/paths)


def test_time_dependent_correlation_hw3():
    m = sd()
    from tests.cases import HW
    paths = 100_000
    P = m.hull_white_process(paths=paths, steps=51, seed=8, **HW)
    P._dump_increments = True
    x = P(np.linspace(0, 5, 11))
    assert x.shape == (11, paths) and np.isfinite(np.asarray(x)).all()
    dW = P._last_run.dump[0]['dW'].cpu().numpy()      # [50, 3, paths]
    grid = np.linspace(0, 5, 51)
    for n in (0, 25, 49):
        z = dW[n]/np.sqrt(grid[n + 1] - grid[n])
        c = np.corrcoef(z)
        want = HW['corr'](grid[n] + (grid[n + 1] - grid[n])/2)
        assert np.abs(c - want).max() < 5/np.sqrt(paths)


def test_increment_stream_quality():
    """Statistical checks of the in-kernel draws beyond the first moments:
    distribution (KS), tails, independence across steps and across neighbouring
    paths, independence of the two normals of a Box-Muller pair across the
    odd/even step reuse (one-factor models keep the second normal for the next
    step)."""
    import scipy.stats
    m = sd()
    paths, n = 200_000, 64
    P = m.wiener_process(paths=paths, steps=n + 1, seed=2025)
    P._dump_increments = True
    P((0., 1.))
    z = (P._last_run.dump[0]['dW'].cpu().numpy()[:, 0, :]*np.sqrt(n))     # [n, paths] ~ N(0,1)
    N = z.size
    assert scipy.stats.kstest(z[::7].ravel()[:500_000], 'norm').pvalue > 1e-3
    for thr in (2., 3., 4.):
        p = 2*scipy.stats.norm.sf(thr)
        k = (np.abs(z) > thr).sum()
        assert abs(k - N*p) < 5*np.sqrt(N*p), thr
    # moments up to 6
    for k, want in ((2, 1.), (4, 3.), (6, 15.)):
        se = np.sqrt((scipy.stats.norm.moment(2*k) - want**2)/N)
        assert abs((z**k).mean() - want) < 5*se
    se = 5/np.sqrt(paths*(n - 1))
    assert abs(np.mean(z[1:]*z[:-1])) < se                   # lag-1 across steps (pair reuse!)
    assert abs(np.mean(z[2:]*z[:-2])) < se
    assert abs(np.mean(z[:, 1:]*z[:, :-1])) < se             # neighbouring paths
    assert abs(np.mean((z[1::2]**2 - 1)*(z[0::2]**2 - 1))) < 3*se   # pair members independent
    # two different seeds / two calls: unrelated streams
    Q = m.wiener_process(paths=paths, steps=n + 1, seed=2026)
    Q._dump_increments = True
    Q((0., 1.))
    z2 = Q._last_run.dump[0]['dW'].cpu().numpy()[:, 0, :]*np.sqrt(n)
    assert abs(np.mean(z*z2)) < 5/np.sqrt(N)


def _normal_draws(m, draws, groups, paths, seed):
    """groups x paths standard normals, one single-step Wiener path each
    (x(1) = 0 + dw exactly), resident in HBM."""
    x = m.wiener_process(paths=paths, vshape=(groups,), x0=0., mu=0., sigma=1., seed=seed,
                         draws=draws, output='device')((0., 1.))
    return x.x[1]                                    # [groups, paths]


@pytest.mark.parametrize('draws', ['fast', 'full'])
def test_normal_draws_fine_grid_uniformity_and_far_tails(draws):
    """Beyond the 2^-16 tail / 2e5-sample KS checks above: 1e9 draws of each
    resolution (default: 64 random bits per pair, 32-bit radius uniform;
    draws='full': 96 bits, 52-bit radius uniform).
      * chi-square of Phi(z) on 2^16 equal cells (65535 d.o.f.: mean 65535,
        sd 362) within 5 sd;
      * two-sided tail counts at 2^-12, 2^-16, 2^-20, 2^-24 within 5 sd of
        their binomial expectation (2^-24 of 1e9 draws = 59.6 +/- 7.7)."""
    import torch
    from scipy import stats
    m = sd()
    groups, paths = 10, 100_000_000
    z = _normal_draws(m, draws, groups, paths, seed=77)
    n = groups*paths
    cells = 1 << 16
    counts = torch.zeros(cells, dtype=torch.int64, device=z.device)
    tails = {k: 0 for k in (12, 16, 20, 24)}
    thr = {k: float(stats.norm.isf(2.0**-(k + 1))) for k in tails}
    s1 = s2 = s4 = 0.
    for g in range(groups):
        row = z[g]
        u = torch.special.ndtr(row)
        idx = torch.clamp((u*cells).to(torch.int64), 0, cells - 1)
        counts += torch.bincount(idx, minlength=cells)
        a = row.abs()
        for k in tails:
            tails[k] += int((a > thr[k]).sum())
        s1 += float(row.sum()); s2 += float((row*row).sum()); s4 += float((row**4).sum())
        del u, idx, a
    assert int(counts.sum()) == n
    expect = n/cells
    chi2 = float(((counts.double() - expect)**2).sum()/expect)
    assert abs(chi2 - (cells - 1)) < 5*np.sqrt(2*(cells - 1)), chi2
    for k, c in tails.items():
        p = 2.0**-k
        assert abs(c - n*p) < 5*np.sqrt(n*p*(1 - p)) + 1, (k, c, n*p)
    assert abs(s1/n) < 5/np.sqrt(n)
    assert abs(s2/n - 1) < 5*np.sqrt(2/n)
    assert abs(s4/n - 3) < 5*np.sqrt(96/n)
    if draws == 'full':
        # the 52-bit radius reaches beyond the 32-bit map's |z| <= 6.76
        assert float(z.abs().max()) > 5.5


def test_full_resolution_draws_heston_price_and_stream_independence():
    """draws='full' runs the same kernels compiled with -DSDEB_DRAW_FULL=1:
    Monte Carlo price within 4 standard errors of the closed form (+ the Euler
    bias bound), results independent of sharding, and a stream distinct from
    the default resolution's."""
    m = sd()
    grid = np.linspace(0., 1., 253)
    kw = dict(steps=grid, x0=100., mu=.03, sigma=1., y0=.04, theta=.04, k=2., xi=.3, rho=-.7,
              seed=5, output='stats', payoff=('call', 100., float(np.exp(-.03))), getinfo=False)
    n = 4_000_000
    full = m.heston_process(paths=n, draws='full', **kw)((0., 1.))
    fast = m.heston_process(paths=n, **kw)((0., 1.))
    pf, ef = (float(np.asarray(z)[-1, 0]) for z in (full.payoff_mean(), full.payoff_stderr()))
    assert abs(pf - 9.2425) < 4*ef + 1e-2
    assert full.sums[-1].ravel()[0] != fast.sums[-1].ravel()[0]
    a = m.heston_process(paths=n//4, draws='full', **kw)((0., 1.))
    b = m.heston_process(paths=n - n//4, path_offset=n//4, draws='full', **kw)((0., 1.))
    assert np.allclose((a.sums + b.sums)[-1].ravel()[:4], full.sums[-1].ravel()[:4], rtol=1e-10,
                       atol=1e-6)
    with pytest.raises(ValueError):
        m.heston_process(paths=10, draws='53bit')
