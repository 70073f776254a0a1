"""Philox-mode quantitative tests against the closed forms of sdepy.analytical
(values generated from the reference, tests/golden/analytical.npz) -- the
counterpart of the reference's tests/test_quant.py:182-353 and test_bs
(:592-654): every preset, several time points, mean / variance / cdf / chf
within Monte Carlo error plus a small allowance for the Euler bias."""
import numpy as np
import pytest

from tests.cases import golden

pytestmark = pytest.mark.gpu

PATHS = 400_000
STEPS = 601          # dt = 0.005 on [0, 3]


def sd():
    import sdepy_b200
    return sdepy_b200


def run(cls, seed, **kw):
    g = golden('analytical')
    tl = np.concatenate(((0.,), g['tq']))
    return g, cls(paths=PATHS, steps=STEPS, seed=seed, output='device', **kw)(tl)


def close(sample_mean, want, var, bias=0.):
    return np.all(np.abs(sample_mean - want) < 4.5*np.sqrt(var/PATHS) + bias)


def test_wiener_mean_var_cdf_chf():
    m = sd()
    g, x = run(m.wiener_process, 1, x0=1., mu=.5, sigma=.8)
    mean, var = np.asarray(x.pmean())[1:, 0], np.asarray(x.pvar())[1:, 0]
    assert close(mean, g['wiener_mean'], g['wiener_var'])
    assert np.all(np.abs(var/g['wiener_var'] - 1) < 5*np.sqrt(2/PATHS))
    assert np.abs(x.cdf(g['tq'], g['xq']) - g['wiener_cdf']).max() < 4.5*.5/np.sqrt(PATHS)
    assert np.abs(x.chf(g['tq'], g['uq']) - g['wiener_chf']).max() < 5/np.sqrt(PATHS)


def test_lognorm_mean_std_cdf():
    m = sd()
    g, x = run(m.lognorm_process, 2, x0=1., mu=.05, sigma=.3)
    mean, std = np.asarray(x.pmean())[1:, 0], np.asarray(x.pstd())[1:, 0]
    assert close(mean, g['lognorm_mean'], g['lognorm_std']**2)
    assert np.all(np.abs(std/g['lognorm_std'] - 1) < 2e-2)
    assert np.abs(x.cdf(g['tq'], g['xq'][2:] + .2) - g['lognorm_cdf']).max() < 4.5*.5/np.sqrt(PATHS)


def test_ornstein_uhlenbeck_and_hull_white_2factor():
    m = sd()
    g, x = run(m.ornstein_uhlenbeck_process, 3, x0=1., theta=.3, k=1.5, sigma=.4)
    mean, var = np.asarray(x.pmean())[1:, 0], np.asarray(x.pvar())[1:, 0]
    assert close(mean, g['oruh_mean'], g['oruh_var'], bias=2e-3)
    assert np.all(np.abs(var/g['oruh_var'] - 1) < 5*np.sqrt(2/PATHS) + 1e-2)   # O(dt) Euler bias
    g, h = run(m.hull_white_process, 4, factors=2, x0=((.2,), (.1,)), theta=((.3,), (0.,)),
               k=((1.,), (.5,)), sigma=((.2,), (.3,)), rho=.4)
    mean, var = np.asarray(h.pmean())[1:, 0], np.asarray(h.pvar())[1:, 0]
    assert close(mean, g['hw2f_mean'], g['hw2f_var'], bias=2e-3)
    assert np.all(np.abs(var/g['hw2f_var'] - 1) < 5*np.sqrt(2/PATHS) + 1e-2)


def test_cox_ingersoll_ross():
    m = sd()
    g, x = run(m.cox_ingersoll_ross_process, 5, x0=.5, theta=.3, k=1.2, xi=.3)
    mean, var = np.asarray(x.pmean())[1:, 0], np.asarray(x.pvar())[1:, 0]
    assert close(mean, g['cir_mean'], g['cir_var'], bias=1e-3)
    assert np.all(np.abs(var/g['cir_var'] - 1) < 5*np.sqrt(3/PATHS) + 1.5e-2)


def test_heston_log_moments_and_chf():
    m = sd()
    g, x = run(m.heston_process, 6, x0=1., mu=.05, sigma=1., y0=.04, theta=.05, k=1.5,
               xi=.3, rho=-.6)
    lx = m.device_process(x.t, x.x.log())
    mean, var = np.asarray(lx.pmean())[1:, 0], np.asarray(lx.pvar())[1:, 0]
    assert close(mean, g['heston_log_mean'], g['heston_log_var'], bias=5e-4)
    assert np.all(np.abs(var/g['heston_log_var'] - 1) < 5*np.sqrt(3/PATHS) + 1e-2)
    assert np.abs(lx.chf(g['tq'], g['uq']) - g['heston_log_chf']).max() < 5/np.sqrt(PATHS) + 3e-3


def test_merton_and_kou_mean_and_log_chf():
    m = sd()
    g, x = run(m.merton_jumpdiff_process, 7, x0=1., mu=.05, sigma=.2, lam=2., a=-.1, b=.15)
    xs = x.x.cpu().numpy()[1:]
    assert np.all(np.abs(xs.mean(-1) - g['mjd_mean']) < 4.5*xs.std(-1)/np.sqrt(PATHS) + 2e-3)
    lx = m.device_process(x.t, x.x.log())
    assert np.abs(lx.chf(g['tq'], g['uq']) - g['mjd_log_chf']).max() < 5/np.sqrt(PATHS) + 4e-3
    g, x = run(m.kou_jumpdiff_process, 8, x0=1., mu=.05, sigma=.2, lam=2., a=.1, b=.15, pa=.4)
    xs = x.x.cpu().numpy()[1:]
    assert np.all(np.abs(xs.mean(-1) - g['kou_mean']) < 4.5*xs.std(-1)/np.sqrt(PATHS) + 2e-3)
    lx = m.device_process(x.t, x.x.log())
    assert np.abs(lx.chf(g['tq'], g['uq']) - g['kou_log_chf']).max() < 5/np.sqrt(PATHS) + 4e-3


def test_black_scholes_call_put_within_3_stderr():
    """Reference tests/test_quant.py::test_bs: discounted payoff of a lognormal
    underlying vs bscall / bsput (analytical.py:934-1035)."""
    m = sd()
    g = golden('analytical')
    for T, call, put in zip(g['tq'], g['bscall'], g['bsput']):
        for kind, want in (('call', call), ('put', put)):
            st = m.lognorm_process(paths=2_000_000, steps=101, x0=1., mu=.03, sigma=.25,
                                   seed=int(100*T) + (kind == 'put'), output='stats',
                                   payoff=(kind, 1.1, float(np.exp(-.03*T))))((0., T))
            price = float(np.asarray(st.payoff_mean())[-1, 0])
            err = float(np.asarray(st.payoff_stderr())[-1, 0])
            assert abs(price - want) < 3.5*err, (T, kind, price, want, err)


def test_kfunc_shortcuts_evaluate_on_the_device():
    """``sdepy.lognorm(...)(timeline, paths=..., steps=...)`` and the one-shot
    ``sdepy.dw(t, dt, paths=...)`` forms (reference shortcuts.py:73-99)."""
    import sdepy_b200 as m
    tt = np.linspace(0., 1., 5)
    P = m.lognorm(x0=1., mu=.05, sigma=.2, seed=3)
    x = P(tt, paths=20000, steps=50)
    assert isinstance(x, m.process) and x.shape == (5, 20000)
    assert abs(float(x[-1].mean()) - np.exp(.05)) < 4*.2123744/np.sqrt(20000)
    z = m.dw(0., .25, paths=50000, seed=4)
    assert tuple(z.shape) == (50000,)
    assert abs(float(z.std()) - .5) < .01
    xy = m.heston_xy(tt, paths=1000, steps=20, seed=5)
    assert len(xy) == 2 and xy[0].shape == (5, 1000)


def test_quickguide_basket_lookback_option():
    """The reference's worked example (doc/quickguide.rst:616-671): 4 correlated
    lognormal securities, dividend yields as a constant process, a term
    structure of volatility as a process (interpolated in time), antithetic
    paths through ``dw=odd_wiener_source``; knock-in lookback call on the
    basket.  The guide prints 4.997 +/- 0.027.  Evaluated twice: on the host
    ``process`` with the guide's NumPy expressions, and on the resident slab
    with the device summaries -- same payoffs, bit for bit."""
    import sdepy_b200 as sdepy
    corr = [[1, 0.50, 0.37, 0.35], [0.50, 1, 0.47, 0.46],
            [0.37, 0.47, 1, 0.19], [0.35, 0.46, 0.19, 1]]
    dividend_yield = sdepy.process(c=(0.20, 4.40, 0., 4.80))/100
    riskfree = 0
    vol_timepoints = (0.1, 0.2, 0.5, 1, 2, 3)
    vol = np.array([[0.40, 0.38, 0.30, 0.28, 0.27, 0.27],
                    [0.31, 0.29, 0.22, 0.16, 0.18, 0.21],
                    [0.24, 0.22, 0.19, 0.19, 0.21, 0.22],
                    [0.35, 0.31, 0.21, 0.18, 0.19, 0.19]])
    sigma = sdepy.process(t=vol_timepoints, v=vol.T)
    assert sigma.shape == (6, 4, 1)
    maturity = 2
    timeline = np.linspace(0, maturity, 4*maturity + 1)
    kw = dict(x0=100, corr=corr, dw=sdepy.odd_wiener_source,
              mu=(riskfree - dividend_yield), sigma=sigma,
              vshape=4, paths=100*1000, steps=maturity*250, seed=77)
    x = sdepy.lognorm_process(**kw)(timeline)
    assert x.shape == (9, 4, 100000)
    x_worst = x.min(axis=1)
    x_basket = x.mean(axis=1)
    down_and_in_paths = (x_worst.min(axis=0) < 80)
    lookback_x_basket = x_basket.max(axis=0)
    payoff = np.maximum(0, lookback_x_basket - 105)
    payoff[np.logical_not(down_and_in_paths)] = 0
    a = sdepy.montecarlo(np.asarray(payoff), use='even')
    price, err = float(a.mean()), float(a.stderr())
    assert .02 < err < .035
    assert abs(price - 4.997) < 3*np.hypot(err, .027)

    # the same on the device: nothing but the price leaves HBM
    d = sdepy.lognorm_process(output='device', **kw)(timeline)
    knocked = d.vmin().tmin().x[0] < 80
    pay = (d.vmean().tmax().x[0] - 105).clamp_min(0.)*knocked
    assert np.array_equal(pay.cpu().numpy(), np.asarray(payoff))
    b = sdepy.montecarlo(pay, use='even')
    assert abs(float(b.mean()) - price) < 1e-12 and abs(float(b.stderr()) - err) < 1e-12
