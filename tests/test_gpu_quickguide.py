"""The reference's quick guide (doc/quickguide.rst) walked through on the CUDA
path: the statements of the guide, in order, with its own shape / type checks
and a few statistical ones.  A user switching packages meets exactly these
calls first.  (Plot lines are left out; the one thing that cannot run on the
device -- a Python ``next`` method -- must say so.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def guide():
    import sdepy_b200 as sdepy

    @sdepy.integrate
    def my_process(t, x, theta=1., k=1., sigma=1.):
        return {'dt': k*(theta - x), 'dw': sigma}

    return sdepy, my_process


coarse_timeline = (0., 0.25, 0.5, 0.75, 1.0)
timeline = np.linspace(0., 1., 500)


def test_integrate_decorator_and_parameters(guide):
    """quickguide.rst:29-170."""
    sdepy, my_process = guide
    assert issubclass(my_process, sdepy.integrator) and issubclass(my_process, sdepy.SDE)
    x = my_process(x0=1, paths=100*1000, steps=100)(coarse_timeline)
    assert x.shape == (5, 100000)
    # OU from 1 towards theta = 1: mean stays 1, var = (1 - exp(-2t))/2
    assert abs(x[-1].mean() - 1.) < 4*np.sqrt((1 - np.exp(-2.))/2/1e5)
    assert abs(x[-1].var()/((1 - np.exp(-2.))/2) - 1) < .03
    x = my_process(x0=1, paths=1000, steps=100)(timeline)
    assert x.shape == (500, 1000)

    corr = ((1, .2, -.3), (.2, 1, .1), (-.3, .1, 1))
    x = my_process(x0=1, vshape=3, corr=corr, paths=1000)(timeline)
    assert x.shape == (500, 3, 1000)

    sigma = lambda t: 0.1 + t                                    # noqa: E731
    theta = lambda t: 2 - t                                      # noqa: E731
    k = lambda t: 2/(t + 1)                                      # noqa: E731
    c02 = lambda t: -0.1*np.cos(3*t)                             # noqa: E731
    c12 = lambda t: 0.1*np.sign(0.5 - t)                         # noqa: E731
    corr = lambda t: ((1, -.2, c02(t)), (-.2, 1, c12(t)), (c02(t), c12(t), 1))   # noqa: E731
    x = my_process(x0=1, vshape=3, corr=corr, theta=theta, k=k, sigma=sigma,
                   paths=10*1000)(timeline)
    assert x.shape == (500, 3, 10000)
    # realised correlations of the increments follow corr(t) (the guide plots them)
    dx = np.diff(x, axis=0)
    for n in (50, 250, 450):
        z = dx[n]
        c = np.corrcoef(z)
        want = np.asarray(corr(timeline[n] + (timeline[1] - timeline[0])/2))
        assert np.abs(c - want).max() < 5/np.sqrt(10000)

    # path-dependent x0 and sigma, integration backwards from the last point
    x0, sigma = np.zeros(1000), np.zeros(1000)
    x0[::2], x0[1::2] = 0., 2.
    sigma[::2], sigma[1::2] = 0.5, 0.1
    x = my_process(x0=x0, sigma=sigma, paths=1000, theta=1, k=-2, i0=-1)(timeline)
    assert x.shape == (500, 1000)
    assert (x[-1, :] == x0).all()

    # parameters broadcast against vshape
    sigma = np.linspace(0., 1., 10).reshape(10, 1, 1)
    k = np.linspace(1., 2., 15).reshape(1, 15, 1)
    x = my_process(x0=1, theta=2, k=k, sigma=sigma, vshape=(10, 15),
                   paths=10*1000)(coarse_timeline)
    assert x.shape == (5, 10, 15, 10000)
    # mean reverts from 1 towards 2 at rate k, whatever sigma (steps=None:
    # the four Euler steps of the coarse timeline, dt = 1/4)
    want = 2 - (1 - k[0, :, 0]/4)**4
    got = x[-1].mean(axis=-1)
    assert np.abs(got - want).max() < 5*1./np.sqrt(10000)
    assert x[-1, 0].std(axis=-1).max() < 1e-12            # sigma = 0 row


def test_kfuncs_and_sources(guide):
    """quickguide.rst:58-60, 182-294."""
    sdepy, my_process = guide
    myp = sdepy.kfunc(my_process)
    assert issubclass(myp, sdepy.integrator) and issubclass(myp, sdepy.SDE)
    p = myp(x0=1, sigma=1, paths=1000)
    x = p(timeline)
    x1, x2 = p(timeline, sigma=0.5), p(timeline, sigma=1.5)
    q = p(paths=100, vshape=(3,), k=2)
    y = q(timeline, sigma=0.5)
    assert x.shape == x1.shape == x2.shape == (500, 1000) and y.shape == (500, 3, 100)
    assert x1[-1].std() < x[-1].std() < x2[-1].std()
    x = myp(timeline, x0=1, sigma=1, paths=1000)
    assert x.shape == (500, 1000)
    assert q.params['k'] == 2 and q.params['paths'] == 100 and q.params['vshape'] == (3,)
    assert (sdepy.iskfunc(myp), sdepy.iskfunc(p), sdepy.iskfunc(my_process)) == (True, True, False)

    # a process as the driving source, shared by the three components
    my_dw = sdepy.integrate(lambda t, x: {'dw': 1})(vshape=1, paths=1000)(timeline)
    p = myp(dw=my_dw, vshape=3, paths=1000, x0=1, sigma=((1,), (2,), (3,)))
    x = p(timeline)
    assert x.shape == (500, 3, 1000)
    # same noise, scaled: the deviations from the noiseless path are proportional
    x_det = myp(vshape=3, paths=1, x0=1, sigma=0)(timeline)
    dev = np.asarray(x) - np.asarray(x_det)
    assert np.allclose(dev[:, 1], 2*dev[:, 0], atol=1e-9)
    assert np.allclose(dev[:, 2], 3*dev[:, 0], atol=1e-9)
    x = p(coarse_timeline, steps=timeline)
    assert x.shape == (5, 3, 1000)

    # a Wiener source with memory: same realisations on any timeline
    my_dw = sdepy.true_wiener_source(paths=1000)
    p = myp(x0=1, theta=1, k=1, sigma=1, dw=my_dw, paths=1000)
    t1 = np.linspace(0., 1., 30)
    t2 = np.linspace(0., 1., 100)
    t3 = t = np.linspace(0., 1., 300)
    x1, x2, x3 = p(t1), p(t2), p(t3)
    y1, y2, y3 = p(t, theta=1.5), p(t, theta=1.75), p(t, theta=2)
    # refining the grid converges path by path
    e13 = np.abs(x1(t)[-1] - x3(t)[-1]).mean()
    e23 = np.abs(x2(t)[-1] - x3(t)[-1]).mean()
    assert e23 < e13 < .2
    # same noise, different theta: paths ordered and equally spaced in theta
    assert (y1[-1] < y2[-1]).all() and (y2[-1] < y3[-1]).all()
    assert np.allclose(y3[-1] - y2[-1], y2[-1] - y1[-1], atol=1e-9)


def test_processes_and_montecarlo(guide):
    """quickguide.rst:307-465."""
    sdepy, my_process = guide
    timeline = np.linspace(0., 1., 101)
    x = my_process(x0=1, vshape=3, paths=1000)(timeline)
    assert x.shape == (101, 3, 1000)
    assert type(x) is sdepy.process
    assert np.isclose(timeline, x.t).all()
    assert x.shape == x.t.shape + x.vshape + (x.paths,)
    y = x(coarse_timeline)
    assert y.shape == (5, 3, 1000) and type(y) is np.ndarray
    assert type(x[0]) is np.ndarray and type(x.mean(axis=0)) is np.ndarray
    assert x['t', ::2].shape == (51, 3, 1000)
    assert x['v', 0].shape == (101, 1000)
    y = x['p', :10]
    assert y.shape == (101, 3, 10) and isinstance(y, sdepy.process)
    i_negative = x.min(axis=(0, 1)) < 0
    y = x['p', i_negative]
    assert y.shape == (101, 3, i_negative.sum())
    x_const = x['t', 0]
    x_one_path = x['p', 0]
    y = np.exp(x) - x_const
    z = np.maximum(x, x_one_path)
    assert isinstance(y, sdepy.process) and isinstance(z, sdepy.process)
    assert np.array_equal(y.t, x.t) and np.array_equal(z.t, x.t)

    # a process as a (stochastic, path-dependent) parameter
    stochastic_vol = my_process(x0=1, paths=10*1000)(timeline)
    stochastic_vol_x = sdepy.lognorm_process(x0=1, vshape=3, paths=10*1000, mu=0,
                                             sigma=stochastic_vol)(timeline)
    assert stochastic_vol_x.shape == (101, 3, 10000)
    assert np.isfinite(stochastic_vol_x).all()
    assert abs(stochastic_vol_x[-1].mean() - 1) < .05       # martingale

    cdf = x.cdf(0.5, x=np.linspace(-2, 2, 100))
    chf = x.chf(0.5, u=np.linspace(-2, 2, 100))
    assert cdf.shape == chf.shape == (100, 3)
    assert (np.diff(cdf, axis=0) >= 0).all() and np.iscomplexobj(chf)
    assert x.pstd().shape == (101, 3, 1)
    assert x.tmax().shape == (1, 3, 1000)

    y = x(1)[0]
    a = sdepy.montecarlo(y, bins=30)
    ygrid = np.linspace(y.min(), y.max(), 200)
    pdf, cdf = a.pdf(ygrid), a.cdf(ygrid)
    near = a.pdf(ygrid, method='interp', kind='nearest')
    assert pdf.shape == cdf.shape == near.shape == (200,)
    assert abs(np.trapezoid(pdf, ygrid) - 1) < .02 and (np.diff(cdf) >= -1e-12).all()

    p = my_process(x0=1, vshape=3, paths=10*1000)
    a = sdepy.montecarlo(bins=100)
    for _ in range(10):
        x = p(timeline)
        a.update(x(1))
    assert a.paths == 100000
    assert a[0].pdf(ygrid).shape == (200,)
    assert np.abs(np.asarray(a.mean()) - 1).max() < 5*np.sqrt((1 - np.exp(-2.))/2/1e5)


def test_custom_python_integrator_fails_loudly(guide):
    """quickguide.rst:474-528: a user ``next`` written in Python cannot run
    inside the kernel -- no silent CPU path; the Milstein scheme built in
    covers the guide's purpose (strong order 1) on the device."""
    sdepy, _ = guide

    class my_integrator(sdepy.integrator):
        def next(self):
            raise AssertionError('must not be called')

    class my_SDE(sdepy.SDE):
        def sde(self, t, x):
            return {'dt': 0, 'dw': x}

    class euler(my_SDE, sdepy.integrator):
        pass

    class rk(my_SDE, my_integrator):
        pass

    args = dict(dw=sdepy.true_wiener_source(paths=100), paths=100, x0=10)
    exact = sdepy.lognorm_process(mu=0, sigma=1, **args)((0, 1))[-1].mean()
    err = {}
    for method in ('euler', 'milstein'):
        err[method] = [abs(euler(**args, steps=s, method=method)((0, 1))[-1].mean()/exact - 1)
                       for s in (100, 1000)]
    assert err['milstein'][1] < err['euler'][1] and err['milstein'][1] < 2e-3
    with pytest.raises(NotImplementedError):
        rk(**args, steps=10)((0, 1))


def test_fokker_planck_example(guide):
    """quickguide.rst:545-613: backward Wiener paths from a grid of end points,
    Green function by montecarlo pdf; u2(x, t1) = exp(-k (t1 - t0)) sin(x)."""
    sdepy, _ = guide
    from scipy.integrate import quad
    k = .5
    x0, x1, t0, t1 = 0, 10, 0, 1
    xgrid = np.linspace(x0, x1, 51)
    tgrid = np.linspace(t0, t1, 5)
    xp = sdepy.wiener_process(paths=10000, steps=100, sigma=np.sqrt(2*k),
                              vshape=xgrid.shape, x0=xgrid[..., np.newaxis],
                              i0=-1)(timeline=tgrid)
    assert xp.shape == (5, 51, 10000)
    assert np.array_equal(xp[-1], np.broadcast_to(xgrid[:, None], (51, 10000)))
    a = sdepy.montecarlo(xp, bins=100)
    u2 = np.array([quad(lambda y: np.sin(y)*a[0, j].pdf(y), -np.inf, np.inf)[0]
                   for j in (5, 20, 35)])
    want = np.exp(-k*(t1 - t0))*np.sin(xgrid[[5, 20, 35]])
    assert np.abs(u2 - want).max() < .03
