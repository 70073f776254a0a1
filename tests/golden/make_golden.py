"""
Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED
REFERENCE (/root/reference/sdepy).  Run once in the authoring container:

    python tests/golden/make_golden.py

The reference is only importable there; the fixtures (small .npz files)
travel with the repo and pin both the CPU oracle (tests/test_oracle_golden.py,
no GPU) and the CUDA path (tests/test_gpu_parity.py, -m gpu).

Two kinds of fixtures:

* ``replay_*``: the reference integrates a preset process while a recording
  wrapper (obeying the reference's source protocol, infrastructure.py:
  1321-1330) logs every increment its own sources return.  Stored: the
  parameters, timeline, merged step grid, the logged increments, the
  reference output and ``info`` counters.
* ``seeded_*``: plain same-seed runs (``rng=np.random.default_rng(seed)``),
  storing only inputs and outputs -- these pin the oracle's restatement of
  the sources' generator consumption.
* ``stats_*``: ``process.pmean/pvar/pstd`` and ``montecarlo`` results.
"""
import os
import sys

import numpy as np

sys.path.insert(0, '/root/reference')
import sdepy  # noqa: E402
from sdepy import infrastructure as infra  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


class recorder:
    """Source-protocol wrapper logging (t, dt, value[, dn_value])."""

    def __init__(self, inner):
        self.inner = inner
        self.paths, self.vshape = inner.paths, inner.vshape
        self.t, self.dt, self.dz, self.dn = [], [], [], []
        self._has_dn = False

    def __call__(self, t, dt):
        z = self.inner(t, dt)
        self.t.append(float(t)); self.dt.append(float(dt))
        self.dz.append(np.array(z))
        if hasattr(self.inner, 'dn_value'):
            # forwarded because jumpdiff_SDE.info_next probes it
            # (integration.py:2615)
            self.dn_value = self.inner.dn_value
            self.dn.append(np.array(self.inner.dn_value))
        return z


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-34s %7.1f KB' % (name, os.path.getsize(path)/1024))


def grid_of(P, timeline):
    tt = np.asarray(timeline, dtype=float)
    target = P.pace(tt)
    target = target[(target >= tt[0]) & (target <= tt[-1])]
    return np.unique(np.concatenate((target, tt)))


def replay_case(name, cls, timeline, paths, steps, vshape=(), seed=7,
                jumps=False, **params):
    rng = np.random.default_rng(seed)
    probe = cls(paths=paths, vshape=vshape, steps=steps, rng=rng, **params)
    rec_w = recorder(probe.sources['dw'])
    kw = dict(dw=rec_w)
    if jumps:
        rec_j = recorder(probe.sources['dj'])
        kw['dj'] = rec_j
    # second instance consuming the recording wrappers
    src_keys = ('corr', 'rho', 'lam', 'a', 'b', 'pa', 'y', 'ptype')
    P = cls(paths=paths, vshape=vshape, steps=steps,
            **{k: v for k, v in params.items() if k not in src_keys}, **kw)
    out = P(timeline)
    outs = out if isinstance(out, tuple) else (out,)
    arrays = dict(tt=np.asarray(timeline, dtype=float),
                  grid=grid_of(P, timeline),
                  dW=np.stack(rec_w.dz), t=np.array(rec_w.t),
                  dt=np.array(rec_w.dt))
    for i, o in enumerate(outs):
        arrays['out%d' % i] = np.asarray(o)
    if jumps:
        arrays['dJ'] = np.stack(rec_j.dz)
        arrays['dN'] = np.stack(rec_j.dn)
        arrays['jump_count'] = P.info['jump_count']
        arrays['jump_rate'] = P.info['jump_rate']
    if 'negative_y_count' in P.info:
        arrays['negative_y_count'] = P.info['negative_y_count']
    arrays['computed_steps'] = np.array(P.info['computed_steps'])
    arrays['stored_steps'] = np.array(P.info['stored_steps'])
    for k, v in params.items():
        if not callable(v):
            arrays['p_' + k] = np.asarray(v)
    save(name, **arrays)


def seeded_case(name, cls, timeline, paths, steps, seed, vshape=(), **params):
    P = cls(paths=paths, vshape=vshape, steps=steps,
            rng=np.random.default_rng(seed), **params)
    out = P(timeline)
    outs = out if isinstance(out, tuple) else (out,)
    arrays = dict(tt=np.asarray(timeline, dtype=float), seed=np.array(seed))
    for i, o in enumerate(outs):
        arrays['out%d' % i] = np.asarray(o)
    for k in ('negative_y_count', 'jump_count', 'jump_rate'):
        if k in P.info:
            arrays[k] = P.info[k]
    save(name, **arrays)


def theta_t(t):
    return .2 + .1*t


def hw_theta(t):
    return np.array(((.02 + .001*t,), (0.,), (0.,)))


def hw_corr(t):
    c01, c02, c12 = .3*np.cos(t), -.2 + .05*t, .1
    return np.array(((1, c01, c02), (c01, 1, c12), (c02, c12, 1)))


HW = dict(factors=3, x0=((.01,), (0.,), (0.,)), theta=hw_theta,
          k=((.1,), (.5,), (1.,)), sigma=((.01,), (.008,), (.005,)),
          corr=hw_corr)
HESTON = dict(x0=100., mu=.03, sigma=1., y0=.04, theta=.04, k=2., xi=.3,
              rho=-.7)
MERTON = dict(x0=1., mu=.05, sigma=.2, lam=2., a=-.1, b=.15)
KOU = dict(x0=1., mu=.05, sigma=.2, lam=2., a=.1, b=.15, pa=.4)


def main():
    P = 193  # deliberately not a multiple of 32
    # ---- replay fixtures (BASELINE.json configs, scaled down) ------------
    replay_case('replay_wiener', sdepy.wiener_process, np.linspace(0, 1, 6),
                P, 21, x0=.5, mu=.1, sigma=.7)
    replay_case('replay_lognorm', sdepy.lognorm_process,
                np.linspace(0, 1, 11), P, 41, x0=1., mu=.05, sigma=.2)
    replay_case('replay_lognorm_v3', sdepy.lognorm_process,
                (0., .3, 1.), 67, 17, vshape=(3,), x0=((1.,), (2.,), (3.,)),
                mu=((.05,), (.0,), (-.1,)), sigma=((.2,), (.3,), (.1,)),
                corr=((1, .5, -.2), (.5, 1, .1), (-.2, .1, 1)))
    replay_case('replay_oruh_tdep', sdepy.ornstein_uhlenbeck_process,
                np.linspace(0, 5, 11), P, 51, x0=.1, theta=theta_t, k=1.,
                sigma=.3)
    replay_case('replay_hw3_tdep', sdepy.hull_white_process,
                np.linspace(0, 5, 11), P, 51, **HW)
    replay_case('replay_cir', sdepy.cox_ingersoll_ross_process,
                np.linspace(0, 2, 5), P, 81, x0=.05, theta=.04, k=1.5, xi=.6)
    replay_case('replay_heston', sdepy.heston_process, (0., 1.), P,
                np.linspace(0, 1, 64), **dict(HESTON, xi=.9))
    replay_case('replay_heston_full', sdepy.full_heston_process,
                np.linspace(0, 1, 5), P, 33, **HESTON)
    replay_case('replay_heston_v2', sdepy.full_heston_process,
                (0., .5, 1.), 67, 21, vshape=(2,),
                **dict(HESTON, x0=((100.,), (50.,)), rho=(-.7, .2),
                       xi=((.3,), (1.1,))))
    replay_case('replay_merton', sdepy.merton_jumpdiff_process,
                np.linspace(0, 1, 5), P, 101, jumps=True,
                **dict(MERTON, lam=20.))
    replay_case('replay_kou', sdepy.kou_jumpdiff_process,
                np.linspace(0, 1, 5), P, 101, jumps=True,
                **dict(KOU, lam=20.))
    # uneven grid: explicit step points not aligned with the timeline
    replay_case('replay_oruh_ragged', sdepy.ornstein_uhlenbeck_process,
                (0., .37, 1.), 5, (0.1, .2, .21, .5, .93, 1.5), x0=1.,
                theta=.5, k=2., sigma=.4)

    # ---- same-seed fixtures ---------------------------------------------
    seeded_case('seeded_lognorm', sdepy.lognorm_process, (0., .5, 1.), 101,
                30, 11, x0=1., mu=.05, sigma=.2)
    seeded_case('seeded_hw3_tdep', sdepy.hull_white_process, (0., 2.5, 5.),
                101, 26, 12, **HW)
    seeded_case('seeded_heston', sdepy.heston_process, (0., 1.), 101, 50, 13,
                **HESTON)
    seeded_case('seeded_merton', sdepy.merton_jumpdiff_process, (0., 1.), 101,
                80, 14, **dict(MERTON, lam=15.))
    seeded_case('seeded_kou', sdepy.kou_jumpdiff_process, (0., 1.), 101, 80,
                15, **dict(KOU, lam=15.))

    # ---- statistics -------------------------------------------------------
    rng = np.random.default_rng(21)
    x = sdepy.lognorm_process(paths=4001, steps=20, rng=rng, x0=1., mu=.05,
                              sigma=.2, vshape=(2,))((0., .5, 1.))
    mc1 = sdepy.montecarlo(x[-1], bins=25)
    a = sdepy.montecarlo(bins=25)
    chunks = (x[-1][:, :1500], x[-1][:, 1500:2700], x[-1][:, 2700:])
    for c in chunks:
        a.update(c)
    arrays = dict(x=np.asarray(x), t=x.t,
                  pmean=np.asarray(x.pmean()), pvar=np.asarray(x.pvar()),
                  pstd=np.asarray(x.pstd()), pvar1=np.asarray(x.pvar(ddof=1)))
    for tag, m in (('one', mc1), ('chunk', a)):
        arrays.update({
            tag + '_mean': m.mean(), tag + '_var': m.var(),
            tag + '_std': m.std(), tag + '_stderr': m.stderr(),
            tag + '_skew': m.skew(), tag + '_kurt': m.kurtosis(),
            tag + '_counts': np.stack([m[i].histogram()[0] for i in range(2)]),
            tag + '_edges': np.stack([m[i].histogram()[1] for i in range(2)]),
            tag + '_outpaths': np.stack([m[i].outpaths for i in range(2)]),
        })
    save('stats_lognorm', **arrays)

    # process.cdf / chf / interpolation and montecarlo pdf / cdf on the same sample
    tq = np.array([-.1, 0., .25, .5, .8, 1., 1.2])
    xq = np.linspace(.5, 2., 7)
    uq = np.linspace(-2., 2., 5)
    xs = np.linspace(.4, 2.2, 11)
    save('stats_cdf_chf', tq=tq, xq=xq, uq=uq, xs=xs,
         interp=np.asarray(x(tq)), incr=np.asarray(x(.25, .5)),
         cdf_tl=x.cdf(xq), cdf_t=x.cdf(tq, xq), chf_tl=x.chf(uq), chf_t=x.chf(tq, uq),
         mc_pdf=np.stack([mc1[i].pdf(xs) for i in range(2)]),
         mc_cdf=np.stack([mc1[i].cdf(xs) for i in range(2)]),
         mc_pdf_i=np.stack([mc1[i].pdf(xs, method='interp') for i in range(2)]),
         mc_cdf_i=np.stack([mc1[i].cdf(xs, method='interp') for i in range(2)]))

    # closed forms of sdepy.analytical (reference analytical.py) at the parameter
    # sets of tests/test_gpu_quant.py: the Philox-mode oracle values
    A = sdepy.analytical if hasattr(sdepy, 'analytical') else __import__('sdepy.analytical').analytical
    tq = np.array([.5, 1., 2., 3.])
    uq = np.linspace(-2., 2., 9)
    xq = np.linspace(-1., 2., 7)
    wp = dict(x0=1., mu=.5, sigma=.8)
    lp = dict(x0=1., mu=.05, sigma=.3)
    op = dict(x0=1., theta=.3, k=1.5, sigma=.4)
    hp = dict(x0=(.2, .1), theta=(.3, .0), k=(1., .5), sigma=(.2, .3), rho=.4)
    cp = dict(x0=.5, theta=.3, k=1.2, xi=.3)
    ep = dict(x0=1., mu=.05, sigma=1., y0=.04, theta=.05, k=1.5, xi=.3, rho=-.6)
    mp = dict(x0=1., mu=.05, sigma=.2, lam=2., a=-.1, b=.15)
    kp = dict(x0=1., mu=.05, sigma=.2, lam=2., a=.1, b=.15, pa=.4)
    save('analytical',
         tq=tq, uq=uq, xq=xq,
         wiener_mean=A.wiener_mean(tq, **wp), wiener_var=A.wiener_var(tq, **wp),
         wiener_cdf=np.stack([A.wiener_cdf(t, xq, **wp) for t in tq]),
         wiener_chf=np.stack([A.wiener_chf(t, uq, **wp) for t in tq]),
         lognorm_mean=A.lognorm_mean(tq, **lp), lognorm_std=A.lognorm_std(tq, **lp),
         lognorm_cdf=np.stack([A.lognorm_cdf(t, xq[2:] + .2, **lp) for t in tq]),
         oruh_mean=A.oruh_mean(tq, **op), oruh_var=A.oruh_var(tq, **op),
         hw2f_mean=A.hw2f_mean(tq, **hp), hw2f_var=A.hw2f_var(tq, **hp),
         cir_mean=A.cir_mean(tq, **cp), cir_var=A.cir_var(tq, **cp),
         heston_log_mean=A.heston_log_mean(tq, **ep), heston_log_var=A.heston_log_var(tq, **ep),
         heston_log_chf=np.stack([A.heston_log_chf(t, uq, **ep) for t in tq]),
         mjd_mean=A.mjd_mean(tq, **mp),
         mjd_log_chf=np.stack([A.mjd_log_chf(t, uq, **mp) for t in tq]),
         kou_mean=A.kou_mean(tq, **kp),
         kou_log_chf=np.stack([A.kou_log_chf(t, uq, **kp) for t in tq]),
         bscall=A.bscall(1.1, tq, x0=1., r=.03, sigma=.25), bsput=A.bsput(1.1, tq, x0=1., r=.03, sigma=.25))

    # the reference's own known-answer check: Euler on log x is exact for
    # constant-parameter lognormal on shared increments
    # (sdepy/tests/test_processes.py:681-708)
    t = np.linspace(0, 1, 9)
    dw = sdepy.true_wiener_source(paths=31, rng=np.random.default_rng(3))
    xe = sdepy.lognorm_process(paths=31, x0=1., mu=.05, sigma=.2, dw=dw)(t)
    w = dw(t) - dw(0.)  # realised Brownian path
    save('known_lognorm_exact', t=t, w=np.asarray(w), x=np.asarray(xe),
         dW=np.diff(np.asarray(dw(t)), axis=0))


def user_system(t, x, y, mu=0., sigma=1., xi=1.):
    """The system of the reference's SDEs docstring (integration.py:1600-1625)
    with a mean-reverting second equation."""
    return ({'dt': mu*x, 'dw': y*x}, {'dt': sigma*(1. - y), 'dw': xi*y})


def user_system_cases():
    """User-defined systems through ``integrate(q=2)``, both stacking layouts
    (addaxis False / True, integration.py:1630-1645)."""
    for name, addaxis in (('replay_system_stacked', False),
                          ('replay_system_addaxis', True)):
        cls = sdepy.integrate(q=2, sources={'dt', 'dw'}, addaxis=addaxis)(user_system)
        kw = dict(paths=53, vshape=(2, 3), steps=17, x0=(1., .3), mu=.05,
                  sigma=np.array(((.5,), (1.5,), (1.,))),
                  xi=np.array((((.2,),), ((.4,),))))
        probe = cls(rng=np.random.default_rng(5), **kw)
        rec = recorder(probe.sources['dw'])
        P = cls(dw=rec, **kw)
        tt = np.array((0., .25, 1.))
        x, y = P(tt)
        save(name, tt=tt, grid=grid_of(P, tt), dW=np.stack(rec.dz),
             out0=np.asarray(x), out1=np.asarray(y),
             p_sigma=kw['sigma'], p_xi=kw['xi'])


def time_axis_case():
    """process summaries across values and along the timeline
    (infrastructure.py:894-1122) on an irregular timeline."""
    rng = np.random.default_rng(33)
    t = np.cumsum(rng.uniform(.01, .3, size=19))
    x = rng.normal(size=(19, 2, 3, 23))*3 + 1
    p = sdepy.process(t=t, x=x)
    out = dict(t=t, x=x)
    for k in ('tmin', 'tmax', 'tsum', 'tmean', 'tvar', 'tstd', 'tcumsum',
              'tder', 'tint', 'tdiff', 'vmin', 'vmax', 'vsum', 'vmean', 'vvar',
              'vstd'):
        out[k] = np.asarray(getattr(p, k)())
    out['tvar1'] = np.asarray(p.tvar(ddof=1))
    out['vstd1'] = np.asarray(p.vstd(ddof=1))
    out['tdiff_half_bwd'] = np.asarray(p.tdiff(dt_exp=.5, fwd=False))
    out['tdiff_half_bwd_t'] = p.tdiff(dt_exp=.5, fwd=False).t
    out['tmin_t'] = p.tmin().t
    save('stats_time_axis', **out)


def jump_system(t, x=0, y=0, k=1):
    """2 equations with Poisson jumps and Wiener increments
    (reference tests/test_integrator.py:312-327)."""
    return ({'dt': x, 'dn': k, 'dw': y}, {'dt': 1, 'dn': k*y, 'dw': x - y})


def k_of_t(t):
    return .5 + .1*t


def jump_system_case():
    """User system with 'dt', 'dn' and 'dw' terms, time-dependent parameter,
    intensity and correlation.  The reference sums the three terms in the
    iteration order of a SET of ids (integration.py:1725-1729), i.e. in an
    order that depends on the interpreter's string-hash seed; the kernel sums
    in sorted-id order, so this fixture must be generated under a hash seed
    where the two coincide (the script checks and says so)."""
    cls = sdepy.integrate(q=2, sources={'dt', 'dn', 'dw'})(jump_system)
    kw = dict(paths=41, steps=30, x0=(1., .5), k=k_of_t)
    src = dict(lam=lambda t: 3. + t, rho=lambda t: -.5 + .1*t)
    probe = cls(rng=np.random.default_rng(8), **kw, **src)
    order = list(probe.A(0., np.ones((2, 41))).keys())
    if order != sorted(order):
        raise SystemExit('set order %s != sorted order: rerun with another '
                         'PYTHONHASHSEED' % order)
    rec_w, rec_n = recorder(probe.sources['dw']), recorder(probe.sources['dn'])
    P = cls(dw=rec_w, dn=rec_n, **kw)
    tt = np.linspace(0., 1., 7)
    x, y = P(tt)
    save('replay_system_jumps', tt=tt, grid=grid_of(P, tt), dW=np.stack(rec_w.dz),
         dN=np.stack(rec_n.dz), out0=np.asarray(x), out1=np.asarray(y))


CONTAINER_KEYS = {
    't_int': ('t', 2), 't_last': ('t', -1), 't_slice': ('t', slice(1, 4)),
    'p_int': ('p', 3), 'p_slice': ('p', slice(0, 5, 2)), 'p_list': ('p', [1, 3]),
    'v_int': ('v', 1), 'v_two': ('v', 1, 2), 'v_tuple': ('v', (0, 1)),
    'v_mixed': ('v', slice(None), 0)}


def process_container_case():
    """process indexing modes, rebase, shapeas, piecewise and ufunc timeline
    propagation (infrastructure.py:505-540, 635-707, 769-801, 1216-1281)."""
    rng = np.random.default_rng(44)
    t = np.linspace(0, 1, 5)
    x = rng.normal(size=(5, 2, 3, 7))
    p = sdepy.process(t, x=x)
    out = dict(t=t, x=x)
    for name, key in CONTAINER_KEYS.items():
        q = p[key]
        out[name] = np.asarray(q)
        out[name + '_t'] = q.t
    q = p.rebase((0.1, .55))
    out['rebase'], out['rebase_t'] = np.asarray(q), q.t
    out['shapeas'] = np.asarray(p['v', 0].shapeas((4, 3)))
    c = sdepy.process(c=(1., 2., 3.))
    q = c.shapeas(p['v', 0])*p['v', 0]
    out['const_times'], out['const_times_t'] = np.asarray(q), q.t
    s = np.linspace(-.2, 1.2, 57)
    out['s'] = s
    for mode in ('mid', 'forward', 'backward'):
        out['piecewise_' + mode] = sdepy.piecewise(t, v=x[..., 0], mode=mode)(s)
    save('process_container', **out)


if __name__ == '__main__':
    if sys.argv[1:] == ['container']:
        process_container_case()
    elif sys.argv[1:] == ['jumpsystem']:
        jump_system_case()
    elif sys.argv[1:] == ['systems']:
        user_system_cases()
    elif sys.argv[1:] == ['timeaxis']:
        time_axis_case()
    else:
        main()
        user_system_cases()
        time_axis_case()
        process_container_case()
        jump_system_case()
