"""Secondary measurements: the other BASELINE.json configs (C1, C2, C4, C5) and the
HBM-bound modes (full-path stores, replay reads).  Device-timed with CUDA events
through the public API (tables are a few KB; their H2D is included).

usage (GPU box): PYTHONPATH=. python tools/bench_modes.py [--scale 1.0] > gpurun_out/modes.json
"""
import argparse
import json
import sys
import time

import numpy as np
import torch

import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import sdepy_b200 as sd  # noqa: E402

HBM_GBS = 6553.6      # MEASURED_PEAKS.json (copy bandwidth, read+write)


from sdepy_b200 import _lib

KERNEL_EVENTS = []
_real_integrate = _lib.lib.sdeb_integrate


def _timed_integrate(p, stream):
    """sdeb_integrate bracketed by CUDA events on the launching stream, so that
    the kernel time is separated from the host-side lowering."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = _real_integrate(p, stream)
    e1.record()
    KERNEL_EVENTS.append((e0, e1))
    return rc


_lib.lib.sdeb_integrate = _timed_integrate


def timed(fn, reps=3, warm=1):
    """(best end-to-end seconds, best kernel-only seconds, last result)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts, ks = [], []
    for _ in range(reps):
        del KERNEL_EVENTS[:]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)*1e-3)
        ks.append(sum(a.elapsed_time(b) for a, b in KERNEL_EVENTS)*1e-3)
    return (min(ts), min(ks)), out


def hw_theta(t):
    return np.array(((.02 + .001*t,), (0.,), (0.,)))


def hw_corr(t):
    c01, c02, c12 = .3*np.cos(t), -.2 + .05*t, .1
    return np.array(((1, c01, c02), (c01, 1, c12), (c02, c12, 1)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scale', type=float, default=1.0)
    a = ap.parse_args()
    res = []

    def rec(name, paths, nsteps, tk, stored=0, read=0, note=''):
        te, t = tk
        r = dict(config=name, paths=paths, steps=nsteps, seconds_api=te, seconds_kernel=t,
                 path_steps_per_s_kernel=paths*nsteps/t, path_steps_per_s_api=paths*nsteps/te)
        if stored or read:
            r['hbm_bytes'] = 8*(stored + read)
            r['hbm_gbs_kernel'] = 8*(stored + read)/t/1e9
            r['hbm_frac_of_measured_copy_peak'] = r['hbm_gbs_kernel']/HBM_GBS
        r['note'] = note
        res.append(r)
        print(json.dumps(r), file=sys.stderr)

    # C1: lognorm 1e5 x 250, full path on device, pmean/pstd
    p, n = 100_000, 250
    t, x = timed(lambda: sd.lognorm_process(x0=1., mu=.05, sigma=.2, paths=p, seed=1,
                                            output='device')(np.linspace(0, 1, n + 1)))
    rec('C1 lognorm full path', p, n, t, stored=p*(n + 1))

    # C2a: OU, time-dependent theta, 1e6 x 500, full path
    p, n = int(1_000_000*a.scale), 500
    tl = np.linspace(0, 5, n + 1)
    t, x = timed(lambda: sd.ornstein_uhlenbeck_process(
        x0=.1, theta=lambda s: .2 + .1*s, k=1., sigma=.3, paths=p, seed=2,
        output='device')(tl))
    rec('C2a OU tdep full path (philox)', p, n, t, stored=p*(n + 1))
    del x
    # C2b: Hull-White 3 factors, time-dependent 3x3 correlation, full path
    t, x = timed(lambda: sd.hull_white_process(
        factors=3, x0=((.01,), (0.,), (0.,)), theta=hw_theta, k=((.1,), (.5,), (1.,)),
        sigma=((.01,), (.008,), (.005,)), corr=hw_corr, paths=p, seed=3,
        output='device')(tl))
    rec('C2b HW3 tdep corr full path (philox)', p, n, t, stored=p*(n + 1))
    del x
    # replay, full path: read 8 B + write 8 B per path-step (lognorm)
    p, n = int(4_000_000*a.scale), 250
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    dW = torch.randn((n, p), dtype=torch.float64, device='cuda', generator=g)*np.sqrt(1/n)
    tl = np.linspace(0, 1, n + 1)
    t, x = timed(lambda: sd.lognorm_process(x0=1., mu=.05, sigma=.2, paths=p,
                                            dw=sd.replay_source(dW), output='device')(tl))
    rec('replay lognorm full path', p, n, t, stored=p*(n + 1), read=p*n)
    del x
    t, x = timed(lambda: sd.lognorm_process(x0=1., mu=.05, sigma=.2, paths=p,
                                            dw=sd.replay_source(dW), steps=n + 1,
                                            output='device')((0., 1.)))
    rec('replay lognorm terminal only', p, n, t, stored=2*p, read=p*n)
    del x
    # the drop-in default: host `process` result (pinned D2H of the 4 GB slab included)
    p_, n_ = int(1_000_000*a.scale), 500
    tl_ = np.linspace(0., 5., n_ + 1)
    run = lambda: sd.ornstein_uhlenbeck_process(
        x0=.1, theta=.2, k=1., sigma=.3, paths=p_, seed=2)(tl_)
    run()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); xh = run(); ts.append(time.perf_counter() - t0)
    res.append(dict(config="OU full path to a host process (output='process', default)",
                    paths=p_, steps=n_, seconds_wall=min(ts),
                    path_steps_per_s_wall=p_*n_/min(ts), d2h_GBps=8.*p_*(n_ + 1)/min(ts)/1e9))
    del xh
    # non-log model: no exp at the store, pure stream (read 8 B + write 8 B per path-step)
    t, x = timed(lambda: sd.ornstein_uhlenbeck_process(
        x0=.1, theta=.2, k=1., sigma=.3, paths=p, dw=sd.replay_source(dW),
        output='device')(tl))
    rec('replay OU full path', p, n, t, stored=p*(n + 1), read=p*n)
    del x, dW
    # philox lognorm terminal (cheapest model: RNG-bound)
    p, n = int(100_000_000*a.scale), 250
    t, x = timed(lambda: sd.lognorm_process(x0=1., mu=.05, sigma=.2, paths=p, steps=n + 1,
                                            seed=4, output='stats')((0., 1.)))
    rec('lognorm terminal stats (philox)', p, n, t)
    # C4: Merton / Kou 1e7 x 1000, terminal
    p, n = int(10_000_000*a.scale), 1000
    t, x = timed(lambda: sd.merton_jumpdiff_process(
        x0=1., mu=.05, sigma=.2, lam=2., a=-.1, b=.15, paths=p, steps=n + 1, seed=5,
        output='stats', getinfo=False)((0., 1.)))
    rec('C4 merton terminal stats (philox)', p, n, t)
    t, x = timed(lambda: sd.kou_jumpdiff_process(
        x0=1., mu=.05, sigma=.2, lam=2., a=.1, b=.15, pa=.4, paths=p, steps=n + 1, seed=6,
        output='stats', getinfo=False)((0., 1.)))
    rec('C4 kou terminal stats (philox)', p, n, t)
    t, x = timed(lambda: sd.merton_jumpdiff_process(
        x0=1., mu=.05, sigma=.2, lam=2., a=-.1, b=.15, paths=p, steps=n + 1, seed=5,
        output='stats')((0., 1.)))
    rec('C4 merton terminal stats (philox, getinfo: jump_count + jump_rate)', p, n, t)
    # C4 in REPLAY mode at its own step count: dW + dJ + dN tables streamed from HBM
    # (24 B read per path-step), terminal value only
    p, n = int(1_000_000*a.scale), 1000
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    dW = torch.randn((n, p), dtype=torch.float64, device='cuda', generator=g)*np.sqrt(1/n)
    dN = torch.poisson(torch.full((n, p), 2./n, dtype=torch.float64, device='cuda'), generator=g).to(torch.int64)
    dJ = dN.double()*(-.1) + dN.double().sqrt()*.15*torch.randn((n, p), dtype=torch.float64, device='cuda', generator=g)
    t, x = timed(lambda: sd.merton_jumpdiff_process(
        x0=1., mu=.05, sigma=.2, lam=2., a=-.1, b=.15, paths=p, steps=n + 1,
        dw=sd.replay_source(dW), dj=sd.replay_source(dJ, dn=dN), output='device')((0., 1.)))
    rec('C4 merton replay (dW + dJ + dN tables), terminal only', p, n, t, stored=2*p, read=3*p*n)
    del x, dW, dJ, dN
    # summaries of a resident slab along the timeline (path-dependent payoffs)
    p, n = int(1_000_000*a.scale), 500
    proc = sd.ornstein_uhlenbeck_process(x0=.1, theta=.2, k=1., sigma=.3, paths=p, seed=8,
                                         output='device')(np.linspace(0., 5., n + 1))
    for name, fn, passes_r, passes_w in (('tmax', proc.tmax, 1, 0), ('tmean', proc.tmean, 1, 0),
                                         ('tstd', proc.tstd, 2, 0), ('tcumsum', proc.tcumsum, 1, 1),
                                         ('tdiff', proc.tdiff, 1, 1), ('tint', proc.tint, 1, 1)):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1)*1e-3)
            del r
        nbytes = 8.*(n + 1)*p*(passes_r + passes_w)
        res.append(dict(config='device_process.%s on (%d, %d)' % (name, n + 1, p),
                        seconds=min(ts), GBps=nbytes/min(ts)/1e9,
                        frac_of_copy_peak=nbytes/min(ts)/1e9/HBM_GBS))
    del proc
    # C5: custom @integrate, Milstein, 1e8 x 2000, montecarlo of the terminal value
    @sd.integrate
    def gbm(t, x, mu=.05, sigma=.2):
        return {'dt': mu*x, 'dw': sigma*x}
    p, n = int(100_000_000*a.scale), 2000
    P = gbm(paths=p, steps=n + 1, x0=1., method='milstein', seed=7, output='device',
            getinfo=False)
    P((0., 1.)); torch.cuda.synchronize()           # NVRTC compile outside the timing
    t, x = timed(lambda: gbm(paths=p, steps=n + 1, x0=1., method='milstein', seed=7,
                             output='device', getinfo=False)((0., 1.)), reps=2, warm=0)
    rec('C5 traced GBM Milstein terminal (philox, NVRTC)', p, n, t, stored=2*p)
    t0 = time.perf_counter()
    mc = sd.montecarlo(x.x[-1], bins=100)
    torch.cuda.synchronize()
    res.append(dict(config='C5 montecarlo(1e8 samples, 100 bins)', seconds_api=time.perf_counter() - t0,
                    mean=float(mc.mean()), stderr=float(mc.stderr()), expected_mean=float(np.exp(.05))))
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
