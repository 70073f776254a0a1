"""One secondary configuration, twice (warm-up + the launch ncu captures):
    python tools/run_mode.py replay_ou | c2a | c2b | mc | tsum | c4 | c5
"""
import os as _os
import sys
sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sdepy_b200 as sd  # noqa: E402


def hw_theta(t):
    return np.array(((.02 + .001*t,), (0.,), (0.,)))


def hw_corr(t):
    c01, c02, c12 = .3*np.cos(t), -.2 + .05*t, .1
    return np.array(((1, c01, c02), (c01, 1, c12), (c02, c12, 1)))


def main():
    mode = sys.argv[1]
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.
    tl = np.linspace(0., 5., 501)
    p = int(1_000_000*scale)
    if mode == 'replay_ou':
        pr, nr = int(4_000_000*scale), 250
        dW = torch.randn((nr, pr), dtype=torch.float64, device='cuda')*np.sqrt(1/nr)
        fn = lambda: sd.ornstein_uhlenbeck_process(
            x0=.1, theta=.2, k=1., sigma=.3, paths=pr, dw=sd.replay_source(dW),
            output='device')(np.linspace(0., 1., nr + 1))
    elif mode == 'c2a':
        fn = lambda: sd.ornstein_uhlenbeck_process(
            x0=.1, theta=lambda s: .2 + .1*s, k=1., sigma=.3, paths=p, seed=2,
            output='device')(tl)
    elif mode == 'c2b':
        fn = lambda: sd.hull_white_process(
            factors=3, x0=((.01,), (0.,), (0.,)), theta=hw_theta, k=((.1,), (.5,), (1.,)),
            sigma=((.01,), (.008,), (.005,)), corr=hw_corr, paths=p, seed=3,
            output='device', getinfo=False)(tl)
    elif mode == 'mc':
        x = torch.empty(int(100_000_000*scale), dtype=torch.float64, device='cuda').normal_()
        fn = lambda: sd.montecarlo(x, bins=100)
    elif mode == 'tsum':
        proc = sd.ornstein_uhlenbeck_process(x0=.1, theta=.2, k=1., sigma=.3, paths=p, seed=8,
                                             output='device')(tl)
        fn = lambda: (proc.tmean(), proc.tcumsum())
    elif mode == 'c4':
        fn = lambda: sd.merton_jumpdiff_process(
            x0=1., mu=.05, sigma=.2, lam=2., a=-.1, b=.15, paths=int(10_000_000*scale),
            steps=1001, seed=5, output='stats', getinfo=False)((0., 1.))
    elif mode == 'c5':
        @sd.integrate
        def gbm(t, x, mu=.05, sigma=.2):
            return {'dt': mu*x, 'dw': sigma*x}
        fn = lambda: gbm(paths=int(20_000_000*scale), steps=2001, x0=1., method='milstein',
                         seed=7, output='device', getinfo=False)((0., 1.))
    else:
        raise SystemExit('unknown mode ' + mode)
    for _ in range(2):
        r = fn()
        torch.cuda.synchronize()
    print(mode, 'ok')


if __name__ == '__main__':
    main()
