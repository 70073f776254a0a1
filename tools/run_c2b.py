"""One C2b run (Hull-White 3 factors, time-dependent theta and correlation, full path) for ncu."""
import sys
import numpy as np
import torch
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import sdepy_b200 as sd  # noqa: E402


def hw_theta(t):
    return np.array(((.02 + .001*t,), (0.,), (0.,)))


def hw_corr(t):
    c01, c02, c12 = .3*np.cos(t), -.2 + .05*t, .1
    return np.array(((1, c01, c02), (c01, 1, c12), (c02, c12, 1)))


p = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
tl = np.linspace(0., 5., 501)
for _ in range(2):
    x = sd.hull_white_process(factors=3, x0=((.01,), (0.,), (0.,)), theta=hw_theta,
                              k=((.1,), (.5,), (1.,)), sigma=((.01,), (.008,), (.005,)),
                              corr=hw_corr, paths=p, seed=3, output='device', getinfo=False)(tl)
    torch.cuda.synchronize()
print(x.shape)
