"""Kernel seconds of C2b (Hull-White 3 factors, time-dependent theta and correlation, full path)."""
import os as _os
import sys
sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sdepy_b200 as sd  # noqa: E402
from sdepy_b200 import _lib  # noqa: E402
from tools.run_mode import hw_theta, hw_corr  # noqa: E402

events = []
real = _lib.lib.sdeb_integrate


def timed(p, stream):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = real(p, stream); e1.record()
    events.append((e0, e1))
    return rc


_lib.lib.sdeb_integrate = timed
tl = np.linspace(0., 5., 501)
best = None
for it in range(5):
    del events[:]
    x = sd.hull_white_process(factors=3, x0=((.01,), (0.,), (0.,)), theta=hw_theta,
                              k=((.1,), (.5,), (1.,)), sigma=((.01,), (.008,), (.005,)),
                              corr=hw_corr, paths=1_000_000, seed=3, output='device',
                              getinfo=False)(tl)
    torch.cuda.synchronize()
    t = sum(a.elapsed_time(b) for a, b in events)*1e-3
    if it and (best is None or t < best):
        best = t
print(_os.environ.get('SDEB_LIB') or 'product', 'C2b %.4f ms  %.4e path-steps/s' % (best*1e3, 5e8/best))
