"""Host-side cost of one public-API call (Heston C3, output='stats'): cProfile of 20 calls."""
import cProfile
import os as _os
import pstats
import sys
sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sdepy_b200 as sd  # noqa: E402

paths = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
grid = np.linspace(0., 1., 253)
kw = dict(x0=100., mu=.03, sigma=1., y0=.04, theta=.04, k=2., xi=.3)


def call(seed):
    r = sd.heston_process(paths=paths, steps=grid, rho=-.7, seed=seed, output='stats',
                          payoff=('call', 100., float(np.exp(-.03))), getinfo=True, **kw)((0., 1.))
    return float(np.asarray(r.payoff_mean())[-1, 0])


for i in range(3):
    call(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    call(10 + i)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(35)
