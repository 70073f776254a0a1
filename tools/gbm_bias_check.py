"""Weak-bias check of the Philox normals: GBM terminal mean, Euler and Milstein, 200 steps.
E[x_N] = (1 + mu*dt)^N exactly for both schemes when E[dw] = 0 and E[dw^2] = dt, so any
deviation beyond the standard error exposes a bias in the first two moments of the draws
(the mean of dw is amplified by N*sigma*sqrt(dt) = 2.8)."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import sdepy_b200 as sd  # noqa: E402


@sd.integrate
def gbm(t, x, mu=.05, sigma=.2):
    return {'dt': mu*x, 'dw': sigma*x}


want = (1 + .05/200)**200
for method in ('euler', 'milstein'):
    tot, tot2, n = 0., 0., 0
    for seed in range(1, 5):
        st = gbm(paths=100_000_000, steps=201, x0=1., method=method, seed=seed,
                 output='stats', getinfo=False)((0., 1.))
        m, se = float(np.asarray(st.pmean())[-1, 0]), float(np.asarray(st.stderr())[-1, 0])
        print('%-8s seed %d: mean %.7f  (want %.7f)  dev %+.2f sigma' % (method, seed, m, want, (m - want)/se))
        tot += m; n += 1; tot2 += se*se
    print('%-8s pooled: dev %+.2f sigma of %.2e' % (method, (tot/n - want)/np.sqrt(tot2)*n, np.sqrt(tot2)/n))
st = sd.lognorm_process(paths=400_000_000, steps=201, x0=1., mu=.05, sigma=.2, seed=9,
                        output='stats', getinfo=False)((0., 1.))
m, se = float(np.asarray(st.pmean())[-1, 0]), float(np.asarray(st.stderr())[-1, 0])
print('lognorm preset (exact in log space): mean %.7f want %.7f dev %+.2f sigma' % (m, np.exp(.05), (m - np.exp(.05))/se))
v = float(np.asarray(st.pvar())[-1, 0]); wantv = np.exp(.1)*(np.exp(.04) - 1)
print('  var %.7f want %.7f rel %.2e (stat. resolution %.1e)' % (v, wantv, v/wantv - 1, np.sqrt(2/4e8)*3))
