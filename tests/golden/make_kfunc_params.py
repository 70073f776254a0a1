"""Golden `params` dictionaries of the reference's kfunc-wrapped sources and
processes (sdepy/kfun.py:171-186; the reference's tests/test_kfunc.py:255-400
pins the same facts).  Run in the build container, where the reference imports:

    PYTHONPATH=/root/reference python tests/golden/make_kfunc_params.py
"""
import json
import os

import numpy as np
import sdepy

CASES = [
    ('wiener_source', dict(paths=2)), ('odd_wiener_source', dict(paths=2)),
    ('true_wiener_source', dict(paths=2)), ('poisson_source', dict(lam=2)),
    ('cpoisson_source', dict(paths=2, lam=3.)),
    ('wiener_process', dict(x0=2)), ('lognorm_process', dict(x0=2)),
    ('ornstein_uhlenbeck_process', dict(theta=2, steps=20)),
    ('hull_white_process', dict(vshape=3, factors=2)),
    ('hull_white_1factor_process', dict(vshape=3)),
    ('cox_ingersoll_ross_process', dict(x0=2)),
    ('full_heston_process', dict(x0=2)), ('heston_process', dict(x0=2)),
    ('jumpdiff_process', dict(x0=2)), ('merton_jumpdiff_process', dict(x0=2)),
    ('kou_jumpdiff_process', dict(x0=2)),
]


def plain(v):
    if isinstance(v, (np.integer, np.floating)):
        return v.item()
    if isinstance(v, np.ndarray):
        return {'__array__': v.tolist()}
    if isinstance(v, tuple):
        return {'__tuple__': [plain(z) for z in v]}
    if v is None or isinstance(v, (bool, int, float, str)):
        return v
    return {'__repr__': type(v).__name__}     # rv objects, dtypes: only the kind


out = {}
for name, kw in CASES:
    cls = getattr(sdepy, name)
    K = cls if sdepy.iskfunc(cls) else sdepy.kfunc(cls)
    out[name] = {'kw': kw, 'params': {k: plain(v) for k, v in K(**kw).params.items()}}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'kfunc_params.json'), 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)
print('wrote', len(out), 'cases')
