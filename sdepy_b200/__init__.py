"""
sdepy_b200 -- B200-native (sm_100a) SDE path integration behind sdepy's API.

The integration step loop, the stochasticity sources' draws and the
process / montecarlo reductions run as hand-written CUDA kernels
(``csrc/``, C ABI in ``include/sdeb.h``); this package is the host-side mirror
of the reference's Python surface for that path.  No CPU fallback.
"""
from . import _lib                                   # fails loudly if the .so is missing
from .infrastructure import (                        # noqa: F401
    process, piecewise, device_process, montecarlo,
    source, wiener_source, poisson_source, cpoisson_source, replay_source,
    odd_wiener_source, even_poisson_source, even_cpoisson_source,
    true_source, true_wiener_source,
    norm_rv, uniform_rv, exp_rv, double_exp_rv, rvmap)
from .integration import (                           # noqa: F401
    paths_generator, integrator, SDE, SDEs, integrate, path_stats,
    wiener_SDE, wiener_process, lognorm_SDE, lognorm_process,
    ornstein_uhlenbeck_SDE, ornstein_uhlenbeck_process,
    hull_white_SDE, hull_white_process, hull_white_1factor_process,
    cox_ingersoll_ross_SDE, cox_ingersoll_ross_process,
    full_heston_SDE, full_heston_process, heston_SDE, heston_process,
    jumpdiff_SDE, jumpdiff_process,
    merton_jumpdiff_SDE, merton_jumpdiff_process,
    kou_jumpdiff_SDE, kou_jumpdiff_process)
from .kfun import kfunc, iskfunc                     # noqa: F401

# interactive shortcuts, wrapped as kfuncs (reference shortcuts.py:73-99 with
# the default _config.KFUNC = 'shortcuts': full names stay plain classes)
dw, dn, dj = kfunc(wiener_source), kfunc(poisson_source), kfunc(cpoisson_source)
odd_dw, even_dn, even_dj = (kfunc(odd_wiener_source), kfunc(even_poisson_source),
                            kfunc(even_cpoisson_source))
true_dw = kfunc(true_wiener_source)
wiener, lognorm = kfunc(wiener_process), kfunc(lognorm_process)
oruh, cir = kfunc(ornstein_uhlenbeck_process), kfunc(cox_ingersoll_ross_process)
hwff, hw1f = kfunc(hull_white_process), kfunc(hull_white_1factor_process)
heston_xy, heston = kfunc(full_heston_process), kfunc(heston_process)
jumpdiff = kfunc(jumpdiff_process)
mjd, kou = kfunc(merton_jumpdiff_process), kfunc(kou_jumpdiff_process)

__version__ = '0.1.0'
