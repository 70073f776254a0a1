"""
sdepy_b200 -- B200-native (sm_100a) SDE path integration behind sdepy's API.

The integration step loop, the stochasticity sources' draws and the
process / montecarlo reductions run as hand-written CUDA kernels
(``csrc/``, C ABI in ``include/sdeb.h``); this package is the host-side mirror
of the reference's Python surface for that path.  No CPU fallback.
"""
from . import _lib                                   # fails loudly if the .so is missing
from .infrastructure import (                        # noqa: F401
    process, device_process, montecarlo,
    source, wiener_source, poisson_source, cpoisson_source, replay_source,
    odd_wiener_source, even_poisson_source, even_cpoisson_source, true_wiener_source,
    norm_rv, uniform_rv, exp_rv, double_exp_rv)
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

__version__ = '0.1.0'
