"""Shared test-case definitions (the time-dependent callables used when the
golden fixtures were generated, see tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def theta_t(t):
    return .2 + .1*t


def hw_theta(t):
    return np.array(((.02 + .001*t,), (0.,), (0.,)))


def hw_corr(t):
    c01, c02, c12 = .3*np.cos(t), -.2 + .05*t, .1
    return np.array(((1, c01, c02), (c01, 1, c12), (c02, c12, 1)))


def user_system(t, x, y, mu=0., sigma=1., xi=1.):
    """Same equations as tests/golden/make_golden.py:user_system."""
    return ({'dt': mu*x, 'dw': y*x}, {'dt': sigma*(1. - y), 'dw': xi*y})


def jump_system(t, x=0, y=0, k=1):
    """Same equations as tests/golden/make_golden.py:jump_system."""
    return ({'dt': x, 'dn': k, 'dw': y}, {'dt': 1, 'dn': k*y, 'dw': x - y})


def k_of_t(t):
    return .5 + .1*t


HW = dict(factors=3, x0=((.01,), (0.,), (0.,)), theta=hw_theta,
          k=((.1,), (.5,), (1.,)), sigma=((.01,), (.008,), (.005,)),
          corr=hw_corr)
HESTON = dict(x0=100., mu=.03, sigma=1., y0=.04, theta=.04, k=2., xi=.3,
              rho=-.7)
MERTON = dict(x0=1., mu=.05, sigma=.2, lam=2., a=-.1, b=.15)
KOU = dict(x0=1., mu=.05, sigma=.2, lam=2., a=.1, b=.15, pa=.4)

# name -> (oracle model, params for the sde, extra kwargs) for replay fixtures
REPLAY = {
    'replay_wiener': ('wiener', dict(mu=.1, sigma=.7), dict(x0=.5)),
    'replay_lognorm': ('lognorm', dict(mu=.05, sigma=.2), dict(x0=1.)),
    'replay_lognorm_v3': ('lognorm',
                          dict(mu=((.05,), (.0,), (-.1,)),
                               sigma=((.2,), (.3,), (.1,))),
                          dict(x0=((1.,), (2.,), (3.,)))),
    'replay_oruh_tdep': ('ornstein_uhlenbeck',
                         dict(theta=theta_t, k=1., sigma=.3), dict(x0=.1)),
    'replay_hw3_tdep': ('hull_white',
                        dict(theta=hw_theta, k=HW['k'], sigma=HW['sigma']),
                        dict(x0=HW['x0'])),
    'replay_cir': ('cox_ingersoll_ross', dict(theta=.04, k=1.5, xi=.6),
                   dict(x0=.05)),
    'replay_heston': ('heston',
                      dict(mu=.03, sigma=1., theta=.04, k=2., xi=.9),
                      dict(x0=100., y0=.04)),
    'replay_heston_full': ('heston',
                           dict(mu=.03, sigma=1., theta=.04, k=2., xi=.3),
                           dict(x0=100., y0=.04, full=True)),
    'replay_heston_v2': ('heston',
                         dict(mu=.03, sigma=1., theta=.04, k=2.,
                              xi=((.3,), (1.1,))),
                         dict(x0=((100.,), (50.,)), y0=.04, full=True)),
    'replay_merton': ('jumpdiff', dict(mu=.05, sigma=.2), dict(x0=1.)),
    'replay_kou': ('jumpdiff', dict(mu=.05, sigma=.2), dict(x0=1.)),
    'replay_oruh_ragged': ('ornstein_uhlenbeck',
                           dict(theta=.5, k=2., sigma=.4), dict(x0=1.)),
}


# --------------------------------------------------------------------------
# user plug points let / info_begin / info_next / info_end on user-defined
# classes (reference integration.py:1154-1197, 1501-1581).  `m` is the package
# providing SDE / SDEs / integrator / process: the reference when the golden
# fixture is generated (tests/golden/make_user_hooks.py), sdepy_b200 in the tests.
# --------------------------------------------------------------------------

def user_hooks_single(m):
    class A_SDE(m.SDE):
        def sde(self, t, x, k=1., s=.4):
            return {'dt': -k*x, 'dw': s}

        def let(self, t, out_x, x):
            out_x[...] = x*x + 1.

        def info_begin(self):
            self.info['neg'] = np.zeros(self.vshape + (self.paths,), dtype=int)
            self.info['big'] = np.zeros(self.vshape + (self.paths,), dtype=int)

        def info_next(self):
            iv = self.itervars
            self.info['neg'] += (iv['last_x'] < 0)
            self.info['big'] += (iv['new_x'] > .25)

        def info_end(self):
            self.info['total_neg'] = int(self.info['neg'].sum())

    class A(A_SDE, m.integrator):
        pass
    return A


def user_hooks_system(m):
    class B_SDE(m.SDEs):
        q = 2

        def sde(self, t, x, y, a=.5):
            return ({'dt': -a*x, 'dw': y}, {'dt': a*(1 - y), 'dw': .2*y})

        def shapes(self, vshape):
            vshape, xshape, wshape = super().shapes(vshape)
            return vshape, vshape, wshape

        def let(self, t, out_x, X):
            x, y = self.unpack(X)
            out_x[...] = x*y

        def result(self, tt, xx):
            return m.process(tt, x=xx)

        def info_begin(self):
            self.info['ylow'] = np.zeros(self.vshape + (self.paths,), dtype=int)

        def info_next(self):
            x, y = self.unpack(self.itervars['last_x'])
            self.info['ylow'] += (y < .9)

    class B(B_SDE, m.integrator):
        pass
    return B


def two_jump_terms(t, x, a=.5, b=.3, c=.2):
    """One equation with a Poisson ('dn') AND a compound Poisson ('dj') term
    next to drift and diffusion: two independent jump sources."""
    return {'dt': -a*x, 'dw': b, 'dn': c*x, 'dj': 1.}
