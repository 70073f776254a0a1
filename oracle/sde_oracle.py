"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the sdepy hot path (SDE Euler step loop, stochasticity
sources, process/montecarlo reductions).  It exists to CHECK the CUDA product
in ``sdepy_b200``; nothing under ``sdepy_b200/`` may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline /
``--impl reference`` legs are allowed to call into this module.

Parity status: PINNED.  ``tests/golden/make_golden.py`` ran the unmodified
reference (``/root/reference/sdepy`` 1.2.1-dev0, NumPy 2.3.5, SciPy 1.18.1)
and committed its inputs/outputs under ``tests/golden``;
``tests/test_oracle_golden.py`` asserts this module reproduces every one of
them BIT-EXACTLY (replay mode: same pre-drawn increments; self-driven mode:
same ``numpy.random`` generator and seed).  The Milstein scheme, which the
reference does not ship, is pinned to the reference's integrator machinery
running it as a plug-in through the ``method=`` hook (see ``milstein`` below).

Third-party arithmetic on the reference path that is not under
/root/reference: ``numpy.random.Generator.normal / multivariate_normal /
poisson`` and ``scipy.stats.<dist>.rvs`` (reference call sites
``sdepy/infrastructure.py:1515, 1532, 1548, 1631, 2035``; the reference pins
only ``numpy>=1.15.2, scipy>=0.19.1``, ``setup.py:48-49``).  The self-driven
mode below calls the very same library entry points in the very same order,
so it consumes the generator stream identically.

Every function cites the reference lines it restates (paths relative to
/root/reference).  The code is organised functionally (a model is a pure
``coefficients(x, p)`` function; one generic driver walks the step grid)
rather than as the reference's cooperating class hierarchy.
"""
import math
import numpy as np

try:  # only needed by the self-driven compound-Poisson draws
    import scipy.stats as _st
except Exception:  # pragma: no cover
    _st = None


# --------------------------------------------------------------------------
# step grid                                   sdepy/integration.py:189-218, 506-538
# --------------------------------------------------------------------------

def step_grid(timeline, steps=None):
    """Output timeline ``tt`` -> merged integration grid.

    Restates ``paths_generator.pace`` (integration.py:209-218) and the merge
    in ``paths_generator.__call__`` (integration.py:529-538): an integer
    ``steps`` is the number of *points* of ``linspace(t0, t1, steps)``, an
    iterable is taken as explicit points, points outside [t0, t1] are
    dropped and the union with the timeline is sorted/deduplicated by
    ``np.unique``.  Returns ``(tt, grid, where)`` with ``grid[where] == tt``.
    """
    tt = np.asarray(timeline)
    if tt.dtype.kind == 'i':
        tt = tt.astype(float)
    if tt.size == 1:
        target = tt
    elif steps is None:
        target = np.array((), dtype=tt.dtype)
    elif np.isscalar(steps):
        target = np.linspace(tt[0], tt[-1], steps, dtype=tt.dtype)
    else:
        target = np.fromiter(steps, dtype=tt.dtype)
    target = target[(target >= tt[0]) & (target <= tt[-1])]
    grid = np.unique(np.concatenate((target, tt)))
    where = np.searchsorted(grid, tt)
    assert np.array_equal(grid[where], tt)
    return tt, grid, where


# --------------------------------------------------------------------------
# models: coefficient functions in the reference's exact operation order
# --------------------------------------------------------------------------
# Each returns a list of (coefficient, differential-id) in the order of the
# reference's ``sde`` dict, so that the Euler sum below associates the same
# way as ``sum(A.get(id, 0)*dZ[id] for id in A.keys())`` (integration.py:718).

def _wiener(x, p):            # integration.py:2069-2070
    return [(p['mu'], 'dt'), (p['sigma'], 'dw')]


def _lognorm(x, p):           # integration.py:2129-2130 (on a = log x)
    return [(p['mu'] - p['sigma']*p['sigma']/2, 'dt'), (p['sigma'], 'dw')]


def _oruh(x, p):              # integration.py:2198-2199, 2268-2269
    return [(p['k']*(p['theta'] - x), 'dt'), (p['sigma'], 'dw')]


def _cir(x, p):               # integration.py:2351-2354
    xp = np.maximum(x, 0.)
    return [(p['k']*(p['theta'] - xp), 'dt'), (p['xi']*np.sqrt(xp), 'dw')]


def _jumpdiff(x, p):          # integration.py:2594-2608 (no martingale corr.)
    return [(p['mu'] - p['sigma']*p['sigma']/2, 'dt'), (p['sigma'], 'dw'),
            (1, 'dj')]


def _heston_pair(x, y, p):    # integration.py:2425-2433
    yp = np.maximum(y, 0.)
    return ([(p['mu'] - p['sigma']*p['sigma']*yp/2, 'dt'),
             (p['sigma']*np.sqrt(yp), 'dw')],
            [(p['k']*(p['theta'] - yp), 'dt'),
             (p['xi']*np.sqrt(yp), 'dw')])


SCALAR_MODELS = {
    'wiener': (_wiener, False), 'lognorm': (_lognorm, True),
    'ornstein_uhlenbeck': (_oruh, False), 'hull_white': (_oruh, False),
    'cox_ingersoll_ross': (_cir, False), 'jumpdiff': (_jumpdiff, True),
}


def _at(p, t):
    """Parameters may be callables of time (evaluated at the LEFT endpoint of
    the step, integration.py:1218-1221)."""
    return {k: (np.asarray(v(t)) if callable(v) else np.asarray(v))
            for k, v in p.items()}


def _euler_sum(x, terms, dz):
    # integration.py:718 -- python ``sum`` starts from int 0 and adds the
    # separately rounded products left to right.
    acc = 0
    for coeff, ident in terms:
        acc = acc + coeff*dz[ident]
    return x + acc


# --------------------------------------------------------------------------
# generic driver
# --------------------------------------------------------------------------

def euler_replay(model, params, x0, grid, where, dW, dJ=None, dN=None,
                 factors=None, y0=None, full=False, scheme='euler',
                 diffusion_dx=None):
    """Integrate ``model`` over ``grid`` consuming pre-drawn increments.

    dW[n] (and dJ[n], dN[n]) are the increments the reference's sources
    returned for step n, i.e. already scaled by sqrt|dt| (infrastructure.py:
    1558-1559).  Shapes follow the reference working shape ``wshape+(paths,)``.

    Restates the loop ``paths_generator._generate_paths``
    (integration.py:392-474) + ``integrator.euler_next`` (707-723) +
    ``SDE.begin/store/exit`` (1154-1197): initial condition (``log`` for
    log-processes, 1163-1164), one Euler update per grid interval, store at
    the grid points that belong to the output timeline, ``exp`` at exit
    (1195-1196; Heston x-component by hand 2443, 2540).

    Returns ``(xx, info)``; ``xx`` has shape ``(len(where),)+xshape+(paths,)``
    (a tuple ``(xx, yy)`` for ``model='heston'`` with ``full=True``).
    """
    grid = np.asarray(grid, dtype=float)
    nsteps = grid.size - 1
    dW = np.asarray(dW)
    paths = dW.shape[-1]
    wshape = dW.shape[1:-1]
    info = {}
    out_of = {int(j): i for i, j in enumerate(where)}

    if model == 'heston':
        # integration.py:2416-2419: state = (log x0, y0) stacked on axis -2
        half = wshape[-1]//2
        vshape_p = wshape[:-1] + (half,) + (paths,)
        a = np.empty(vshape_p)
        a[...] = np.log(np.asarray(x0))
        y = np.empty(vshape_p)
        y[...] = np.asarray(y0)
        neg = np.zeros(vshape_p, dtype=np.int64)      # 2421-2423
        xs, ys = [], []
        if 0 in out_of:
            xs.append(a.copy()); ys.append(y.copy())
        for n in range(nsteps):
            s, ds = grid[n], grid[n+1] - grid[n]
            p = _at(params, s)
            tx, ty = _heston_pair(a, y, p)
            neg += (y < 0)                            # 2435-2439 (pre-step y)
            dwx, dwy = dW[n][..., :half, :], dW[n][..., half:, :]
            a, y = (_euler_sum(a, tx, {'dt': ds, 'dw': dwx}),
                    _euler_sum(y, ty, {'dt': ds, 'dw': dwy}))
            if n + 1 in out_of:
                xs.append(a.copy()); ys.append(y.copy())
        info['negative_y_count'] = neg
        xx = np.exp(np.stack(xs))                     # 2443 / 2540
        if wshape == (2,):                            # vshape == () (addaxis)
            xx = xx[:, 0]
            neg.shape = (paths,)
        if full:
            yy = np.stack(ys)
            if wshape == (2,):
                yy = yy[:, 0]
            return (xx, yy), info
        return xx, info

    coeffs, is_log = SCALAR_MODELS[model]
    x = np.empty(wshape + (paths,))
    x[...] = np.asarray(x0)                           # integration.py:1469
    if is_log:
        x = np.log(x)                                 # integration.py:1163-1164
    has_jumps = model == 'jumpdiff'
    if has_jumps:
        info['jump_count'] = np.zeros(wshape + (paths,), dtype=np.int64)

    def emit(x):
        # hull_white_SDE.let sums the factor axis (integration.py:2271-2272)
        return x.sum(axis=-2) if model == 'hull_white' else x.copy()

    rows = []
    if 0 in out_of:
        rows.append(emit(x))
    for n in range(nsteps):
        s, ds = grid[n], grid[n+1] - grid[n]
        p = _at(params, s)
        dz = {'dt': ds, 'dw': dW[n]}
        if has_jumps:
            dz['dj'] = dJ[n]
            if dN is not None:
                info['jump_count'] += dN[n]           # integration.py:2618
        terms = coeffs(x, p)
        if scheme == 'milstein':
            x = milstein(x, terms, dz, diffusion_dx(x, p))
        else:
            x = _euler_sum(x, terms, dz)
        if n + 1 in out_of:
            rows.append(emit(x))
    xx = np.stack(rows)
    if is_log:
        xx = np.exp(xx)                               # integration.py:1195-1196
    return xx, info


def generic_replay(sde, params, x0, grid, where, dW, log=False,
                   scheme='euler', diffusion_dx=None, dN=None, dJ=None,
                   let=None, info_next=None):
    """Replay driver for a user ``sde(t, x, **params) -> {'dt':.., 'dw':..}``
    function (the ``integrate`` decorator path, integration.py:1843-1955);
    optionally with 'dn' / 'dj' terms fed from recorded increments, a user
    ``let`` (the value stored at output points, SDE.store -> let, 1175-1186,
    1501-1528) and a user ``info_next(last_x, new_x) -> {key: increment}``
    accumulated after every step (SDE.next -> info_next, 1169-1173).  Returns
    ``xx``, or ``(xx, info)`` when ``info_next`` is given."""
    grid = np.asarray(grid, dtype=float)
    out_of = {int(j): i for i, j in enumerate(where)}
    dW = np.asarray(dW)
    x = np.empty(dW.shape[1:])
    x[...] = np.asarray(x0)
    if log:
        x = np.log(x)
    store = (lambda z: z.copy()) if let is None else let
    rows = [store(x)] if 0 in out_of else []
    info = {}
    for n in range(grid.size - 1):
        s, ds = grid[n], grid[n+1] - grid[n]
        p = _at(params, s)
        A = sde(s, x, **p)
        terms = [(A[k], k) for k in A.keys()]
        dz = {'dt': ds, 'dw': dW[n]}
        if dN is not None:
            dz['dn'] = dN[n]
        if dJ is not None:
            dz['dj'] = dJ[n]
        last = x
        if scheme == 'milstein':
            x = milstein(x, terms, dz, diffusion_dx(s, x, **p))
        else:
            x = _euler_sum(x, terms, dz)
        if info_next is not None:
            for k, v in info_next(last, x).items():
                info[k] = info.get(k, 0) + v
        if n + 1 in out_of:
            rows.append(store(x))
    xx = np.stack(rows)
    xx = np.exp(xx) if log else xx
    return xx if info_next is None else (xx, info)


def system_replay(sde, q, params, x0s, grid, where, dW, addaxis, dN=None, dJ=None,
                  let=None, info_next=None):
    """Replay driver for a user system ``sde(t, x1..xq, **params) -> (dict,)*q``
    (``SDEs``, integration.py:1584-1835).  The q equations are stacked along a
    new axis -2 (``addaxis``) or along the last axis of vshape, equation k
    owning ``[k*d, (k+1)*d)`` (unpack/pack, integration.py:1735-1777); the
    stacked coefficients then go through the plain Euler update (718)."""
    grid = np.asarray(grid, dtype=float)
    out_of = {int(j): i for i, j in enumerate(where)}
    dW = np.asarray(dW)
    wshape, paths = dW.shape[1:-1], dW.shape[-1]
    if addaxis:
        vshape = wshape[:-1]
        unpack = lambda X: tuple(X[..., k, :] for k in range(q))
        idx = np.index_exp[..., np.newaxis, :]
    else:
        d = wshape[-1]//q
        vshape = wshape[:-1] + (d,)
        unpack = lambda X: tuple(X[..., k*d:(k + 1)*d, :] for k in range(q))
        idx = np.index_exp[...]
    pack = lambda zs: np.concatenate(
        tuple(np.broadcast_to(z, vshape + (paths,))[idx] for z in zs), axis=-2)
    X = pack(tuple(np.asarray(z, dtype=float) for z in x0s))
    rows = [X.copy()] if 0 in out_of else []
    info = {}
    for n in range(grid.size - 1):
        s, ds = grid[n], grid[n+1] - grid[n]
        As = sde(s, *unpack(X), **_at(params, s))
        ids = sorted(set().union(*(a.keys() for a in As)))
        terms = [(pack(tuple(a.get(k, 0) for a in As)), k) for k in ids]
        # the reference sums in the iteration order of a set of ids (1725-1729);
        # sorted order is the convention of the kernel and of the fixtures
        dz = {'dt': ds, 'dw': dW[n]}
        if dN is not None:
            dz['dn'] = dN[n]
        if dJ is not None:
            dz['dj'] = dJ[n]
        last = X
        X = _euler_sum(X, terms, dz)
        if info_next is not None:
            for k, v in info_next(unpack(last), unpack(X)).items():
                info[k] = info.get(k, 0) + v
        if n + 1 in out_of:
            rows.append(X.copy())
    if let is not None:
        # a user let() over the unpacked equations (one stored array)
        xx = np.stack([let(*unpack(r)) for r in rows])
        return xx if info_next is None else (xx, info)
    out = tuple(np.stack([unpack(r)[k] for r in rows]) for k in range(q))
    return out if info_next is None else (out, info)


def milstein(x, terms, dz, b_dx):
    """Milstein update ``x + a ds + b dw + (1/2) b b' (dw^2 - ds)``.

    sdepy ships Euler-Maruyama only (integration.py:615-618), so there is no
    reference Milstein to copy.  PINNED instead to the reference's own integrator
    machinery running the scheme through its documented ``method='<id>'`` ->
    ``<id>_next`` hook (integration.py:675-685): tests/golden/
    make_milstein_plugin.py plugs a ``milstein_next`` (the reference's
    ``euler_next`` followed by this correction, every product and sum separately
    rounded) into a class generated by ``sdepy.integrate`` and records the run;
    tests/test_oracle_golden.py reproduces it bit for bit.
    """
    a = dict((k, c) for c, k in terms)
    ds, dw = dz['dt'], dz['dw']
    euler = _euler_sum(x, terms, dz)
    b = a['dw']
    corr = ((b*b_dx)*0.5)*(dw*dw - ds)
    return euler + corr


# --------------------------------------------------------------------------
# self-driven sources: same numpy/scipy calls, same order as the reference
# --------------------------------------------------------------------------

def rho_to_corr(rho):
    """infrastructure.py:128-150: scalar rho -> [[1,r],[r,1]]; vector of
    length K -> [[I, diag r], [diag r, I]]."""
    rho = np.asarray(rho)
    if rho.size == 1:
        r = rho.reshape(())
        return np.array(((1, r), (r, 1)))
    r = rho.reshape(rho.size)
    eye, dg = np.eye(r.size), np.diag(r)
    return np.block([[eye, dg], [dg, eye]])


def draw_wiener(rng, t, dt, wshape, paths, corr=None):
    """infrastructure.py:1503-1560 for scalar ``t``/``dt``: iid normals, or
    ``multivariate_normal`` (drawn as (..., paths, M) then swapped, 1532-1552;
    time-dependent correlation sampled at the midpoint t+dt/2, 1540), then
    scaled by sqrt|dt|."""
    if corr is None:
        dz = rng.normal(0., 1., size=wshape + (paths,))
    else:
        cov = corr(t + dt/2) if callable(corr) else np.asarray(corr)
        if cov.ndim == 3:
            cov = cov[..., 0]
        m = wshape[-1]
        dz = rng.multivariate_normal(
            mean=np.zeros(m), cov=cov, size=wshape[:-1] + (paths,))
        dz = dz.swapaxes(-1, -2)
    dz *= np.sqrt(np.abs(np.asarray(dt).reshape((1,)*(len(wshape) + 1))))
    return dz


def draw_poisson(rng, t, dt, wshape, paths, lam):
    """infrastructure.py:1617-1633: sign(dt) * Poisson(|dt|*lam(t+dt/2))."""
    lam_ = lam(t + dt/2) if callable(lam) else lam
    sign = int(np.sign(dt))
    return sign*rng.poisson(abs(dt)*np.asarray(lam_), wshape + (paths,))


class _double_exp:
    """infrastructure.py:1757-1776: three full-size draws (exp+, exp-,
    uniform) for every element, then a select."""
    def __init__(self, a, b, pa):
        self.a, self.b, self.pa = a, b, pa

    def rvs(self, size, random_state):
        plus = _st.expon(scale=self.a).rvs(size=size, random_state=random_state)
        minus = _st.expon(scale=self.b).rvs(size=size, random_state=random_state)
        u = _st.uniform(scale=1.).rvs(size=size, random_state=random_state)
        return np.where(u <= self.pa, plus, -minus) + 0


def jump_law(kind, **kw):
    """The reference's preset jump-size laws (infrastructure.py:1653-1776)."""
    if kind == 'norm':
        return _st.norm(loc=kw['a'], scale=kw['b'])
    if kind == 'uniform':
        return _st.uniform(loc=kw['a'], scale=kw['b'] - kw['a'])
    if kind == 'double_exp':
        return _double_exp(kw['a'], kw['b'], kw['pa'])
    raise ValueError(kind)


def draw_cpoisson(rng, t, dt, wshape, paths, lam, law):
    """infrastructure.py:2017-2040: Poisson counts first, then for each
    realised count j = 1..max one block of (n_j, j) variates summed over the
    last axis.  Returns (dj, dn)."""
    dn = draw_poisson(rng, t, dt, wshape, paths, lam)
    sign = int(np.sign(dt))
    dj = np.zeros(wshape + (paths,))
    pos = sign*dn
    for j in range(1, int(pos.max()) + 1):
        hit = (pos == j)
        if hit.any():
            y = law.rvs(size=(int(hit.sum()), j), random_state=rng)
            dj[hit] = sign*y.sum(axis=-1)
    return dj, dn


def self_driven(model, params, x0, timeline, steps, paths, rng, wshape=(),
                corr=None, lam=None, law=None, y0=None, full=False):
    """Run ``model`` drawing increments the way the reference does
    (sources called in sorted-id order dj < dn < dt < dw,
    integration.py:1135, 1233).  This is the CPU baseline in bench.py and
    reproduces a same-seed reference run bit for bit."""
    tt, grid, where = step_grid(timeline, steps)
    n = grid.size - 1
    dW = np.empty((n,) + wshape + (paths,))
    dJ = dN = None
    if model == 'jumpdiff':
        dJ = np.empty((n,) + wshape + (paths,))
        dN = np.empty((n,) + wshape + (paths,), dtype=np.int64)
    for i in range(n):
        s, ds = grid[i], grid[i+1] - grid[i]
        if model == 'jumpdiff':
            dJ[i], dN[i] = draw_cpoisson(rng, s, ds, wshape, paths, lam, law)
        dW[i] = draw_wiener(rng, s, ds, wshape, paths, corr)
    return euler_replay(model, params, x0, grid, where, dW, dJ, dN,
                        y0=y0, full=full)


def heston_stream(params, x0, y0, rho, grid, paths, rng):
    """Memory-light Heston terminal run (no increment table kept): the CPU
    baseline kernel for the north-star config.  Same draws/arithmetic as
    ``self_driven('heston', ...)``; returns terminal x and negative_y_count."""
    corr = rho_to_corr(rho)
    a = np.full(paths, math.log(x0))
    y = np.full(paths, float(y0))
    neg = np.zeros(paths, dtype=np.int64)
    p = {k: np.asarray(v) for k, v in params.items()}
    for n in range(len(grid) - 1):
        s, ds = grid[n], grid[n+1] - grid[n]
        dw = draw_wiener(rng, s, ds, (2,), paths, corr)
        tx, ty = _heston_pair(a, y, p)
        neg += (y < 0)
        a, y = (_euler_sum(a, tx, {'dt': ds, 'dw': dw[0]}),
                _euler_sum(y, ty, {'dt': ds, 'dw': dw[1]}))
    return np.exp(a), neg


# --------------------------------------------------------------------------
# statistics                        infrastructure.py:861-889, 2869-3076
# --------------------------------------------------------------------------

def pmean(xx):
    return xx.mean(axis=-1, keepdims=True)


def pvar(xx, ddof=0):
    return xx.var(axis=-1, ddof=ddof, keepdims=True)


def pstd(xx, ddof=0):
    return xx.std(axis=-1, ddof=ddof, keepdims=True)


class moments_histogram:
    """``montecarlo`` restated (infrastructure.py:2869-3076): the first
    sample fixes the centring constant (its mean, 2934) and the bin edges
    (``np.histogram(bins, range)``, 2999-3004); later samples reuse them, and
    values outside the edges are counted apart (3013).  Moments 1..4 of
    ``sample - centre`` are cumulated as running means (2944-2953)."""

    def __init__(self, bins=100, range=None):
        self.bins, self.range = bins, range
        self.n = 0

    def update(self, sample):
        sample = np.asarray(sample, dtype=float)
        m = sample.shape[-1]
        if self.n == 0:
            self.centre = sample.mean(axis=-1)
            self.mom = [np.zeros(sample.shape[:-1]) for _ in range(4)]
            self.mean_ = np.zeros(sample.shape[:-1])
        d = sample - self.centre[..., None]
        pw = d
        for k in range(4):
            self.mom[k] = (self.n*self.mom[k] + m*pw.mean(axis=-1))/(self.n + m)
            pw = pw*d
        self.mean_ = (self.n*self.mean_ + m*sample.mean(axis=-1))/(self.n + m)
        if self.bins is not None:
            flat = sample.reshape(-1, m)
            if self.n == 0:
                self.edges, self.counts = [], []
                for row in flat:
                    c, e = np.histogram(row, bins=self.bins, range=self.range)
                    self.counts.append(c.astype(np.int64)); self.edges.append(e)
                self.outside = [m - int(c.sum()) for c in self.counts]
            else:
                for i, row in enumerate(flat):
                    c, _ = np.histogram(row, bins=self.edges[i])
                    self.counts[i] += c
                    self.outside[i] += m - int(c.sum())
        self.n += m

    def mean(self):
        return self.mean_

    def var(self):
        return self.mom[1] - self.mom[0]*self.mom[0]

    def std(self):
        return np.sqrt(self.var())

    def stderr(self):
        return np.sqrt(self.var()/(self.n - 1))

    def skew(self):
        m1, m2, m3 = self.mom[:3]
        return (m3 - 3*m1*m2 + 2*m1**3)/(m2 - m1*m1)**1.5

    def kurtosis(self):
        m1, m2, m3, m4 = self.mom
        # NB infrastructure.py:3062-3067: by operator precedence the ``-3.0``
        # only applies to the paths<2 branch -- the value is the RAW kurtosis.
        return (m4 - 4*m1*m3 + 6*m1*m1*m2 - 3*m1**4)/(m2 - m1*m1)**2
