"""
SDE integration framework -- host-side mirror of ``sdepy/integration.py``
(``paths_generator``, ``integrator``, ``SDE``, ``SDEs``, ``integrate`` and the
preset processes) whose step loop is ONE fused CUDA kernel launch
(``csrc/sde_engine.cuh`` through ``sdeb_integrate``) instead of the
reference's per-step Python/NumPy loop (``integration.py:392-474``).

The construction surface is the reference's: keyword-only ``__init__``,
argument consolidation by signature inspection (``integration.py:986-1055``),
overridable ``sde / init / shapes / more / source_*`` hooks, ``method=`` and
``dw= / dj=`` plug points, the same error conventions.  Added keywords:
``seed`` (Philox key), ``output`` ('process' host container -- the drop-in
default --, 'device' HBM-resident container, 'stats' fused statistics only),
``device``, ``path_offset`` (global index of this shard's first path),
``payoff``, ``draws`` ('fast': 64 random bits per pair of normals, the default;
'full': 96 bits, a 52-bit radius uniform like numpy's 53-bit ziggurat).  There
is no CPU fallback.
"""
import numpy as np
import torch

from . import _cuda, _engine, _jit, _lib
from .infrastructure import (
    process, device_process, wiener_source, poisson_source, cpoisson_source,
    odd_wiener_source, even_cpoisson_source, even_poisson_source,
    replay_source, norm_rv, double_exp_rv, lane_values, stack_lane_columns,
    _shape_setup, _const_param_setup, _variable_param_setup, _source_setup,
    _get_default_rng, _signature, _empty)


# --------------------------------------------------------------------------
# paths_generator: timeline handling (reference integration.py:40-584)
# --------------------------------------------------------------------------

class paths_generator:
    """Timeline validation, step-grid construction and launch of the device
    integration.  ``pace`` may be overridden as in the reference; the
    ``begin/next/store/end`` hooks of the reference's Python loop have no role
    on the device path (the loop is the kernel) and are kept as inert hooks."""

    depth = 2

    def __init__(self, *, paths=1, xshape=(), wshape=(), dtype=None,
                 steps=None, i0=0, info=None, getinfo=True):
        self.paths = paths
        self.xshape = _shape_setup(xshape)
        self.wshape = _shape_setup(wshape)
        self.dtype = dtype
        self.steps, self.i0 = steps, i0
        self.info = {} if info is None else info
        self.getinfo = getinfo
        super().__init__()

    def pace(self, timeline):
        """Target integration points to be merged with the output timeline
        (reference integration.py:189-218): an integer ``steps`` is the number
        of points of ``linspace(t0, t1, steps)``."""
        steps, ttype = self.steps, timeline.dtype
        if timeline.size == 1:
            return timeline
        if steps is None:
            return np.array((), dtype=ttype)
        if np.isscalar(steps):
            return np.linspace(timeline[0], timeline[-1], steps, dtype=ttype)
        return np.fromiter(steps, dtype=ttype)

    def begin(self):
        pass

    def next(self):
        pass

    def store(self, i, k):
        pass

    def end(self):
        pass

    def exit(self, tt, xx):
        return tt, xx

    # per-step hooks of the reference's Python loop (integration.py:392-474,
    # 1154-1234): the kernel replaces them, so a user override cannot take effect
    _stepping_hooks = ('begin', 'next', 'store', 'end', 'A', 'dZ', 'info_begin',
                       'info_next', 'info_store', 'info_end')

    def _check_no_python_stepping(self):
        for name in self._stepping_hooks:
            f = getattr(type(self), name, None)
            if f is not None and not getattr(f, '__module__', '').startswith(__package__):
                raise NotImplementedError(
                    '{}.{} is a Python per-step hook: the integration runs as '
                    'one CUDA kernel and cannot call it (no CPU stepping path). '
                    "Available device schemes: method='euler' | 'milstein'"
                    .format(type(self).__name__, name))

    def _device_run(self, tt, grid):
        raise NotImplementedError(
            '{} does not define a device integration: the CUDA path runs SDE '
            'subclasses (preset or traced), not arbitrary Python next() hooks'
            .format(type(self).__name__))

    def __call__(self, timeline):
        """Integrate along ``timeline`` (reference integration.py:476-584)."""
        # the state and all arithmetic are fp64 in registers; float32 / float16
        # are STORAGE types of the output (reference integration.py:495-496, 550)
        dtype = float if self.dtype is None else self.dtype
        if np.dtype(dtype) not in _OUT_DTYPES:
            raise NotImplementedError(
                'the CUDA path stores float64, float32 or float16 paths '
                '(dtype={} requested)'.format(dtype))
        if np.dtype(dtype) != np.dtype(float) and getattr(self, 'output', 'process') == 'device':
            raise NotImplementedError(
                "output='device' containers are float64: use the default "
                "output='process' with dtype={}".format(np.dtype(dtype)))
        if self.depth < 2:
            raise ValueError('the depth of the integrator algorithm should be '
                             '>= 2, not {}'.format(self.depth))
        tt = np.asarray(timeline)
        if tt.shape != (tt.size,):
            raise ValueError(
                'the integration timeline should be a one-dimensional array, '
                'not an array of shape {}'.format(tt.shape))
        if not np.array_equal(tt, np.unique(tt)):
            raise ValueError(
                'the integration timeline sholud be an array of strictly '
                'increasing numbers, but {} was given'.format(tt))
        i0 = self.i0
        try:
            if not np.isscalar(tt[i0]):
                raise IndexError()
        except IndexError:
            raise IndexError(
                'i0 should be an integer indexing an element of the '
                'integration timeline, but {} was given'.format(i0))
        ttype = float if tt.dtype.kind == 'i' else tt.dtype
        tt, tt_asis = tt.astype(ttype), tt

        target = self.pace(tt)
        target = target[np.logical_and(target >= tt[0], target <= tt[-1])]
        grid = np.unique(np.concatenate((target, tt)))
        if len(grid) < self.depth - 1:
            raise ValueError('at least {} time points are needed for a '
                             'paths_generator with depth {}'
                             .format(self.depth - 1, self.depth))
        if self.getinfo:
            self.info.update(t0=tt[i0], tmin=tt[0], tmax=tt[-1],
                             computed_steps=0, stored_steps=0)
        self._check_no_python_stepping()
        xx = self._device_run(tt, grid)
        return self.exit(tt_asis, xx)


# --------------------------------------------------------------------------
# integrator (reference integration.py:596-774)
# --------------------------------------------------------------------------

class integrator(paths_generator):
    """Integration scheme selection.  ``method='euler'`` (Euler-Maruyama,
    reference integration.py:707-723) and ``method='milstein'`` are device
    schemes; as in the reference a method ``<id>`` is recognised when the
    class has an ``<id>_next`` attribute (integration.py:675-685)."""

    def _check_integration_method(self, id):
        if not hasattr(self, id + '_next'):
            raise ValueError(
                'unrecognized integration method {}: use \'euler\' or provide '
                'a properly defined `{}_next` integrator class method'
                .format(id, id))

    def __init__(self, *, paths=1, xshape=(), wshape=(), dtype=None,
                 steps=None, i0=0, info=None, getinfo=True, method='euler'):
        self.method = method
        self._check_integration_method(method)
        super().__init__(paths=paths, xshape=xshape, wshape=wshape,
                         dtype=dtype, steps=steps, i0=i0, info=info,
                         getinfo=getinfo)

    # device schemes: the attributes only declare that the kernel implements
    # the scheme (see sde_engine.cuh step functors / _jit.py)
    def euler_next(self):
        """x + sum_id A[id]*dZ[id], each product and sum separately rounded,
        association of integration.py:718 -- executed in-kernel."""
        raise RuntimeError('euler_next runs inside the CUDA kernel')

    def milstein_next(self):
        """Euler + (1/2) b b' (dw^2 - dt) -- executed in-kernel (traced SDEs)."""
        raise RuntimeError('milstein_next runs inside the CUDA kernel')

    _device_schemes = ('euler', 'milstein')

    def exit(self, tt, xx):
        return _wrap(tt, xx)

    def A(self, t, x):
        return {'dt': x + np.nan}

    def dZ(self, t, dt):
        return {'dt': dt + np.nan}


_OUT_DTYPES = {np.dtype(np.float64): _lib.F64, np.dtype(np.float32): _lib.F32,
               np.dtype(np.float16): _lib.F16}


def _wrap(tt, xx):
    """process for host arrays, device_process for CUDA tensors."""
    if isinstance(xx, torch.Tensor):
        return device_process(tt, xx)
    return process(t=tt, x=xx)


# --------------------------------------------------------------------------
# fused statistics result (output='stats')
# --------------------------------------------------------------------------

class path_stats:
    """Across-path statistics of the values a process would have stored, one
    entry per output time point and component -- what ``process.pmean / pvar /
    pstd`` (infrastructure.py:861-889) and ``montecarlo`` moments
    (2924-2959) would give, without materialising the paths.

    ``sums`` has shape ``(N,) + xshape + (NSTAT,)``: centred power sums S1..S4
    of ``(x - centre)``, min, max, payoff sum and payoff square sum.
    """

    def __init__(self, t, sums, centre, paths, info=None):
        self.t = np.asarray(t)
        self.sums = np.asarray(sums, dtype=float)
        self.centre = np.asarray(centre, dtype=float)
        self.paths = int(paths)
        self.info = info

    def allreduce(self, group=None):
        """Combine the shards of all ranks with ONE all-reduce (SUM): every rank
        places its packed block (centre, power sums, min, max, payoff sums, path
        count) in its own slot of a [world, ...] vector, so that after the
        collective each rank holds all the blocks and folds them in rank order
        -- re-centring the power sums on rank 0's centre first (the centres
        differ when x0 is path-dependent).  NCCL when the process group is NCCL,
        gloo on CPU."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return self
        backend = dist.get_backend(group)
        dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else 'cpu'
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        s = self.sums
        c = np.broadcast_to(self.centre, s.shape[1:-1]) if s.ndim > 1 else self.centre
        c = np.broadcast_to(c, s.shape[:-1])
        block = np.concatenate((s.ravel(), c.ravel(), [float(self.paths)]))
        buf = np.zeros((world, block.size))
        buf[rank] = block
        buf = torch.from_numpy(buf).to(dev)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        buf = buf.cpu().numpy()
        return _fold_rank_blocks(self, buf)

    def _m(self, k):
        return self.sums[..., k - 1]/self.paths

    def _proc(self, v):
        return process(t=self.t, x=np.asarray(v)[..., np.newaxis])

    def pmean(self):
        return self._proc(self.centre + self._m(1))

    def pvar(self, ddof=0):
        n = self.paths
        return self._proc((self.sums[..., 1] - n*self._m(1)**2)/(n - ddof))

    def pstd(self, ddof=0):
        return self._proc(np.sqrt(np.asarray(self.pvar(ddof=ddof))[..., 0]))

    def stderr(self):
        return self._proc(np.sqrt(np.asarray(self.pvar())[..., 0]/(self.paths - 1)))

    def skew(self):
        m1, m2, m3 = self._m(1), self._m(2), self._m(3)
        with np.errstate(invalid='ignore', divide='ignore'):     # zero variance at t0: nan
            return self._proc((m3 - 3*m1*m2 + 2*m1**3)/(m2 - m1*m1)**1.5)

    def kurtosis(self):
        m1, m2, m3, m4 = (self._m(k) for k in (1, 2, 3, 4))
        with np.errstate(invalid='ignore', divide='ignore'):
            return self._proc((m4 - 4*m1*m3 + 6*m1*m1*m2 - 3*m1**4)/(m2 - m1*m1)**2)

    def pmin(self):
        return self._proc(self.sums[..., 4])

    def pmax(self):
        return self._proc(self.sums[..., 5])

    def payoff_mean(self):
        return self._proc(self.sums[..., 6]/self.paths)

    def payoff_stderr(self):
        n = self.paths
        m, m2 = self.sums[..., 6]/n, self.sums[..., 7]/n
        return self._proc(np.sqrt(np.maximum(m2 - m*m, 0.)/(n - 1)))


def _fold_rank_blocks(proto, buf):
    """Fold the per-rank blocks [world, sums | centres | paths] of
    ``path_stats.allreduce`` in rank order; power sums about centre c_r are
    moved to rank 0's centre c_0 binomially (d = c_r - c_0)."""
    s0 = proto.sums
    ns, nc = s0.size, s0[..., 0].size
    out, c0, paths = None, None, 0
    for r in range(buf.shape[0]):
        s = buf[r, :ns].reshape(s0.shape).copy()
        c = buf[r, ns:ns + nc].reshape(s0.shape[:-1])
        n = int(round(buf[r, ns + nc]))
        if n == 0:
            continue
        if out is None:
            out, c0, paths = s, c, n
            continue
        d = c - c0
        S1, S2, S3, S4 = (s[..., k] for k in range(4))
        if np.any(d != 0.):
            T1 = S1 + n*d
            T2 = S2 + 2*d*S1 + n*d**2
            T3 = S3 + 3*d*S2 + 3*d**2*S1 + n*d**3
            T4 = S4 + 4*d*S3 + 6*d**2*S2 + 4*d**3*S1 + n*d**4
            S1, S2, S3, S4 = T1, T2, T3, T4
        for k, S in enumerate((S1, S2, S3, S4)):
            out[..., k] += S
        out[..., 4] = np.minimum(out[..., 4], s[..., 4])
        out[..., 5] = np.maximum(out[..., 5], s[..., 5])
        out[..., 6:] += s[..., 6:]
        paths += n
    if out is None:
        return proto
    centre = c0[0] if proto.centre.ndim < c0.ndim else c0
    return path_stats(proto.t, out, centre, paths, proto.info)


# --------------------------------------------------------------------------
# SDE (reference integration.py:781-1581)
# --------------------------------------------------------------------------

class SDE(_jit._traced):
    """A user- or preset-defined Ito SDE, cooperating with ``integrator``
    (which must follow in the MRO).  Same construction protocol as the
    reference ``SDE`` class; the equation runs on the GPU either through a
    hand-written preset functor (``_preset``) or, for an arbitrary ``sde``
    method, through tracing + NVRTC (``_jit.py``)."""

    sources = {'dt', 'dw'}
    log = False
    q = None
    addaxis = None
    _preset = None          # name of the hand-written kernel functor, if any
    # traced equations: let() and info_next() are compiled into the kernel
    # (_jit._trace_let / _trace_info), info_begin / info_end run on the host
    # before / after the launch; the other per-step hooks cannot take effect
    _stepping_hooks = ('begin', 'next', 'store', 'end', 'A', 'dZ', 'info_store')

    # ---- argument bookkeeping (reference 970-1076) -----------------------
    def _check_source_id(self, id):
        if not hasattr(self, 'source_' + id):
            raise ValueError(
                "unrecognized source {}: use one of 'dt', 'dw', 'dn', 'dj', "
                'or provide a properly defined SDE class method `source_{}`'
                .format(id, id))

    def _get_args(self, keys):
        return {k: z for k, z in self._args.items() if k in keys}

    def _inspect_args_defaults(self, sde_nvars=1):
        groups = []
        self._source_args_keys = {}
        for id in self.sources:
            self._check_source_id(id)
            d = dict(_signature(getattr(self, 'source_' + id)))
            self._source_args_keys[id] = set(d)
            groups.append(d)
        init_d = dict(_signature(self.init)[2:])            # after (t, out_x)
        sde_d = dict(_signature(self.sde)[1 + sde_nvars:])  # after (t, x...)
        more_d = dict(_signature(self.more))
        groups += [init_d, sde_d, more_d]
        expected, seen = {}, {}
        for d in groups:
            expected.update(d)
            for k, z in d.items():
                seen.setdefault(k, []).append(z)
        repeated = {k for k, zs in seen.items() if len(set(zs)) > 1}
        if repeated:
            raise TypeError('two or more incompatible defaults found for SDE '
                            'parameter(s) {}'.format(repeated))
        self._init_args_keys = set(init_d)
        self._sde_args_keys = set(sde_d)
        self._more_args_keys = set(more_d)
        self._expected_args = expected

    def _consolidate_args(self, **args):
        all_args = {**self._expected_args, **args}
        unexpected = set(all_args).difference(self._expected_args)
        missing = {k for k, v in all_args.items() if v is _empty}
        if unexpected:
            raise TypeError('unexpected keyword(s): {}'.format(unexpected))
        if missing:
            raise TypeError('no value and no default found for sde '
                            'parameter(s) {}'.format(missing))
        return all_args

    # ---- construction (reference 1081-1148) ------------------------------
    def __init__(self, *, paths=1, vshape=(), dtype=None, rng=None,
                 steps=None, i0=0, info=None, getinfo=True, method='euler',
                 seed=None, output='process', device=None, path_offset=0,
                 payoff=None, draws='fast', **args):
        if not isinstance(self, integrator):
            raise TypeError(
                'cannot instantiate SDE subclass {} that is not a subclass of '
                'a cooperating integrator class'.format(type(self)))
        if hasattr(self, 'method'):
            raise TypeError(
                'improper method resolution order in class {}: the integrator '
                'class cannot precede the the SDE class'.format(type(self)))
        if output not in ('process', 'device', 'stats'):
            raise ValueError("output must be 'process', 'device' or 'stats', "
                             'not {!r}'.format(output))
        if draws not in ('fast', 'full'):
            raise ValueError("draws must be 'fast' (64 random bits per pair of normals: "
                             "32-bit radius uniform, 32-bit angle) or 'full' (96 bits: 52-bit "
                             'radius uniform), not {!r}'.format(draws))
        self.draws = draws
        self.output, self.device = output, device
        self.path_offset, self.payoff = int(path_offset), payoff
        self.vshape = _shape_setup(vshape)
        self._inspect_args_defaults()
        self._args = self._consolidate_args(**args)
        self._args.update({k: _const_param_setup(z) for k, z in
                           self._get_args(self._init_args_keys).items()})
        self._args.update({k: _variable_param_setup(z) for k, z in
                           self._get_args(self._sde_args_keys).items()})
        self.vshape, self.xshape, self.wshape = self.shapes(self.vshape)
        self.paths, self.dtype = paths, dtype
        self._rng_asis = rng
        self._rng = _get_default_rng() if rng is None else rng
        self._seed = seed
        self.sources = {
            id: getattr(self, 'source_' + id)(
                **self._get_args(self._source_args_keys[id]))
            for id in self.sources}
        self._ordered_source_ids = sorted(self.sources)
        self.more(**self._get_args(self._more_args_keys))
        super().__init__(paths=self.paths, xshape=self.xshape,
                         wshape=self.wshape, dtype=self.dtype, steps=steps,
                         i0=i0, info=info, getinfo=getinfo, method=method)

    rng = property(lambda self: self._rng)
    args = property(lambda self: self._args)

    # ---- hooks with the reference's signatures ---------------------------
    def shapes(self, vshape):
        return vshape, vshape, vshape

    def source_dt(self):
        def dt(s, ds):
            return ds
        dt.paths, dt.vshape = self.paths, self.wshape
        return dt

    def _source_kw(self):
        kw = dict(rng=self._rng_asis)
        if self._seed is not None:
            kw['seed'] = self._seed
        return kw

    def source_dw(self, dw=None, corr=None, rho=None):
        """Wiener source: instance, class or None (reference 1311-1344)."""
        import inspect
        extra = self._source_kw() if (dw is None or (
            inspect.isclass(dw) and issubclass(dw, wiener_source))) else dict(rng=self._rng_asis)
        return _source_setup(dw, wiener_source, paths=self.paths,
                             vshape=self.wshape, dtype=self.dtype, corr=corr,
                             rho=rho, **extra)

    def source_dn(self, dn=None, ptype=int, lam=1.):
        """Poisson source (reference 1346-1381); ``SDE(seed=...)`` reaches it
        like the other Philox-backed sources."""
        import inspect
        extra = self._source_kw() if (dn is None or (
            inspect.isclass(dn) and issubclass(dn, poisson_source))) else dict(rng=self._rng_asis)
        return _source_setup(dn, poisson_source, paths=self.paths,
                             vshape=self.wshape, dtype=ptype, lam=lam, **extra)

    def source_dj(self, dj=None, dn=None, ptype=int, lam=1., y=None):
        """Compound Poisson source (reference 1383-1423)."""
        import inspect
        extra = self._source_kw() if (dj is None or (
            inspect.isclass(dj) and issubclass(dj, cpoisson_source))) else dict(rng=self._rng_asis)
        return _source_setup(dj, cpoisson_source, paths=self.paths,
                             vshape=self.wshape, dtype=self.dtype, dn=dn,
                             ptype=ptype, lam=lam, y=y, **extra)

    def more(self):
        pass

    def init(self, t, out_x, x0=1.):
        out_x[...] = x0

    def sde(self, t, x):
        return {'dt': x + np.nan}

    def let(self, t, out_x, x):
        out_x[...] = x

    def result(self, tt, xx):
        return _wrap(tt, xx)

    def info_begin(self):
        pass

    def info_next(self):
        pass

    def info_store(self):
        pass

    def info_end(self):
        pass

    def exit(self, tt, xx):
        """Final wrap-up (reference 1193-1197); the ``exp`` of log-processes
        has already been applied by the kernel at store time."""
        if isinstance(xx, path_stats):
            return xx
        return self.result(tt, xx)

    # ---- reference-facing evaluation of the equation (host, for probing) --
    def _check_sde_values(self, A):
        if not isinstance(A, dict):
            raise TypeError('invalid {} return values: a dict, not a {} object '
                            'expected'.format(self.sde, type(A)))
        if not set(A.keys()).issubset(self.sources):
            raise KeyError(
                'invalid {} return values: {} entries expected (one per '
                'stochasticity source), not {}'
                .format(self.sde, set(self.sources), set(A.keys())))

    def _sde_args_at(self, t):
        return {k: (z(t) if callable(z) else z)
                for k, z in self._get_args(self._sde_args_keys).items()}

    def A(self, t, x):
        A_ = self.sde(t, x, **self._sde_args_at(t))
        self._check_sde_values(A_)
        return A_

    # ---- lowering to the kernel -------------------------------------------
    def _initial_state(self, t0):
        """Working state at t0 on the host, through the (overridable) ``init``
        hook and the log transform (reference begin(), 1154-1164).  Shape
        ``wshape + (1,)`` or, for path-dependent x0, ``wshape + (paths,)``."""
        init_args = self._get_args(self._init_args_keys)
        dtype = float if self.dtype is None else self.dtype
        for npaths in (1, self.paths):
            # the reference's working array has the requested dtype: x0 is
            # rounded to it when `init` writes it (integration.py:552-555)
            w = np.full(self.wshape + (npaths,), np.nan, dtype=dtype)
            try:
                self._init_paths = npaths
                self.init(t0, w, **init_args)
                break
            except ValueError:
                if npaths == self.paths:
                    raise
        if self.log:
            w = np.log(w)
        return w

    def _lanes(self):
        """(lead_shape, ncomp): lanes = prod(lead_shape), each owning the last
        ``ncomp`` working components.  Scalar equations are coupled along the
        last working axis only when the Wiener source is correlated."""
        dw = self.sources.get('dw')
        if isinstance(dw, wiener_source) and dw.corr is not None:
            return self.wshape[:-1], self.wshape[-1]
        return self.wshape, 1

    def _noise_plan(self, segs):
        """Decide philox vs replay and gather the replay tables.

        Device-generated draws need every source to be one of this package's
        Philox-backed sources; otherwise ALL sources are evaluated on the host
        step by step -- in sorted-id order like the reference (integration.py:
        1233) -- and fed to the kernel in replay mode.  That covers replay
        tables, the reference's own source objects and ``process`` instances.
        """
        dw, slots = self.sources.get('dw'), self._jump_slots()
        jsrc = [z for _, z in slots]
        unknown = set(self.sources) - {'dt', 'dw', 'dj', 'dn'}
        if unknown:
            raise NotImplementedError(
                'sources {} have no device implementation'.format(unknown))
        philox_ok = (dw is None or type(dw) in (wiener_source, odd_wiener_source)) and all(
            (type(z) in (cpoisson_source, even_cpoisson_source) and z.device_ready()) or
            type(z) in (poisson_source, even_poisson_source) for z in jsrc)
        if philox_ok:
            return None
        # tables per sweep: 'dW' array, 'dJ' LIST of arrays (one per jump slot:
        # 'dj' then 'dn'), 'dN' array (single slot exposing its counts)
        tables = []
        if all(isinstance(z, replay_source) for z in [dw] + jsrc if z is not None):
            # whole tables at once (no per-step host calls, no copies)
            at = 0
            for seg in segs:
                n = seg.n_steps
                tab = {'dW': dw.table[at:at + n]} if dw is not None else {
                    'dW': np.zeros((n,) + self.wshape + (self.paths,))}
                if jsrc:
                    tab['dJ'] = [z.table[at:at + n] for z in jsrc]
                    if len(jsrc) == 1 and jsrc[0].dn_table is not None:
                        tab['dN'] = jsrc[0].dn_table[at:at + n]
                for k, v in tab.items():
                    for u in (v if isinstance(v, list) else [v]):
                        if u.shape[0] != n:
                            raise ValueError(
                                'replay table {} holds {} steps, {} needed'
                                .format(k, u.shape[0], n))
                tables.append(tab)
                at += n
            return tables
        by_id = dict(slots)
        for seg in segs:
            rows = {'dW': [], 'dN': []}
            jrows = {id: [] for id, _ in slots}
            for s, ds in zip(seg.s, seg.ds):
                for id in self._ordered_source_ids:
                    if id in by_id:
                        z = by_id[id](s, ds)
                        jrows[id].append(self._as_lane_table(z))
                        if len(slots) == 1 and hasattr(by_id[id], 'dn_value'):
                            rows['dN'].append(self._as_lane_table(by_id[id].dn_value, integer=True))
                    elif id == 'dw':
                        # device-resident sources hand over CUDA tensors
                        call = getattr(dw, 'device_call', dw)
                        rows['dW'].append(self._as_lane_table(call(s, ds)))

            def stacked(v):
                return torch.stack(v) if isinstance(v[0], torch.Tensor) else np.stack(v)
            tab = {k: stacked(v) for k, v in rows.items() if v}
            # no Wiener term, or a sweep without steps (single-point timeline)
            empty = (seg.n_steps,) + self.wshape + (self.paths,)
            if 'dW' not in tab:
                tab['dW'] = np.zeros(empty)
            if slots:
                tab['dJ'] = [stacked(jrows[id]) if jrows[id] else np.zeros(empty)
                             for id, _ in slots]
            tables.append(tab)
        return tables

    def _lane_tables(self, tab, seg, spec):
        """Replay tables of one sweep in the kernel's layout [steps, lanes, paths];
        jump tables of several slots are interleaved per lane group:
        [steps, groups, slot, nw, paths]."""
        n, paths = seg.n_steps, self.paths

        def lanes(v):
            return self._to_lanes(v, 1).reshape((n, -1 if n else 0, paths))
        out = {}
        for k, v in tab.items():
            if not isinstance(v, list):
                out[k] = lanes(v)
            elif len(v) == 1:
                out[k] = lanes(v[0])
            else:
                parts = [lanes(u) for u in v]
                if any(isinstance(u, torch.Tensor) for u in parts):
                    parts = [u if isinstance(u, torch.Tensor) else torch.from_numpy(
                        np.ascontiguousarray(u)).to(next(
                            w for w in parts if isinstance(w, torch.Tensor)).device)
                        for u in parts]
                    parts = [u.to(torch.float64).reshape((n, spec.groups, spec.nw, paths))
                             for u in parts]
                    out[k] = torch.stack(parts, dim=2).reshape((n, -1, paths))
                else:
                    parts = [np.asarray(u, dtype=float).reshape((n, spec.groups, spec.nw, paths))
                             for u in parts]
                    out[k] = np.stack(parts, axis=2).reshape((n, -1, paths))
        return out

    def _as_lane_table(self, z, integer=False):
        shape = self.wshape + (self.paths,)
        if isinstance(z, torch.Tensor):
            return z.expand(shape) if tuple(z.shape) != shape else z
        z = np.asarray(z)
        if integer:
            z = z.astype(np.int64)
        return np.broadcast_to(z, shape)

    def _jump_slots(self):
        """[(id, source)] feeding the kernel's jump slots, in sorted-id order
        like the reference's calls (integration.py:1135): the compound Poisson
        source 'dj' and / or a plain Poisson source 'dn' (unit jump sizes)."""
        return [(id, self.sources[id]) for id in ('dj', 'dn') if id in self.sources]

    def _jump_source(self):
        slots = self._jump_slots()
        return slots[0][1] if slots else None

    def _jumps_time_dependent(self):
        """True when a jump intensity or jump-size law changes in time (both
        are sampled at step midpoints, infrastructure.py:1630, 2031)."""
        for _, dj in self._jump_slots():
            lam_src = getattr(dj, 'dn', dj)
            law = getattr(dj, 'y', None)
            if callable(getattr(lam_src, 'lam', None)) or any(
                    callable(z) for z in getattr(law, 'params', {}).values()):
                return True
        return False

    def _philox_key(self):
        for id in ('dw', 'dj', 'dn'):
            src = self.sources.get(id)
            if hasattr(src, 'next_key'):
                return src.next_key()
        return 0

    # _spec / _records / _stats_centre: generic traced lowering inherited from
    # _jit._traced; the preset equations below override them.

    def _device_run(self, tt, grid):
        if self.method not in self._device_schemes:
            raise NotImplementedError(
                'integration method {!r} has no device implementation '
                '(available: {})'.format(self.method, self._device_schemes))
        segs = _engine.segments_of(tt, grid, self.i0)
        # user info hooks of a traced equation: info_begin / info_end run here,
        # on the host, around the launch (reference integration.py:1166, 1189)
        user_info = self.getinfo and self._preset is None and not isinstance(
            self, _preset_SDE) and any(self._hook_overridden(h) for h in
                                       ('info_begin', 'info_next', 'info_end'))
        if user_info:
            self.itervars = dict(tt=tt, steps_tt=grid, i0=self.i0)
            self.info_begin()
        spec, lead = self._spec()
        replay = self._noise_plan(segs)
        # the lowered tables depend only on the grid and on the parameters:
        # repeated calls (Monte Carlo batches) reuse them
        key = self._lowering_key(tt, grid, replay is not None)
        cached = getattr(self, '_lowered', None)
        if cached is not None and cached[0] == key:
            w0l, w0_arg, records = cached[1]
        else:
            w0 = self._to_lanes(self._initial_state(tt[self.i0]))
            npaths0 = w0.shape[-1]
            w0l = w0.reshape((spec.groups, spec.nw, npaths0))
            w0_arg = w0l if npaths0 > 1 else w0l[..., 0]
            records = [self._records(spec, seg, lead, replay is not None)
                       for seg in segs]
            self._lowered = (key, (w0l, w0_arg, records))
        if replay is not None:
            replay = [self._lane_tables(tab, seg, spec) for tab, seg in zip(replay, segs)]
        want_stats = self.output == 'stats'
        centre = self._stats_centre(w0l) if want_stats else None
        jumps = spec.jumps
        anti = {}
        if replay is None:
            for key, src in (('anti_dw_half', self.sources.get('dw')),
                             ('anti_dj_half', self._jump_source())):
                if getattr(src, 'antithetic', False):
                    if self.path_offset:
                        raise NotImplementedError(
                            'antithetic sources pair paths k and K+k of ONE '
                            'launch: they cannot be combined with path_offset')
                    anti[key] = self.paths//2
        res = _engine.run(
            spec, segs, tt.size, records, w0_arg, paths=self.paths,
            path_offset=self.path_offset, seed=self._philox_key(),
            dev=self.device, replay=replay, want_out=not want_stats,
            want_stats=want_stats, centre=centre, payoff=self.payoff,
            counters=self.getinfo, dn_sums=self.getinfo and jumps,
            dump=getattr(self, '_dump_increments', False),
            out_dtype=np.dtype(float if self.dtype is None else self.dtype),
            trace_kernels=getattr(self, '_trace_kernels', False), **anti)
        if self.getinfo:
            self.info['computed_steps'] = int(sum(s.n_steps for s in segs))
            self.info['stored_steps'] = int(sum((s.store_row >= 0).sum() for s in segs))
            self._device_info(res, tt, segs, replay)
            if user_info:
                self.info_end()
                del self.itervars
        self._last_run = res
        xshape = self.xshape
        if want_stats:
            torch.cuda.current_stream(res.stats.device).synchronize()
            sums = self._from_lanes(res.stats.cpu().numpy().reshape(
                (tt.size, -1, _lib.NSTAT)), (_lib.NSTAT,))
            centre = self._from_lanes(np.asarray(centre).reshape((1, -1)), ())[0]
            return path_stats(tt, sums, centre, self.paths, self.info)
        xx = self._from_lanes(res.out.reshape((tt.size, -1, self.paths)),
                              (self.paths,))
        if self.output == 'process':
            xx = _cuda.to_host(xx)
        return xx

    # layout hooks: the kernel wants the components of one lane adjacent
    # (element-major); SDEs with addaxis=False store them variable-major
    def _to_lanes(self, a, pre=0):
        return a

    def _from_lanes(self, a, tail):
        return a.reshape(a.shape[:1] + self.xshape + tail)

    def _lowering_key(self, tt, grid, replay):
        def ident(z):
            if isinstance(z, np.ndarray):
                return (z.shape, z.tobytes()) if z.size <= 4096 else id(z)
            if isinstance(z, (int, float, complex, str, type(None))):
                return z
            return id(z)
        srcs = []
        for k in ('dw', 'dj', 'dn'):
            src = self.sources.get(k)
            srcs.append((id(src), ident(getattr(src, 'corr', None)), ident(getattr(src, 'lam', None)),
                         id(getattr(src, 'y', None))))
        return (tt.tobytes(), grid.tobytes(), self.i0, replay, self.method,
                tuple((k, ident(v)) for k, v in sorted(self._args.items())), tuple(srcs))

    # (_device_info: counters compiled from info_next, _jit._traced)

    def _counter(self, res, shape):
        c = res.counter.reshape(shape + (self.paths,))
        return c.cpu().numpy() if self.output == 'process' else c


# --------------------------------------------------------------------------
# SDEs (reference integration.py:1584-1835)
# --------------------------------------------------------------------------

class SDEs(SDE):
    """System of ``q`` equations stacked along the last working axis
    (``addaxis`` semantics of the reference, integration.py:1630-1645)."""

    q = 1
    addaxis = False
    _system = True

    def _inspect_args_defaults(self):
        if self.q < 1:
            raise ValueError('the number of equations q should be positive, '
                             'but {} was given'.format(self.q))
        if self.vshape == ():
            self.addaxis = True
        super()._inspect_args_defaults(sde_nvars=self.q)

    def _check_sde_values(self, As):
        if not isinstance(As, (list, tuple)):
            raise TypeError(
                'invalid {} return values: a list or tuple of {} dict (one per '
                'equation) expected, not a {} object'
                .format(self.sde, self.q, type(As)))
        if len(As) != self.q:
            raise ValueError('invalid {} return values: {} equations expected, '
                             'not {}'.format(self.sde, self.q, len(As)))
        for a in As:
            super()._check_sde_values(a)

    def A(self, t, X):
        As = self.sde(t, *self.unpack(X), **self._sde_args_at(t))
        self._check_sde_values(As)
        ids = set()
        for a in As:
            ids.update(a.keys())
        return {id: self.pack(tuple(a.get(id, 0) for a in As)) for id in ids}

    def unpack(self, X):
        """Split the stacked array into one array (or tensor view) per
        equation (reference 1735-1757)."""
        q = self.q
        if self.addaxis:
            return tuple(X[..., k, :] for k in range(q))
        d = self.vshape[-1]
        return tuple(X[..., k*d:(k + 1)*d, :] for k in range(q))

    def pack(self, xs):
        target = self.vshape + (getattr(self, '_init_paths', self.paths),)
        idx = np.index_exp[..., np.newaxis, :] if self.addaxis else np.index_exp[...]
        return np.concatenate(
            tuple(np.broadcast_to(x, target)[idx] for x in xs), axis=-2)

    def shapes(self, vshape):
        q = self.q
        if self.addaxis:
            xshape = wshape = vshape + (q,)
        else:
            xshape = wshape = vshape[:-1] + (vshape[-1]*q,)
        return vshape, xshape, wshape

    def init(self, t, out_X, x0=1.):
        x0s = x0
        if x0s.shape == ():
            x0s = (x0s,)*self.q
        out_X[...] = self.pack(x0s)

    def sde(self, t, x):
        return ({'dt': x + np.nan},)

    def result(self, tt, XX):
        return tuple(_wrap(tt, xx) for xx in self.unpack(XX))

    def _stacked_coupled(self):
        """addaxis=False with a correlated Wiener source: the correlation
        matrix couples the whole stacked axis (q*d components)."""
        dw = self.sources.get('dw')
        return (not self.addaxis and isinstance(dw, wiener_source)
                and dw.corr is not None)

    def _lanes(self):
        """Traced systems: one lane owns the q variables of an element of
        vshape.  With addaxis=True they are adjacent in the working array;
        with addaxis=False variable k of element h sits at k*d + h along the
        last axis (reference 1735-1757) and ``_to_lanes`` / ``_from_lanes``
        transpose between the two layouts on the way in and out -- unless
        the Wiener source is correlated along that axis: then one lane owns
        all its q*d components, in the reference's own order."""
        if self._stacked_coupled():
            return self.vshape[:-1], self.wshape[-1]
        return self.vshape, self.q

    def _to_lanes(self, a, pre=0):
        if self.addaxis or self._stacked_coupled():
            return a
        v, q = self.vshape, self.q
        a = a.reshape(a.shape[:pre] + v[:-1] + (q, v[-1]) + a.shape[pre + len(v):])
        ax = pre + len(v) - 1
        return a.swapaxes(ax, ax + 1)

    def _from_lanes(self, a, tail):
        if (self.addaxis or self._stacked_coupled()
                or (getattr(self, '_jit', None) or {}).get('let') == 'single'):
            return super()._from_lanes(a, tail)
        v, q = self.vshape, self.q
        a = a.reshape(a.shape[:1] + v + (q,) + tail)
        return a.swapaxes(len(v), len(v) + 1).reshape(a.shape[:1] + self.xshape + tail)


# --------------------------------------------------------------------------
# integrate decorator (reference integration.py:1843-1955)
# --------------------------------------------------------------------------

def _SDE_from_function(f, q=None, sources=None, log=False, addaxis=False):
    if q is not None and sources is not None:
        neq, ids = q, set(sources)
        base = SDE if neq == 0 else SDEs
    else:
        try:
            try:
                test = f()
            except Exception:
                test = f(np.array(1.), np.array(1.))
        except Exception:
            raise TypeError('test evaluation of {} failed'.format(f))
        if isinstance(test, (tuple, list)):
            neq = len(test)
            if neq == 0:
                raise ValueError('non empty list or tuple expected')
            base = SDEs
        else:
            neq, base, test = 0, SDE, (test,)
        ids = set()
        for z in test:
            ids.update(z.keys())
        if (q is not None and neq != q) or (
                sources is not None and set(sources) != ids):
            raise TypeError("test evaluation of {} inconsistent with given 'q' "
                            "or 'sources'".format(f))
    flags = dict(q=neq, sources=ids, log=log, addaxis=addaxis,
                 sde=staticmethod(f))
    return type('SDE_wrapper', (base,), flags)


def integrate(sde=None, *, q=None, sources=None, log=False, addaxis=False):
    """Decorator turning a function ``f(t, x, ..., **params) -> dict`` (or
    tuple of dicts) into an integrator class, as ``sdepy.integrate``
    (reference integration.py:1894-1955).  The function is traced once with
    symbolic state variables and compiled with NVRTC for sm_100a."""
    if sde is None:
        def decorator(sde):
            return integrate(sde, q=q, sources=sources, log=log, addaxis=addaxis)
        return decorator
    cls = _SDE_from_function(sde, q=q, sources=sources, log=log, addaxis=addaxis)
    return type('sde_integrator', (cls, integrator), {})


# --------------------------------------------------------------------------
# preset equations: hand-written kernel functors (sde_engine.cuh)
# --------------------------------------------------------------------------

class _preset_SDE(SDE):
    """Shared lowering of the preset scalar equations."""

    _model = None           # SDEB_MODEL_*
    _coupled = False        # lanes own the last working axis regardless of corr
    _device_schemes = ('euler',)
    # hand-written functors: no user hook can be compiled in
    _stepping_hooks = paths_generator._stepping_hooks + ('let',)

    def _device_info(self, res, tt, segs, replay):
        pass

    def _lanes(self):
        if self._coupled:
            return self.wshape[:-1], self.wshape[-1]
        return super()._lanes()

    # the hand-written functors address the reference's own working layout
    # (Heston: N x-components then N y-components, integration.py:2460-2466)
    _to_lanes = SDE._to_lanes
    _from_lanes = SDE._from_lanes

    def _spec(self):
        lead, ncomp = self._lanes()
        groups = int(np.prod(lead, dtype=int))
        if self.draws == 'full':
            # full-resolution draws: the same functor compiled with
            # -DSDEB_DRAW_FULL=1 (NVRTC, cached per model and component count)
            handle = _jit.instantiate_preset(self._model, ncomp, full=True)
            return _engine.problem_spec(_lib.MODEL_JIT, ncomp, groups,
                                        jit_handle=handle), lead
        try:
            return _engine.problem_spec(self._model, ncomp, groups), lead
        except _lib.SdebError:
            # component count not pre-instantiated in libsdeb.so: instantiate
            # the same hand-written functor with NVRTC
            handle = _jit.instantiate_preset(self._model, ncomp)
            return _engine.problem_spec(_lib.MODEL_JIT, ncomp, groups,
                                        jit_handle=handle), lead

    def _coeffs(self, p):
        """Per-component parameter tuple from the evaluated sde args, in the
        reference's operation order."""
        raise NotImplementedError

    def _records(self, spec, seg, lead, replay):
        sde_args = self._get_args(self._sde_args_keys)
        dw = self.sources.get('dw')
        dj = self.sources.get('dj')
        tdep = any(callable(z) for z in sde_args.values())
        corr_t = (not replay and isinstance(dw, wiener_source) and callable(dw.corr))
        jumps_t = spec.jumps and not replay and self._jumps_time_dependent()
        n = seg.n_steps if (tdep or corr_t or jumps_t) else 1
        n = max(n, 1) if seg.n_steps else 1
        ncomp = spec.ncomp
        # parameters broadcast against the working shape (+ the paths axis);
        # its prod(lead) x ncomp elements map to [group, component]
        full = self._param_target()
        assert int(np.prod(full, dtype=int)) == spec.groups*ncomp
        blocks = []
        for i in range(n):
            s = seg.s[i] if seg.n_steps else 0.
            ds = seg.ds[i] if seg.n_steps else 0.
            p = {k: (z(s) if callable(z) else z) for k, z in sde_args.items()}
            cols = [self._lane_matrix(c, full) for c in self._coeffs(p)]
            if spec.jumps:
                cols += self._jump_cols(dj, s, ds, full, replay)
            block = stack_lane_columns(cols, spec.groups, ncomp)   # [G, npc(, paths)]
            L = None
            if spec.nchol and not replay and isinstance(dw, wiener_source):
                L = dw.chol_at(s + ds/2)                           # midpoint, 1540
            blocks.append((block, _engine.chol_entries(L, spec.ndw)))
        return _engine.assemble_records(blocks, spec)

    def _param_target(self):
        return self.wshape

    def _lane_matrix(self, value, full):
        """Per-lane values of a coefficient: [lanes] or [lanes, paths]."""
        return lane_values(value, full, 'SDE parameter', paths=self.paths)

    def _jump_cols(self, dj, s, ds, full, replay):
        zero = np.zeros(int(np.prod(full, dtype=int)))
        if replay:
            return [zero]*6
        mid = s + ds/2
        lam = self._lane_matrix(dj.dn.lam_at(mid), full)     # midpoint, 1630
        kind, a, b, pa = dj.y.at(mid)                        # midpoint, 2031
        # the kernel forms |dt|*lam itself: the record is constant in time
        # unless lam or the jump law are
        return [lam, zero, zero + kind,
                self._lane_matrix(a, full), self._lane_matrix(b, full),
                self._lane_matrix(pa, full)]

    def _stats_centre(self, w0l):
        w = w0l.mean(axis=-1)                                # [G, nw]
        return np.exp(w) if self.log else w


class wiener_SDE(_preset_SDE):
    """dx = mu dt + sigma dw (reference integration.py:2056-2070)."""
    _model = _lib.MODEL_LINEAR

    def init(self, s, out_x, x0=0.):
        super().init(s, out_x, x0)

    def sde(self, t, x, mu=0., sigma=1.):
        return {'dt': mu, 'dw': sigma}

    def _coeffs(self, p):
        return p['mu'], p['sigma']


class wiener_process(wiener_SDE, integrator):
    """Wiener process with drift (reference integration.py:2073-2116)."""


class lognorm_SDE(_preset_SDE):
    """dx = mu x dt + sigma x dw, integrated as da = (mu - sigma^2/2) dt +
    sigma dw on a = log x (reference integration.py:2119-2130)."""
    _model = _lib.MODEL_LINEAR_LOG
    log = True

    def sde(self, t, x, mu=0., sigma=1.):
        return {'dt': mu - sigma*sigma/2, 'dw': sigma}

    def _coeffs(self, p):
        mu, sigma = p['mu'], p['sigma']
        return mu - sigma*sigma/2, sigma


class lognorm_process(lognorm_SDE, integrator):
    """Lognormal process (reference integration.py:2133-2182)."""


class ornstein_uhlenbeck_SDE(_preset_SDE):
    """dx = k (theta - x) dt + sigma dw (reference integration.py:2185-2199)."""
    _model = _lib.MODEL_MEANREV

    def init(self, s, out_x, x0=0.):
        super().init(s, out_x, x0)

    def sde(self, s, x, theta=0., k=1., sigma=1.):
        return {'dt': k*(theta - x), 'dw': sigma}

    def _coeffs(self, p):
        return p['theta'], p['k'], p['sigma']


class ornstein_uhlenbeck_process(ornstein_uhlenbeck_SDE, integrator):
    """Ornstein-Uhlenbeck process (reference integration.py:2202-2244)."""


class hull_white_SDE(_preset_SDE):
    """F-factor Hull-White: sum of F correlated mean-reverting factors
    (reference integration.py:2247-2272)."""
    _model = _lib.MODEL_HULL_WHITE
    _coupled = True

    def init(self, s, out_x, x0=0.):
        super().init(s, out_x, x0)

    def more(self, factors=1):
        pass

    def shapes(self, vshape):
        return vshape, vshape, vshape + (self.args['factors'],)

    def sde(self, s, x, theta=0., k=1., sigma=1.):
        return {'dt': k*(theta - x), 'dw': sigma}

    def let(self, s, out_x, x):
        out_x[...] = x.sum(axis=-2)

    def _coeffs(self, p):
        return p['theta'], p['k'], p['sigma']

    def _stats_centre(self, w0l):
        return w0l.mean(axis=-1).sum(axis=-1, keepdims=True)


class hull_white_process(hull_white_SDE, integrator):
    """F-factor Hull-White process (reference integration.py:2275-2315)."""


class hull_white_1factor_process(ornstein_uhlenbeck_process):
    """Synonym of ornstein_uhlenbeck_process (reference 2318-2339)."""


class cox_ingersoll_ross_SDE(_preset_SDE):
    """dx = k (theta - x+) dt + xi sqrt(x+) dw (reference 2342-2354)."""
    _model = _lib.MODEL_CIR

    def sde(self, s, x, theta=1., k=1., xi=1.):
        x_plus = np.maximum(x, 0.)
        return {'dt': k*(theta - x_plus), 'dw': xi*np.sqrt(x_plus)}

    def _coeffs(self, p):
        return p['theta'], p['k'], p['xi']


class cox_ingersoll_ross_process(cox_ingersoll_ross_SDE, integrator):
    """Cox-Ingersoll-Ross process (reference integration.py:2357-2400)."""


class full_heston_SDE(_preset_SDE, SDEs):
    """Heston stochastic volatility, full truncation, on (log x, y)
    (reference integration.py:2403-2444)."""
    _model = _lib.MODEL_HESTON_FULL
    q = 2
    addaxis = False
    log = False

    def init(self, s, out_X, x0=1., y0=1.):
        out_x, out_y = self.unpack(out_X)
        out_x[...] = np.log(x0)
        out_y[...] = y0

    def sde(self, t, x, y, mu=0., sigma=1., theta=1., k=1., xi=1.):
        y_plus = np.maximum(y, 0.)
        return ({'dt': mu - sigma*sigma*y_plus/2, 'dw': sigma*np.sqrt(y_plus)},
                {'dt': k*(theta - y_plus), 'dw': xi*np.sqrt(y_plus)})

    def _lanes(self):
        # one lane owns the N x- and N y-components of the last axis
        return self.wshape[:-1], self.wshape[-1]//2

    def _param_target(self):
        # parameters broadcast against vshape (x and y share them)
        return self.wshape[:-1] + (self.wshape[-1]//2,)

    def _coeffs(self, p):
        sigma = p['sigma']
        # sigma*sigma/2: halving commutes with rounding, so the kernel's
        # (s*s/2)*y+ equals the reference's s*s*y+/2 bit for bit
        return p['mu'], sigma*sigma/2, sigma, p['theta'], p['k'], p['xi']

    def result(self, tt, xx):
        xs, ys = self.unpack(xx)
        return _wrap(tt, xs), _wrap(tt, ys)

    def _stats_centre(self, w0l):
        w = w0l.mean(axis=-1)
        n = w.shape[-1]//2
        c = w.copy()
        c[:, :n] = np.exp(w[:, :n])
        return c

    def _device_info(self, res, tt, segs, replay):
        if res.counter is not None:
            self.info['negative_y_count'] = self._counter(res, self.vshape)


class full_heston_process(full_heston_SDE, integrator):
    """Heston process returning (x, y) (reference integration.py:2447-2514)."""


class heston_SDE(full_heston_SDE):
    """Heston process storing x only (reference integration.py:2517-2541)."""
    _model = _lib.MODEL_HESTON

    def shapes(self, vshape):
        vshape, xshape, wshape = super().shapes(vshape)
        return vshape, vshape, wshape

    def result(self, tt, xx):
        return _wrap(tt, xx)

    def _stats_centre(self, w0l):
        w = w0l.mean(axis=-1)
        return np.exp(w[:, :w.shape[-1]//2])


class heston_process(heston_SDE, integrator):
    """Heston process (reference integration.py:2544-2573)."""


class jumpdiff_SDE(_preset_SDE):
    """Lognormal diffusion with compound Poisson log-jumps, no martingale
    correction (reference integration.py:2576-2623)."""
    _model = _lib.MODEL_JUMPDIFF
    log = True
    sources = {'dt', 'dw', 'dj'}

    def sde(self, s, x, mu=0., sigma=1.):
        return {'dt': mu - sigma*sigma/2, 'dw': sigma, 'dj': 1}

    def _coeffs(self, p):
        mu, sigma = p['mu'], p['sigma']
        return mu - sigma*sigma/2, sigma

    def _device_info(self, res, tt, segs, replay):
        """jump_count / jump_rate (reference 2588-2623).  In replay mode they
        are only available when the dj source exposes ``dn_value`` (2615)."""
        have_dn = replay is None or all('dN' in tab for tab in replay)
        if replay is None and isinstance(self.sources.get('dj'), even_cpoisson_source):
            have_dn = False      # the reference's antithetic wrapper hides dn_value
        if not have_dn:
            return
        if res.counter is not None:
            self.info['jump_count'] = self._counter(res, self.vshape)
        # info_begin re-initialises the diagnostics at every sweep (2588-2592),
        # so they describe the LAST sweep (the forward one unless i0 is the
        # final point); index i is the next output point of that sweep.
        seg, dn = segs[-1], res.dn_sum[-1]
        rev = bool(seg.n_steps and seg.ds[0] < 0)
        tts = tt[self.i0::-1] if rev else tt[self.i0:]
        rate = np.zeros(tts.shape, dtype=float)
        i = 1
        for n in range(seg.n_steps if dn is not None else 0):
            rate[i - 1] += dn[n]/(tts[i] - tts[i - 1])/self.paths
            if seg.store_row[n] >= 0:
                i += 1
        if tts.size > 1:
            rate[-1] = rate[-2]
        self.info['jump_rate'] = rate


class jumpdiff_process(jumpdiff_SDE, integrator):
    """Jump-diffusion process (reference integration.py:2626-2695)."""


class merton_jumpdiff_SDE(jumpdiff_SDE):
    """Normal jump sizes (reference integration.py:2698-2710)."""

    def source_dj(self, dj=None, dn=None, ptype=int, lam=1., a=0., b=1.):
        return super().source_dj(dj=dj, dn=dn, ptype=ptype, lam=lam,
                                 y=norm_rv(a=a, b=b))


class merton_jumpdiff_process(merton_jumpdiff_SDE, integrator):
    """Merton jump-diffusion (reference integration.py:2713-2731)."""


class kou_jumpdiff_SDE(jumpdiff_SDE):
    """Double-exponential jump sizes (reference integration.py:2734-2747)."""

    def source_dj(self, dj=None, dn=None, ptype=int, lam=1., a=0.5, b=0.5,
                  pa=0.5):
        return super().source_dj(dj=dj, dn=dn, ptype=ptype, lam=lam,
                                 y=double_exp_rv(a=a, b=b, pa=pa))


class kou_jumpdiff_process(kou_jumpdiff_SDE, integrator):
    """Kou jump-diffusion (reference integration.py:2750-2770)."""
