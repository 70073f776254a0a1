"""
Containers, stochasticity sources and Monte Carlo statistics -- the host side
mirror of ``sdepy/infrastructure.py`` for the hot path only.

* ``process`` : host ``ndarray`` subclass ``(N,)+vshape+(paths,)`` with a
  timeline (reference ``infrastructure.py:272-456, 861-889``).
* ``device_process`` : the same container resident in HBM; across-path
  reductions run as CUDA kernels (``sdeb_moments``).
* ``wiener_source / poisson_source / cpoisson_source`` : descriptors of the
  in-kernel Philox draws (reference ``1388-1560, 1566-1633, 1881-2040``); they
  also obey the reference's source protocol ``dz(t, dt)`` by launching a draw
  kernel.  ``replay_source`` feeds pre-drawn increments (replay mode).
* ``montecarlo`` : cumulated moments + histograms (reference ``2718-3312``)
  computed on the device.
"""
import inspect

import numpy as np
import torch

from . import _cuda, _lib

# --------------------------------------------------------------------------
# default generator: only used to derive Philox seeds (reference:
# infrastructure.py:26-51)
# --------------------------------------------------------------------------
default_rng = np.random.default_rng()


def _get_default_rng():
    return default_rng


def _seed_from(rng):
    """64-bit Philox key drawn from a numpy Generator / RandomState."""
    if isinstance(rng, (int, np.integer, list, np.ndarray)):
        # reference infrastructure.py:1344-1348
        raise TypeError('`rng` must be an instance of a random number '
                        'generator, not {}.'.format(type(rng)))
    rng = default_rng if rng is None else rng
    if hasattr(rng, 'integers'):
        return int(rng.integers(0, 2**63 - 1, dtype=np.int64))
    if hasattr(rng, 'randint'):
        return (int(rng.randint(0, 2**31 - 1)) << 31) | int(rng.randint(0, 2**31 - 1))
    raise TypeError('`rng` must be a numpy.random Generator or RandomState, '
                    'not {}'.format(type(rng)))


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30))*0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27))*0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def _shape_setup(shape):
    return (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)


def _const_param_setup(z):
    return z if z is None else np.asarray(z)


def _variable_param_setup(z):
    """Reference infrastructure.py:64-108: arrays stay arrays, callables are
    probed once at t=1. to learn their shape and wrapped in asarray."""
    if z is None or isinstance(z, process):
        return z
    if callable(z):
        try:
            shape = np.asarray(z(1.)).shape
        except Exception:
            shape = None

        def wrapped(s, _z=z):
            return np.asarray(_z(s))
        wrapped.shape = shape
        return wrapped
    return np.asarray(z)


def _param_shape(z):
    if z is None:
        return None
    if isinstance(z, process):
        return z.vshape + (z.paths,)
    return z.shape


def _rho_to_corr(rho):
    """Reference infrastructure.py:128-150."""
    if rho is None:
        return None
    rho = np.asarray(rho)
    n = rho.size
    if rho.shape not in {(), (n,), (n, 1)}:
        raise ValueError(
            'correlation ``rho`` should be a vector, possibly with a trailing '
            '1-dimensional axis matching the paths axis, not an array with '
            'shape {}'.format(rho.shape))
    if n == 1:
        r = rho.reshape(())
        return np.array(((1, r), (r, 1)))
    r = rho.reshape(n)
    eye, dg = np.eye(n), np.diag(r)
    return np.block([[eye, dg], [dg, eye]])


def _corr_matrix(corr, rho):
    """Reference infrastructure.py:153-206: ``corr`` overrides ``rho``; either
    may be time-dependent.  Returns None, an array, or a callable."""
    if corr is None and rho is None:
        return None
    if corr is not None:
        corr = _variable_param_setup(corr)
        cs = _param_shape(corr)
        if cs is not None and (len(cs) not in (2, 3) or cs[0] != cs[1] or
                               (len(cs) == 3 and cs[2] != 1)):
            raise ValueError(
                'the correlation matrix ``corr`` should be square, possibly '
                'with a trailing 1-dimensional axis matching the paths axis, '
                'not an array with shape {}'.format(cs))
        return corr
    rho = _variable_param_setup(rho)
    rs = _param_shape(rho)
    if rs is not None and (len(rs) > 2 or (len(rs) == 2 and rs[1] != 1)):
        raise ValueError(
            'correlation ``rho`` should be a vector, possibly with a trailing '
            '1-dimensional axis matching the paths axis, not an array with '
            'shape {}'.format(rs))
    if callable(rho):
        def corr_t(t, _rho=rho):
            return _rho_to_corr(_rho(t))
        corr_t.shape = (None if rs is None else (2, 2) if rs == () else
                        (2*rs[0], 2*rs[0]))
        return corr_t
    return _rho_to_corr(rho)


def _check_source(src, paths, vshape):
    """Reference infrastructure.py:209-232."""
    if callable(src) and hasattr(src, 'paths') and hasattr(src, 'vshape'):
        ok = (src.paths == paths)
        try:
            np.broadcast_to(np.empty(src.vshape), vshape)
        except ValueError:
            ok = False
        if not ok:
            raise ValueError(
                'invalid stochasticity source: expecting source paths={} and '
                'vshape broadcastable to {}, but paths={}, vshape={} were '
                'found'.format(paths, vshape, src.paths, src.vshape))
        return
    raise ValueError(
        "stochasticity source of type '{}', not compliant with the source "
        'protocol (should be callable with properly defined paths and vshape '
        'attributes)'.format(type(src).__name__))


def _source_setup(dz, source_type, paths, vshape, **args):
    """Reference infrastructure.py:235-243: None -> default class, a class ->
    instantiate, an instance -> validate and use as is."""
    if dz is None:
        return source_type(paths=paths, vshape=vshape, **args)
    if inspect.isclass(dz):
        return dz(paths=paths, vshape=vshape, **args)
    _check_source(dz, paths, vshape)
    return dz


_empty = inspect.Signature.empty


_SIGNATURES = {}


def _signature(f):
    """[(name, default)] of f's parameters; cached per function (a process is
    constructed at every Monte Carlo batch: inspect.signature is 40 us a call)."""
    key = (getattr(f, '__func__', f), getattr(f, '__self__', None) is not None)
    try:
        sig = _SIGNATURES.get(key)
    except TypeError:                    # unhashable callable
        key, sig = None, None
    if sig is None:
        sig = tuple((k, p.default) for k, p in inspect.signature(f).parameters.items())
        if key is not None:
            _SIGNATURES[key] = sig
    return list(sig)


# --------------------------------------------------------------------------
# process (host) and device_process (HBM)
# --------------------------------------------------------------------------

class process(np.ndarray):
    """Host container of a realised process: ``x[i]`` are the values at time
    ``t[i]``, shaped ``vshape + (paths,)`` (reference infrastructure.py:
    423-456).  Only construction, interpolation-as-source and the
    across-path summaries are provided; everything else is plain ndarray."""

    __array_priority__ = 1.0
    interp_kind = 'linear'

    def __new__(cls, t=0., *, x=None, v=None, c=None, dtype=None):
        t = np.asarray(t)
        if t.ndim > 1 or t.size == 0:
            raise ValueError('the shape of a process timeline should be () '
                             'or (n,), not {}'.format(t.shape))
        t = t.reshape(-1)
        if sum(z is not None for z in (x, v, c)) != 1:
            raise ValueError('when creating a process instance, one and only '
                             'one of x or v or c should be provided')
        if x is not None:
            x = np.asarray(x, dtype=dtype)
        elif v is not None:
            x = np.asarray(v, dtype=dtype)[..., np.newaxis]
        else:
            c = np.asarray(c, dtype=dtype)
            x = np.empty(t.shape + c.shape + (1,), dtype=dtype)
            x[...] = c[np.newaxis, ..., np.newaxis]
        if t.shape != x.shape[:1]:
            raise ValueError('process could not be created from timeline t '
                             'shaped {} and body shaped {}'
                             .format(t.shape, x.shape))
        obj = x.view(cls)
        obj.t = t
        return obj

    def __array_finalize__(self, obj):
        t = getattr(obj, 't', None)
        self.t = t if (t is not None and t.shape == self.shape[:1]) else None

    def _is_compatible(self, other):
        """Two processes broadcast only if they share the timeline (or one is
        constant) and their values and paths broadcast axis by axis, vshapes
        of equal length (reference infrastructure.py:458-471)."""
        t_ok = (self.t.size == 1 or other.t.size == 1
                or np.array_equal(self.t, other.t))
        s1, s2 = self.shape[1:], other.shape[1:]
        return t_ok and len(s1) == len(s2) and all(
            a == b or a == 1 or b == 1 for a, b in zip(s1, s2))

    def __array_wrap__(self, out, context=None, return_scalar=False):
        """ufunc results stay processes, on the common non-constant timeline of
        the process operands, else the constant timeline of the first one;
        incompatible process operands raise ValueError (reference
        infrastructure.py:486-540: its __array_prepare__ check, which NumPy 2
        no longer calls, is made here); ndarray methods (sum, mean ...) return
        plain arrays."""
        if context is None:
            return np.asarray(out)
        procs = [a for a in context[1] if isinstance(a, process)]
        for a in procs:
            if getattr(a, 't', None) is None:
                raise ValueError(
                    'cannot operate on a process without a timeline. '
                    'if this results from array operations on processes, '
                    'try using their array views instead (x attribute)')
        for a in procs:
            if not a._is_compatible(self):
                raise ValueError(
                    'processes could not be broadcast together due to '
                    'incompatible shapes {}, {} and/or timelines'
                    .format(a.shape, self.shape))
        t = procs[0].t if procs else self.t
        for a in procs[1:]:
            if a.t.size > 1:
                t = a.t
                break
        out = np.asarray(out)
        if t is None or t.shape != out.shape[:1]:
            return out
        res = out.view(type(self))
        res.t = t
        return res

    # ---- process-specific indexing (reference infrastructure.py:653-707) ---
    def __getitem__(self, key):
        """Plain ndarray indexing returns plain arrays; ``p['t', i]``,
        ``p['p', i]`` and ``p['v', i, j ...]`` index the timeline, the paths
        and the values and return processes (an integer index keeps its axis)."""
        x = self.view(np.ndarray)
        if isinstance(key, str):
            key = (key,)
        if not (isinstance(key, tuple) and key and isinstance(key[0], str)):
            return x[key]
        mode, idx = key[0], key[1:]
        if len(idx) == 1 and isinstance(idx[0], tuple):
            idx = idx[0]
        if mode == 'v':
            return type(self)(self.t, x=x[(slice(None),) + idx + (slice(None),)])
        if mode not in ('t', 'p'):
            raise IndexError('process indexing error - unsupported indexing '
                             'mode ' + repr(mode))
        if len(idx) > 1:
            raise IndexError("process indexing error - one index expected in "
                             "'t' and 'p' modes")
        if idx and isinstance(idx[0], (int, np.integer)):
            i = int(idx[0])
            idx = (slice(i, None) if i == -1 else slice(i, i + 1),)
        if mode == 't':
            return type(self)(self.t[idx], x=x[idx])
        return type(self)(self.t, x=x[(Ellipsis,) + idx])

    def rebase(self, t, *, kind=None):
        """Process with timeline ``t`` and interpolated values (635-651)."""
        t = np.asarray(t)
        t = t.reshape(1) if t.ndim == 0 else t
        return process(t, x=self(t, kind=kind))

    def shapeas(self, vshape_or_process):
        """Add / remove leading 1-axes of the values so as to broadcast against
        values of the given shape (769-801)."""
        vshape = (vshape_or_process.vshape if isinstance(vshape_or_process, process)
                  else _shape_setup(vshape_or_process))
        k, h = self.ndim - 2, len(vshape)
        if h >= k:
            newshape = self.shape[:1] + (1,)*(h - k) + self.shape[1:]
        else:
            if set(self.shape[1:k - h + 1]) != {1}:
                raise ValueError('could not reshape {} process values as {}'
                                 .format(self.vshape, vshape))
            newshape = self.shape[:1] + self.shape[k - h + 1:]
        return type(self)(self.t, x=self.view(np.ndarray).reshape(newshape))

    def pcopy(self, **args):
        return type(self)(t=self.t.copy(**args), x=self.view(np.ndarray).copy(**args))

    def xcopy(self, **args):
        return type(self)(t=self.t, x=self.view(np.ndarray).copy(**args))

    def tcopy(self, **args):
        return type(self)(t=self.t.copy(**args), x=self.view(np.ndarray))

    @property
    def x(self):
        return self.view(np.ndarray)

    @property
    def paths(self):
        return self.shape[-1]

    @property
    def vshape(self):
        return self.shape[1:-1]

    def interp(self, *, kind=None):
        """Callable ``f(s)`` interpolating the values in time (reference
        infrastructure.py:544-613)."""
        import scipy.interpolate
        kind = self.interp_kind if kind is None else kind
        t, x = self.t, self.x
        if t.size == 1:
            def f(s):
                return np.broadcast_to(x[0], np.asarray(s).shape + x.shape[1:]).copy()
            return f
        g = scipy.interpolate.interp1d(
            t, x, axis=0, kind=kind, assume_sorted=True, copy=False,
            bounds_error=False, fill_value=(x[0], x[-1]))
        return lambda s: g(s).astype(x.dtype, copy=False)

    def __call__(self, s, ds=None, *, kind=None):
        """``p(s)`` values at ``s``; ``p(s, ds)`` increments -- which makes a
        process a valid stochasticity source (reference infrastructure.py:
        615-633)."""
        f = self.interp(kind=kind)
        s = np.asarray(s)
        if ds is None:
            return f(s)
        return f(s + np.asarray(ds)) - f(s)

    def _summary(self, name, **kw):
        if 'dtype' in kw and kw['dtype'] is None:
            kw['dtype'] = self.dtype       # reference 851-889: accumulate in own dtype
        return process(t=self.t, x=getattr(self.x, name)(axis=-1, keepdims=True, **kw))

    def _at(self, t):
        if t is None:
            return self.t, self.x
        t = np.asarray(t)
        return t, self(t)

    def chf(self, t=None, u=None):
        """Path-average of exp(1j*u*p(t)), shape t.shape + u.shape + vshape
        (reference infrastructure.py:1125-1167); ``p.chf(u)`` uses the
        process timeline."""
        if t is None and u is None:
            raise TypeError('u argument missing')
        if u is None:
            t, u = None, t
        t, x = self._at(t)
        u = np.asarray(u)
        uu = u.reshape((1,)*t.ndim + u.shape + (1,)*(x.ndim - t.ndim))
        xx = x.reshape(t.shape + (1,)*u.ndim + x.shape[t.ndim:])
        return np.exp(1j*uu*xx).mean(axis=-1)

    def cdf(self, t=None, x=None):
        """Fraction of paths with p(t) <= x, shape t.shape + x.shape + vshape
        (reference infrastructure.py:1169-1209)."""
        if t is None and x is None:
            raise TypeError('x argument missing')
        if x is None:
            t, x = None, t
        t, y = self._at(t)
        x = np.asarray(x)
        xx = x.reshape((1,)*t.ndim + x.shape + (1,)*(y.ndim - t.ndim))
        yy = y.reshape(t.shape + (1,)*x.ndim + y.shape[t.ndim:])
        return (yy <= xx).sum(axis=-1)/self.paths

    def psum(self, dtype=None, out=None):
        return self._summary('sum', dtype=dtype, out=out)

    def pmean(self, dtype=None, out=None):
        return self._summary('mean', dtype=dtype, out=out)

    def pvar(self, dtype=None, out=None, ddof=0):
        return self._summary('var', dtype=dtype, out=out, ddof=ddof)

    def pstd(self, dtype=None, out=None, ddof=0):
        return self._summary('std', dtype=dtype, out=out, ddof=ddof)

    def pmin(self, out=None):
        return self._summary('min', out=out)

    def pmax(self, out=None):
        return self._summary('max', out=out)

    # ---- timeline helpers, summaries across values and along time ---------
    # (reference infrastructure.py:731-766, 894-1122)
    @property
    def tx(self):
        return self.t.reshape(self.t.shape + (1,)*(self.ndim - 1))

    @property
    def dt(self):
        return np.diff(self.t)

    @property
    def dtx(self):
        dt = self.dt
        return dt.reshape(dt.shape + (1,)*(self.ndim - 1))

    def _vsummary(self, name, **kw):
        axes = tuple(range(1, self.ndim - 1))
        return process(t=self.t, x=getattr(np.asarray(self), name)(axis=axes, **kw))

    def _tsummary(self, name, **kw):
        return process(t=self.t[:1],
                       x=getattr(np.asarray(self), name)(axis=0, keepdims=True, **kw))

    def vmin(self, out=None):
        return self._vsummary('min', out=out)

    def vmax(self, out=None):
        return self._vsummary('max', out=out)

    def vsum(self, dtype=None, out=None):
        return self._vsummary('sum', dtype=dtype, out=out)

    def vmean(self, dtype=None, out=None):
        return self._vsummary('mean', dtype=dtype, out=out)

    def vvar(self, dtype=None, out=None, ddof=0):
        return self._vsummary('var', dtype=dtype, out=out, ddof=ddof)

    def vstd(self, dtype=None, out=None, ddof=0):
        return self._vsummary('std', dtype=dtype, out=out, ddof=ddof)

    def tmin(self, out=None):
        return self._tsummary('min', out=out)

    def tmax(self, out=None):
        return self._tsummary('max', out=out)

    def tsum(self, dtype=None, out=None):
        return self._tsummary('sum', dtype=dtype, out=out)

    def tmean(self, dtype=None, out=None):
        return self._tsummary('mean', dtype=dtype, out=out)

    def tvar(self, dtype=None, out=None, ddof=0):
        return self._tsummary('var', dtype=dtype, out=out, ddof=ddof)

    def tstd(self, dtype=None, out=None, ddof=0):
        return self._tsummary('std', dtype=dtype, out=out, ddof=ddof)

    def tcumsum(self, dtype=None, out=None):
        return process(t=self.t, x=np.asarray(self).cumsum(axis=0, dtype=dtype, out=out))

    def tdiff(self, dt_exp=0, fwd=True):
        """``q[i] = (p[i+1] - p[i])/(t[i+1] - t[i])**dt_exp`` at ``t[i]``
        (forward) or ``t[i+1]`` (reference 1032-1080)."""
        x = np.diff(np.asarray(self), axis=0)
        if dt_exp:
            x = x/self.dtx**dt_exp
        return process(t=self.t[:-1] if fwd else self.t[1:], x=x)

    def tder(self):
        return self.tdiff(dt_exp=1)

    def tint(self):
        """Running integral of the piecewise-constant (left value) process
        (reference 1101-1122)."""
        x = np.zeros(self.shape)
        x[1:] = np.asarray(self)[:-1]*self.dtx
        return process(t=self.t, x=x.cumsum(axis=0))


def piecewise(t=0., *, x=None, v=None, dtype=None, mode='mid'):
    """Process interpolating to a piecewise constant function: ``t[i]`` is the
    midpoint ('mid'), start ('forward') or end ('backward') of the segment
    with value ``x[i]`` (reference infrastructure.py:1216-1281).  Usable as a
    time-dependent SDE parameter (nearest-neighbour interpolation)."""
    p = process(t, x=x, v=v, dtype=dtype)
    t, x = p.t, p.x
    if mode == 'mid':
        s, y = t, x
    elif mode in ('forward', 'backward'):
        # every knot doubled: the value switches AT the knot
        s = np.repeat(t, 2)
        y = np.repeat(x, 2, axis=0)
        if mode == 'forward':
            y[2::2] = x[:-1]          # left copy of knot i carries value i-1
        else:
            y[1:-1:2] = x[1:]         # right copy of knot i carries value i+1
    else:
        raise ValueError("mode should be one of 'mid', 'forward', 'backward', "
                         'but {} was given'.format(mode))
    p = process(s, x=y, dtype=dtype)
    p.interp_kind = 'nearest'
    return p


class device_process:
    """A process resident in HBM: ``x`` is a CUDA float64 tensor shaped
    ``(N,) + vshape + (paths,)`` (paths contiguous, the reference's layout,
    integration.py:550), ``t`` the host timeline.  ``pmean/pvar/pstd`` run the
    reduction kernels and return small host ``process`` objects shaped
    ``(N,) + vshape + (1,)`` as the reference does (infrastructure.py:
    861-889)."""

    def __init__(self, t, x):
        self.t = np.asarray(t).reshape(-1)
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64):
            raise TypeError('device_process needs a CUDA float64 tensor')
        if x.shape[0] != self.t.size:
            raise ValueError('process could not be created from timeline t '
                             'shaped {} and body shaped {}'
                             .format(self.t.shape, tuple(x.shape)))
        self.x = x

    shape = property(lambda self: tuple(self.x.shape))
    paths = property(lambda self: self.x.shape[-1])
    vshape = property(lambda self: tuple(self.x.shape[1:-1]))
    ndim = property(lambda self: self.x.dim())
    dtype = property(lambda self: np.dtype(float))

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, key):
        """Integer / slice on the time axis keep the container; anything else
        returns the indexed tensor."""
        if isinstance(key, slice):
            return device_process(self.t[key], self.x[key])
        return self.x[key]

    def cpu(self):
        return process(t=self.t, x=self.x.cpu().numpy())

    to_process = cpu

    def __array__(self, dtype=None, copy=None):
        a = self.x.cpu().numpy()
        return a if dtype is None else a.astype(dtype)

    def _rows(self):
        x = self.x if self.x.is_contiguous() else self.x.contiguous()
        return x.reshape(-1, x.shape[-1])

    def _moments(self, centre=None):
        return _cuda.moments(self._rows(), self.paths, centre)

    def _wrap(self, flat):
        return process(t=self.t, x=np.asarray(flat).reshape(self.shape[:-1] + (1,)))

    # ---- interpolation in time, cdf, chf (reference 544-633, 1125-1209) ----
    def _bracket(self, t):
        """(i_lo, i_hi, interp) of scipy.interpolate.interp1d(kind='linear',
        fill_value=(x[0], x[-1])): knots lo = searchsorted(t)-1 clipped; no
        extrapolation."""
        tk = self.t
        if tk.size == 1 or t < tk[0]:
            return 0, 0, 0
        if t > tk[-1]:
            return tk.size - 1, tk.size - 1, 0
        idx = int(np.clip(np.searchsorted(tk, t), 1, tk.size - 1))
        return idx - 1, idx, 1

    def _weights(self, lo, hi, t):
        """Interpolation weights of scipy's interp1d._call_linear."""
        t_lo, t_hi = self.t[lo], self.t[hi]
        if lo == hi:
            return 1., 0.
        t = np.float64(t)
        return float((t_hi - t)/(t_hi - t_lo)), float((t - t_lo)/(t_hi - t_lo))

    def _rows2d(self):
        x = self.x if self.x.is_contiguous() else self.x.contiguous()
        return x.reshape(x.shape[0], -1, x.shape[-1])       # [N, V, paths]

    def __call__(self, s, ds=None):
        """Interpolated values ``p(s)`` (or increments ``p(s, ds)``) as a CUDA
        tensor shaped ``s.shape + vshape + (paths,)``."""
        s = np.asarray(s, dtype=float)
        if ds is not None:
            return self(s + np.asarray(ds, dtype=float)) - self(s)
        rows = self._rows2d()
        dev = rows.device
        out = _cuda.empty((s.size,) + tuple(rows.shape[1:]), dev)
        with torch.cuda.device(dev):
            for k, t in enumerate(s.reshape(-1)):
                lo, hi, interp = self._bracket(float(t))
                if not interp:
                    out[k] = rows[lo]
                    continue
                w_lo, w_hi = self._weights(lo, hi, t)
                for v in range(rows.shape[1]):
                    _lib.check(_lib.lib.sdeb_path_interp(
                        _cuda.ptr(rows[lo, v]), _cuda.ptr(rows[hi, v]), w_lo, w_hi,
                        self.paths, _cuda.ptr(out[k, v]), _cuda.stream_ptr(dev)))
        return out.reshape(s.shape + self.vshape + (self.paths,))

    def _eval(self, t, q, mode):
        rows = self._rows2d()
        dev = rows.device
        q = np.asarray(q, dtype=float)
        tq = self.t if t is None else np.asarray(t, dtype=float)
        nq, V = q.size, rows.shape[1]
        qd = _cuda.to_device(q.reshape(-1), dev)
        if mode == 'cdf':
            acc = _cuda.zeros((tq.size, V, nq), dev, torch.int64)
        else:
            acc = _cuda.empty((tq.size, V, nq, 2), dev)
            ws_bytes = _lib.lib.sdeb_path_eval_workspace(nq)
            ws = _cuda.empty((ws_bytes//8,), dev)
        with torch.cuda.device(dev):
            for k, tv in enumerate(tq.reshape(-1)):
                if t is None:
                    lo, hi, interp = k, k, 0
                else:
                    lo, hi, interp = self._bracket(float(tv))
                w_lo, w_hi = self._weights(lo, hi, tv)
                for v in range(V):
                    args = (_cuda.ptr(rows[lo, v]), _cuda.ptr(rows[hi, v]), w_lo, w_hi,
                            interp, self.paths, _cuda.ptr(qd), nq)
                    if mode == 'cdf':
                        _lib.check(_lib.lib.sdeb_path_cdf(*args, _cuda.ptr(acc[k, v]),
                                                          _cuda.stream_ptr(dev)))
                    else:
                        _lib.check(_lib.lib.sdeb_path_chf(*args, _cuda.ptr(acc[k, v]),
                                                          _cuda.ptr(ws), ws_bytes,
                                                          _cuda.stream_ptr(dev)))
        a = acc.cpu().numpy()
        if mode == 'cdf':
            r = a/self.paths                                  # [T, V, nq]
        else:
            r = (a[..., 0] + 1j*a[..., 1])/self.paths
        r = np.moveaxis(r, 1, -1)                             # [T, nq, V]
        return r.reshape(tq.shape + q.shape + self.vshape)

    def cdf(self, t=None, x=None):
        """Fraction of paths with ``p(t) <= x`` (reference 1169-1209)."""
        if t is None and x is None:
            raise TypeError('x argument missing')
        if x is None:
            t, x = None, t
        return self._eval(t, x, 'cdf')

    def chf(self, t=None, u=None):
        """Path-average of ``exp(1j*u*p(t))`` (reference 1125-1167)."""
        if t is None and u is None:
            raise TypeError('u argument missing')
        if u is None:
            t, u = None, t
        return self._eval(t, u, 'chf')

    def psum(self):
        return self._wrap(self._moments()[:, 0])

    def _two_pass(self):
        """(first-pass mean, power sums about it): two HBM passes -- sum / min /
        max, then the centred sums with the centre taken from the first pass ON
        THE DEVICE -- and one device-to-host copy (two-pass accuracy of
        numpy.mean / numpy.var, reference infrastructure.py:861-889)."""
        rows, n = self._rows(), self.paths
        rs_parts, st_parts = [], []
        for r0 in range(0, rows.shape[0], 32768):
            part = rows[r0:r0 + 32768]
            rs = _cuda.mc_range(part, n)
            st_parts.append(_cuda.mc_update(part, n, range_stats=rs))
            rs_parts.append(rs)
        both = torch.cat(rs_parts + st_parts).cpu().numpy()
        k = rows.shape[0]
        return both[:k, 0]/n, both[k:]

    def pmean(self):
        c, m2 = self._two_pass()
        return self._wrap(c + m2[:, 0]/self.paths)

    def pvar(self, ddof=0):
        n = self.paths
        c, m = self._two_pass()
        d = m[:, 0]/n
        return self._wrap((m[:, 1] - n*d*d)/(n - ddof))

    def pstd(self, ddof=0):
        return self._wrap(np.sqrt(np.asarray(self.pvar(ddof=ddof)).reshape(-1)))

    def pmin(self):
        return self._wrap(self._moments()[:, 4])

    def pmax(self):
        return self._wrap(self._moments()[:, 5])

    # ---- summaries across values and along time, on the slab ---------------
    # (reference infrastructure.py:894-1122): results stay in HBM as
    # device_process objects -- path-dependent payoffs (running maximum,
    # time average, realised variance) never leave the device.
    def _axis_reduce(self, outer, n_reduce, cols, what, ddof=0):
        x = self.x if self.x.is_contiguous() else self.x.contiguous()
        dev = x.device
        outs = {k: None for k in ('min', 'max', 'sum', 'ssd')}
        key = {'mean': 'sum', 'var': 'ssd', 'std': 'ssd'}.get(what, what)
        outs[key] = _cuda.empty((outer*cols,), dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.sdeb_axis_reduce(
                _cuda.ptr(x), outer, n_reduce, cols,
                *(None if outs[k] is None else _cuda.ptr(outs[k])
                  for k in ('min', 'max', 'sum', 'ssd')),
                float(n_reduce if what == 'mean' else 1), float(n_reduce - ddof),
                int(what == 'std'), _cuda.stream_ptr(dev)))
        return outs[key]

    def _tsummary(self, what, ddof=0):
        n, cols = self.x.shape[0], int(np.prod(self.shape[1:], dtype=int))
        r = self._axis_reduce(1, n, cols, what, ddof)
        return device_process(self.t[:1], r.reshape((1,) + self.shape[1:]))

    def _vsummary(self, what, ddof=0):
        n, V = self.x.shape[0], int(np.prod(self.vshape, dtype=int))
        r = self._axis_reduce(n, V, self.paths, what, ddof)
        return device_process(self.t, r.reshape((n, self.paths)))

    def tmin(self):
        return self._tsummary('min')

    def tmax(self):
        return self._tsummary('max')

    def tsum(self):
        return self._tsummary('sum')

    def tmean(self):
        return self._tsummary('mean')

    def tvar(self, ddof=0):
        return self._tsummary('var', ddof)

    def tstd(self, ddof=0):
        return self._tsummary('std', ddof)

    def vmin(self):
        return self._vsummary('min')

    def vmax(self):
        return self._vsummary('max')

    def vsum(self):
        return self._vsummary('sum')

    def vmean(self):
        return self._vsummary('mean')

    def vvar(self, ddof=0):
        return self._vsummary('var', ddof)

    def vstd(self, ddof=0):
        return self._vsummary('std', ddof)

    def _scan(self, mode, weights, rows_out):
        x = self.x if self.x.is_contiguous() else self.x.contiguous()
        dev = x.device
        n, cols = x.shape[0], int(np.prod(self.shape[1:], dtype=int))
        out = _cuda.empty((rows_out,) + self.shape[1:], dev)
        w = None if weights is None else _cuda.to_device(
            np.ascontiguousarray(weights, dtype=float), dev)
        if rows_out:
            with torch.cuda.device(dev):
                _lib.check(_lib.lib.sdeb_time_scan(
                    _cuda.ptr(x), n, cols, mode, None if w is None else _cuda.ptr(w),
                    _cuda.ptr(out), _cuda.stream_ptr(dev)))
        return out

    def tcumsum(self):
        return device_process(self.t, self._scan(_lib.SCAN_CUMSUM, None, len(self)))

    def tdiff(self, dt_exp=0, fwd=True):
        w = np.diff(self.t)**dt_exp if dt_exp else None
        return device_process(self.t[:-1] if fwd else self.t[1:],
                              self._scan(_lib.SCAN_DIFF, w, len(self) - 1))

    def tder(self):
        return self.tdiff(dt_exp=1)

    def tint(self):
        return device_process(self.t, self._scan(_lib.SCAN_INT, np.diff(self.t), len(self)))


# --------------------------------------------------------------------------
# sources
# --------------------------------------------------------------------------

class source:
    """Base class of stochasticity sources (reference infrastructure.py:
    1291-1382).  Any callable ``dz(t, dt)`` with ``paths`` and ``vshape``
    attributes obeys the source protocol."""

    def __init__(self, *, paths=1, vshape=(), dtype=None, rng=None, seed=None):
        self.paths, self.dtype = paths, dtype
        self.vshape = _shape_setup(vshape)
        if isinstance(rng, (int, np.int64, list, np.ndarray)):
            raise TypeError('`rng` must be an instance of a random number '
                            'generator, not {}.'.format(type(rng)))
        self._rng = default_rng if rng is None else rng
        # Philox key: explicit seed, or 64 bits drawn from the generator
        self.seed = int(seed) if seed is not None else _seed_from(self._rng)
        self._epoch = 0      # bumped per integration / standalone call

    def __call__(self, t, dt=None):
        dt = 0 if dt is None else dt
        return np.asarray(t) + np.asarray(dt) + np.nan

    @property
    def rng(self):
        return self._rng

    @property
    def size(self):
        return 0

    @property
    def t(self):
        return np.array((), dtype=float)

    def next_key(self):
        """Fresh 64-bit Philox key for one integration (or one draw call):
        successive runs of the same instance are independent, the same
        (seed, call index) always reproduces the same stream."""
        # seed and call index are hashed as SEPARATE words: with seed + epoch,
        # call n of seed s would reuse the key of call 0 of seed s + n, i.e.
        # `for seed in range(k)` loops with repeated calls would repeat paths
        key = _splitmix64(_splitmix64(self.seed & 0xFFFFFFFFFFFFFFFF) ^
                          ((self._epoch*0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF))
        self._epoch += 1
        return key


def _chol(corr):
    corr = np.asarray(corr, dtype=float)
    if corr.ndim == 3:
        if corr.shape[2] != 1:
            raise ValueError('invalid correlation matrix shape {}'
                             .format(corr.shape))
        corr = corr[..., 0]
    try:
        return np.linalg.cholesky(corr)
    except np.linalg.LinAlgError:
        # semi-definite matrices (e.g. |rho| = 1): symmetric square root made
        # triangular by QR keeps L L^T = corr
        w, v = np.linalg.eigh(corr)
        if w.min() < -1e-10:
            raise ValueError('correlation matrix is not positive semidefinite')
        a = v*np.sqrt(np.clip(w, 0, None))
        q, r = np.linalg.qr(a.T)
        L = r.T
        return L*np.sign(np.diag(L) + (np.diag(L) == 0))


class wiener_source(source):
    """dw: standard Wiener increments, optionally correlated along the last
    axis of ``vshape`` (reference infrastructure.py:1388-1560).  Inside an
    integration the draws are generated in-kernel (Philox4x32-10 -> Box-Muller
    -> Cholesky factor of ``corr`` evaluated at ``t + dt/2``); as a standalone
    callable ``dw(t, dt)`` launches ``sdeb_draw_wiener`` and returns a host
    array ``vshape + (paths,)``."""

    def __init__(self, *, paths=1, vshape=(), dtype=None, rng=None,
                 corr=None, rho=None, seed=None):
        super().__init__(paths=paths, vshape=vshape, dtype=dtype, rng=rng, seed=seed)
        self.corr = corr = _corr_matrix(corr, rho)
        cshape = _param_shape(corr)
        if corr is not None:
            if self.vshape == ():
                raise ValueError('if vshape is (), no correlations apply, but '
                                 'corr={}, rho={} were given'.format(corr, rho))
            if cshape is not None and (
                    cshape[:2] != 2*self.vshape[-1:] or
                    (len(cshape) == 3 and cshape[-1] != 1)):
                raise ValueError(
                    'cannot instantiate a Wiener source with values shape {} '
                    'and correlation matrix shape {}'.format(self.vshape, cshape))

    def chol_at(self, t):
        """Lower Cholesky factor of the correlation in force at time ``t``
        (None if uncorrelated)."""
        if self.corr is None:
            return None
        c = self.corr(t) if callable(self.corr) else self.corr
        return _chol(c)

    def __call__(self, t, dt):
        t, dt = np.broadcast_arrays(t, dt)
        if t.shape != ():
            out = np.empty(t.shape + self.vshape + (self.paths,), dtype=self.dtype)
            for i in np.ndindex(t.shape):
                out[i] = wiener_source.__call__(self, t[i], dt[i])
            return out
        dev = _cuda.device()
        vs = self.vshape
        correlated = self.corr is not None
        ndw = vs[-1] if correlated else 1
        groups = int(np.prod(vs[:-1] if correlated else vs, dtype=int))
        if ndw > 32:
            raise NotImplementedError('more than 32 correlated components')
        L = self.chol_at(float(t) + float(dt)/2)
        chol = None
        if L is not None:
            chol = _cuda.to_device(L[np.tril_indices(ndw)], dev)
        out = _cuda.empty((groups*ndw, self.paths), dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.sdeb_draw_wiener(
                _cuda.ptr(out), groups, ndw, self.paths, self.paths, 0,
                self.next_key(), 0, float(np.sqrt(np.abs(dt))),
                _cuda.ptr(chol), _cuda.stream_ptr(dev)))
        return out.cpu().numpy().reshape(vs + (self.paths,)).astype(self.dtype, copy=False)


def _over_times(src, call, t, dt, shape, dtype, extras=()):
    """Source call vectorised over array-valued ``t, dt`` (output shaped
    ``t.shape + vshape + (paths,)``, reference infrastructure.py:1321-1330);
    ``extras`` names attributes (``dn_value`` ...) stacked the same way."""
    t, dt = np.broadcast_arrays(t, dt)
    out = np.empty(t.shape + shape, dtype=dtype)
    side = {k: [] for k in extras}
    for i in np.ndindex(t.shape):
        out[i] = call(src, t[i], dt[i])
        for k in extras:
            if hasattr(src, k):
                side[k].append(getattr(src, k))
    for k, v in side.items():
        if v and k != 'y_value':
            setattr(src, k, np.stack(v).reshape(t.shape + np.shape(v[0])))
    return out


class poisson_source(source):
    """dn: Poisson increments with intensity ``lam`` (reference
    infrastructure.py:1566-1633): ``sign(dt)*Poisson(|dt|*lam(t+dt/2))``."""

    def __init__(self, *, paths=1, vshape=(), dtype=int, rng=None, lam=1.,
                 seed=None):
        super().__init__(paths=paths, vshape=vshape, dtype=dtype, rng=rng, seed=seed)
        self.lam = lam = _variable_param_setup(lam)
        ls = _param_shape(lam)
        if ls is not None:
            try:
                np.broadcast_to(np.empty(ls), self.vshape + (paths,))
            except ValueError:
                raise ValueError(
                    'cannot broadcast lambda parameter shaped {} to requested '
                    'poisson source shape = vshape + (paths,) = {}'
                    .format(ls, self.vshape + (paths,)))

    def lam_at(self, t):
        return np.asarray(self.lam(t) if callable(self.lam) else self.lam,
                          dtype=float)

    def __call__(self, t, dt):
        if np.ndim(t) or np.ndim(dt):
            return _over_times(self, poisson_source.__call__, t, dt,
                               self.vshape + (self.paths,), self.dtype)
        dj, dn = _draw_cpoisson(self, self, None, t, dt, want_dj=False)
        return dn


class _law:
    """Jump-size law with possibly time-dependent parameters (reference
    infrastructure.py:1640-1776).  ``at(t)`` gives (kind, a, b, pa)."""

    def __init__(self, kind, **params):
        self.kind, self.params = kind, params

    def at(self, t):
        v = {k: (np.asarray(z(t)) if callable(z) else np.asarray(z))
             for k, z in self.params.items()}
        a = np.asarray(v.get('a', 0.), dtype=float)
        b = np.asarray(v.get('b', 0.), dtype=float)
        pa = np.asarray(v.get('pa', 0.), dtype=float)
        if self.kind == _lib.LAW_EXP and (a == 0).any():
            raise ValueError('domain error in arguments')
        if self.kind == _lib.LAW_DOUBLE_EXP and (
                (a <= 0).any() or (b <= 0).any() or (pa > 1).any() or (pa < 0).any()):
            raise ValueError('domain error in arguments')
        return self.kind, a, b, pa

    def rvs(self, size, random_state=None):
        """Host-side variates with the ``scipy.stats`` ``rvs`` signature: the
        protocol through which foreign consumers (e.g. the reference's own
        list-based ``true_cpoisson_source``) draw jump sizes.  The device path
        never calls it: in-kernel draws come from ``jump_size`` (csrc)."""
        if any(callable(z) for z in self.params.values()):
            raise TypeError('time-dependent distribution: evaluate it at a '
                            'time first, y(t).rvs(size)')
        _, a, b, pa = self.at(0.)
        rng = (random_state if hasattr(random_state, 'standard_normal')
               else np.random.default_rng(random_state))
        if self.kind == _lib.LAW_NORMAL:
            return a + b*rng.standard_normal(size)
        if self.kind == _lib.LAW_UNIFORM:
            return a + (b - a)*rng.random(size)
        if self.kind == _lib.LAW_EXP:
            return a*rng.standard_exponential(size)
        plus = a*rng.standard_exponential(size)
        minus = b*rng.standard_exponential(size)
        return np.where(rng.random(size) <= pa, plus, -minus) + 0

    # moments used by analytical formulae (reference 1719, 1727, 1743-1791)
    def _const(self):
        return self.at(0.)[1:]

    def mean(self):
        a, b, pa = self._const()
        return {_lib.LAW_NORMAL: a, _lib.LAW_UNIFORM: (a + b)/2,
                _lib.LAW_EXP: a, _lib.LAW_DOUBLE_EXP: pa*a - (1 - pa)*b}[self.kind] + 0

    def var(self):
        a, b, pa = self._const()
        return {_lib.LAW_NORMAL: b*b, _lib.LAW_UNIFORM: (b - a)**2/12,
                _lib.LAW_EXP: a*a,
                _lib.LAW_DOUBLE_EXP: pa*(1 - pa)*(a + b)**2 + (pa*a**2 + (1 - pa)*b**2)
                }[self.kind] + 0

    def std(self):
        return np.sqrt(self.var())

    def exp_mean(self):
        a, b, pa = self._const()
        if self.kind == _lib.LAW_NORMAL:
            return np.exp(a + b*b/2) + 0
        if self.kind == _lib.LAW_UNIFORM:
            return (np.exp(b) - np.exp(a))/(b - a) + 0
        if self.kind == _lib.LAW_EXP:
            return np.where(a < 1, 1/(1 - a), np.inf) + 0
        return (pa/(1 - a) if a < 1 else np.inf) + (1 - pa)/(1 + b) + 0


class _timed_law(_law):
    """A law with callable parameters is itself a callable ``y(t)`` returning
    the law frozen at ``t``, like the reference's time-dependent distributions
    (infrastructure.py:1640-1650)."""

    def __call__(self, t):
        _, a, b, pa = self.at(t)
        return _law(self.kind, a=a, b=b, pa=pa)


def _make_law(kind, **params):
    timed = any(callable(z) for z in params.values())
    return (_timed_law if timed else _law)(kind, **params)


def norm_rv(a=0, b=1):
    """Normal jump sizes, mean ``a`` and standard deviation ``b``
    (reference infrastructure.py:1653-1664)."""
    return _make_law(_lib.LAW_NORMAL, a=a, b=b)


def uniform_rv(a=0, b=1):
    """Uniform jump sizes in [a, b] (reference 1667-1677)."""
    return _make_law(_lib.LAW_UNIFORM, a=a, b=b)


def exp_rv(a=1):
    """Exponential jump sizes with (signed) scale ``a`` (reference 1680-1693)."""
    return _make_law(_lib.LAW_EXP, a=a)


def double_exp_rv(a=1, b=1, pa=0.5):
    """Double exponential: scale ``a`` with probability ``pa``, ``-b``
    otherwise (reference 1696-1712)."""
    return _make_law(_lib.LAW_DOUBLE_EXP, a=a, b=b, pa=pa)


def rvmap(f, y):
    """Distribution of ``f(y)`` -- or of ``f(t, y(t))`` when ``f``'s first
    argument is named ``t`` or ``s`` and/or ``y`` is time-dependent -- as an
    object with an ``rvs(size, random_state=None)`` method, or a callable of
    ``t`` returning one, as accepted by ``cpoisson_source`` (reference
    infrastructure.py:1799-1862).  Like any user-supplied ``rvs`` object it is
    evaluated on the host over device-drawn counts."""
    timed = _signature(f)[0][0] in ('t', 's')

    class mapped:
        def __init__(self, rv, t=None):
            self._rv, self._t = rv, t

        def rvs(self, size, random_state=None):
            z = self._rv.rvs(size=size, random_state=random_state)
            return f(self._t, z) if timed else f(z)

    if callable(y) or timed:
        return lambda t: mapped(y(t) if callable(y) else y, t)
    return mapped(y)


class cpoisson_source(source):
    """dj: compound Poisson increments (reference infrastructure.py:
    1881-2040).  With a preset law (``norm_rv`` ...) the draws happen
    in-kernel; after a standalone call ``dn_value`` holds the counts."""

    def __init__(self, *, paths=1, vshape=(), dtype=None, rng=None, dn=None,
                 ptype=int, lam=1., y=None, seed=None):
        super().__init__(paths=paths, vshape=vshape, dtype=dtype, rng=rng, seed=seed)
        self.dn = _source_setup(dn, poisson_source, paths=paths,
                                vshape=self.vshape, dtype=ptype, rng=rng, lam=lam)
        self.ptype = self.dn.dtype if hasattr(dn, 'dtype') else ptype
        self.lam = self.dn.lam if hasattr(dn, 'lam') else lam
        self.y = uniform_rv(a=0, b=1) if y is None else y

    def device_ready(self):
        """True when both the counts and the sizes can be drawn in-kernel."""
        return type(self.dn) is poisson_source and isinstance(self.y, _law)

    def __call__(self, t, dt):
        if np.ndim(t) or np.ndim(dt):
            return _over_times(self, cpoisson_source.__call__, t, dt,
                               self.vshape + (self.paths,), self.dtype,
                               extras=('dn_value', 'y_value'))
        if not self.device_ready():
            return self._call_with_user_objects(t, dt)
        dj, dn = _draw_cpoisson(self, self.dn, self.y, t, dt, want_dj=True)
        self.dn_value = dn
        return dj.astype(self.dtype, copy=False)

    def _call_with_user_objects(self, t, dt):
        """A user-supplied ``dn`` source or jump-size object (anything with an
        ``rvs(size, random_state)`` method, e.g. a frozen scipy.stats law) is
        Python code and is evaluated where it lives, on the host, with the
        reference's protocol (infrastructure.py:2002-2040): counts from
        ``self.dn`` (the device Poisson draw unless it is a user source too),
        then one ``rvs`` block of shape ``(paths with j jumps, j)`` per count j.
        An SDE fed by such a source runs through the generic replay path."""
        t, dt = np.broadcast_arrays(t, dt)
        if t.shape != ():
            raise NotImplementedError('array-valued t, dt')
        sign = int(np.sign(dt))
        dn = np.asarray(self.dn(t, dt))
        counts = sign*dn
        rv = self.y(float(t) + float(dt)/2) if callable(self.y) else self.y
        dz = np.zeros(self.vshape + (self.paths,), dtype=float if self.dtype is None else self.dtype)
        y_value = []
        for j in range(1, int(counts.max(initial=0)) + 1):
            index = counts == j
            if index.any():
                size = (int(index.sum()), j)
                try:
                    sample = rv.rvs(size=size, random_state=self._rng)
                except TypeError:
                    import warnings
                    sample = rv.rvs(size=size)
                    warnings.warn(
                        'The use of cpoission_source with distributions not '
                        'accepting a `random_state` keyword argument is '
                        'deprecated, and will not be supported in future '
                        'releases.', DeprecationWarning)
                dz[index] = sign*np.asarray(sample).sum(axis=-1)
                y_value.append(sample)
        self.dn_value, self.y_value = dn, y_value
        return dz


def lane_values(z, lead_shape, what='parameter', paths=None):
    """Broadcast a parameter value against ``vshape + (paths,)`` (the
    reference's convention) and return one scalar per lane, shape ``[lanes]``
    -- or, when the value really varies along the paths axis and ``paths`` is
    given, one row per lane, shape ``[lanes, paths]``."""
    z = np.asarray(z, dtype=float)
    lead = tuple(lead_shape)
    if z.size == 1:                       # scalar: the common case, no broadcasting machinery
        return np.full(int(np.prod(lead, dtype=int)), z.reshape(-1)[0])
    try:
        return np.broadcast_to(z, lead + (1,)).reshape(-1).copy()
    except ValueError:
        pass
    if paths is not None:
        try:
            return np.broadcast_to(z, lead + (paths,)).reshape(-1, paths).copy()
        except ValueError:
            pass
    raise ValueError(
        '{} of shape {} does not broadcast to vshape + (paths,) = {}'
        .format(what, z.shape, lead + (paths if paths is not None else 1,)))


def stack_lane_columns(cols, groups, ncomp):
    """cols: per-lane values, each ``[lanes]`` or ``[lanes, paths]``.  Returns the
    record block ``[groups, ncomp*len(cols)]`` or, if any column is
    path-dependent, ``[groups, ncomp*len(cols), paths]`` (component-major)."""
    pp = [c for c in cols if c.ndim == 2]
    if not pp:
        block = np.stack([c.reshape(groups, ncomp) for c in cols], axis=-1)
        return block.reshape(groups, ncomp*len(cols))
    paths = pp[0].shape[-1]
    full = [np.broadcast_to(c[:, None] if c.ndim == 1 else c, (groups*ncomp, paths))
            .reshape(groups, ncomp, paths) for c in cols]
    block = np.stack(full, axis=2)                       # [G, ncomp, per, paths]
    return block.reshape(groups, ncomp*len(cols), paths)


def _draw_cpoisson(src, dn_src, law, t, dt, want_dj):
    t, dt = np.broadcast_arrays(t, dt)
    if t.shape != ():
        raise NotImplementedError('array-valued t, dt')
    t, dt = float(t), float(dt)
    dev = _cuda.device()
    vs, paths = src.vshape, src.paths
    lanes = int(np.prod(vs, dtype=int))
    lam = lane_values(dn_src.lam_at(t + dt/2), vs, 'lam')
    if law is not None:
        kind, a, b, pa = law.at(t + dt/2)
    else:
        kind, a, b, pa = _lib.LAW_UNIFORM, 0., 1., 0.
    a, b, pa = (lane_values(z, vs, 'jump law parameter') for z in (a, b, pa))
    dj = _cuda.empty((lanes, paths), dev) if want_dj else None
    dn = _cuda.empty((lanes, paths), dev, torch.int64)
    key = src.next_key()
    sign = int(np.sign(dt))
    with torch.cuda.device(dev):
        for i in range(lanes):   # one launch per lane: parameters may differ
            _lib.check(_lib.lib.sdeb_draw_cpoisson(
                _cuda.ptr(dj[i]) if want_dj else None, _cuda.ptr(dn[i]), 1,
                paths, paths, 0, (key + i) & 0xFFFFFFFFFFFFFFFF, 0,
                float(abs(dt)*lam[i]), sign, kind, float(a[i]), float(b[i]),
                float(pa[i]), _cuda.stream_ptr(dev)))
    dn_h = dn.cpu().numpy().reshape(vs + (paths,)).astype(dn_src.dtype, copy=False)
    dj_h = dj.cpu().numpy().reshape(vs + (paths,)) if want_dj else None
    return dj_h, dn_h


# ---- antithetic variants (reference infrastructure.py:2047-2150) ---------

def _check_even(paths):
    if paths % 2:
        raise ValueError('the number of paths for sources with antithetics '
                         'should be even, not {}'.format(paths))


class odd_wiener_source(wiener_source):
    """dw with antithetic paths: the trailing K = paths/2 paths repeat the
    leading K with the sign reversed (reference infrastructure.py:2095-2110).
    In-kernel the second half reuses the Philox stream of path p - K."""
    antithetic = True

    def __init__(self, *, paths=2, vshape=(), dtype=None, rng=None, **args):
        _check_even(paths)
        super().__init__(paths=paths, vshape=vshape, dtype=dtype, rng=rng, **args)

    def __call__(self, t, dt):
        self.paths //= 2
        try:
            z = super().__call__(t, dt)
        finally:
            self.paths *= 2
        return np.concatenate((z, -z), axis=-1)


class even_poisson_source(poisson_source):
    """dn with antithetic paths exposing identical increments (reference
    infrastructure.py:2113-2130)."""
    antithetic = True

    def __init__(self, *, paths=2, vshape=(), dtype=None, rng=None, **args):
        _check_even(paths)
        super().__init__(paths=paths, vshape=vshape, dtype=dtype, rng=rng, **args)

    def __call__(self, t, dt):
        self.paths //= 2
        try:
            z = super().__call__(t, dt)
        finally:
            self.paths *= 2
        return np.concatenate((z, z), axis=-1)


class even_cpoisson_source(cpoisson_source):
    """dj with antithetic paths exposing identical jumps (reference
    infrastructure.py:2133-2150).  Like the reference's wrapper it does not
    expose ``dn_value``."""
    antithetic = True

    def __init__(self, *, paths=2, vshape=(), dtype=None, rng=None, **args):
        _check_even(paths)
        super().__init__(paths=paths, vshape=vshape, dtype=dtype, rng=rng, **args)

    def __call__(self, t, dt):
        self.paths //= 2
        self.dn.paths //= 2
        try:
            z = super().__call__(t, dt)
        finally:
            self.paths *= 2
            self.dn.paths *= 2
        del self.dn_value
        return np.concatenate((z, z), axis=-1)


class true_source(source):
    """Base class of sources with memory (reference infrastructure.py:
    2153-2233): ``s(t)`` is the value of the driving process at ``t``,
    ``s(t, dt)`` its increment, realisations are stored and interpolated /
    extended consistently on later calls; ``s[index]`` is a sub-source sharing
    them."""


class true_wiener_source(true_source):
    """dw with memory (reference infrastructure.py:2182-2499): ``dw(t)`` is the
    realised value at time ``t`` of a Wiener path with ``dw(t0) = z0``, and
    ``dw(t, dt) = dw(t + dt) - dw(t)``; new values are drawn conditionally on
    all previously realised ones (extension beyond the known times, Brownian
    bridge between them -- matrix bridge for time-dependent correlations,
    2475-2499), so the SAME driving path can be integrated on different step
    grids.  The realisations live in HBM; calls return CUDA tensors shaped
    ``t.shape + vshape + (paths,)``."""

    def __init__(self, *, paths=1, vshape=(), dtype=None, rng=None, corr=None,
                 rho=None, rtol='max', t0=0., z0=0., seed=None, device=None):
        super().__init__(paths=paths, vshape=vshape, dtype=dtype, rng=rng, seed=seed)
        self._w = wiener_source(paths=paths, vshape=vshape, dtype=dtype, rng=rng,
                                corr=corr, rho=rho, seed=self.seed)
        self.corr = self._w.corr
        self.rtol = np.finfo(float).resolution if rtol == 'max' else float(rtol)
        self.t0, self.z0 = float(t0), z0
        self._device = device
        self._tlist, self._zlist = [self.t0], None
        self._key = _splitmix64(self.seed)
        self._count = 0

    # lanes: groups x ndw as in wiener_source
    def _geometry(self):
        vs = self.vshape
        if self.corr is not None:
            return int(np.prod(vs[:-1], dtype=int)), vs[-1]
        return int(np.prod(vs, dtype=int)), 1

    def _init(self):
        if self._zlist is None:
            dev = _cuda.device(self._device)
            groups, ndw = self._geometry()
            z0 = np.broadcast_to(np.asarray(self.z0, dtype=float),
                                 self.vshape + (self.paths,)).reshape(groups*ndw, self.paths)
            self._zlist = [_cuda.to_device(z0, dev)]

    def _corr_at(self, t):
        groups, ndw = self._geometry()
        if self.corr is None:
            return np.eye(ndw)
        c = np.asarray(self.corr(t) if callable(self.corr) else self.corr, dtype=float)
        return c[..., 0] if c.ndim == 3 else c

    def _launch(self, w1, w2, M1, M2, cov):
        groups, ndw = self._geometry()
        dev = w1.device
        Ly = _chol(cov) if np.any(cov) else np.zeros((ndw, ndw))
        mats = _cuda.to_device(np.stack((M1, M2, Ly)), dev)
        out = _cuda.empty((groups*ndw, self.paths), dev)
        self._count += 1
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.sdeb_bridge_wiener(
                _cuda.ptr(out), _cuda.ptr(w1), _cuda.ptr(w2), _cuda.ptr(mats), groups,
                ndw, self.paths, self.paths, 0, self._key, self._count,
                _cuda.stream_ptr(dev)))
        return out

    def new_outside(self, w, t, s):
        """Value at ``s`` beyond the known times, given ``w`` at the nearest
        known time ``t``; covariance ``corr((t+s)/2)*|s-t|`` (reference 2460-2469)."""
        ndw = self._geometry()[1]
        eye = np.eye(ndw)
        return self._launch(w, None, eye, 0*eye, self._corr_at((t + s)/2)*abs(s - t))

    def new_inside(self, w1, w2, t1, t2, s):
        """Bridge value at ``t1 < s < t2`` (reference 2475-2499)."""
        t0 = self.t0
        if t2 <= t0:
            w1, w2, t1, t2 = w2, w1, t2, t1
        ndw = self._geometry()[1]
        eye = np.eye(ndw)
        a, b = s - t1, t2 - s
        if callable(self.corr):
            A, B = self._corr_at((t1 + s)/2)*a, self._corr_at((s + t2)/2)*b
            Z = B @ np.linalg.inv(A + B)
            return self._launch(w1, w2, Z, eye - Z, (Z @ A)*np.sign(a))
        z = b/(a + b)
        return self._launch(w1, w2, z*eye, (1 - z)*eye, self._corr_at(s)*abs(z*a))

    def _value(self, s):
        import bisect
        self._init()
        t, z = self._tlist, self._zlist
        k = bisect.bisect_right(t, s)
        if k > 0 and np.isclose(s, t[k - 1], rtol=self.rtol, atol=0.):
            return z[k - 1]
        if k == len(t):
            z.append(self.new_outside(z[-1], t[-1], s)); t.append(s)
            return z[-1]
        if k == 0:
            z.insert(0, self.new_outside(z[0], t[0], s)); t.insert(0, s)
            return z[0]
        z.insert(k, self.new_inside(z[k - 1], z[k], t[k - 1], t[k], s)); t.insert(k, s)
        return z[k]

    def _values(self, s):
        s = np.asarray(s, dtype=float)
        rows = [self._value(float(v)) for v in s.reshape(-1)]
        out = torch.stack(rows) if rows else _cuda.empty((0,), _cuda.device(self._device))
        return out.reshape(s.shape + self.vshape + (self.paths,))

    def device_call(self, t, dt=None):
        """``W(t)`` or ``W(t + dt) - W(t)`` as a CUDA tensor (a
        ``_cuda.device_array``: NumPy functions accept it and pull a host
        copy): what the integration kernels consume when the source is passed
        as ``dw=``, without any round trip through the host."""
        if dt is None:
            return _cuda.as_device_array(self._values(t))
        t, dt = np.broadcast_arrays(t, dt)
        return _cuda.as_device_array(self._values(t + dt) - self._values(t))

    def __call__(self, t, dt=None):
        """Host array, like every source of the reference (the realisations
        themselves stay in HBM)."""
        z = self.device_call(t, dt)
        return _cuda.to_host(z.as_subclass(torch.Tensor)).astype(self.dtype, copy=False)

    def __getitem__(self, index):
        """Sub-source over some components, sharing this one's realisations."""
        return _indexed_true_source(self, index)

    @property
    def size(self):
        return 0 if self._zlist is None else sum(z.numel() for z in self._zlist)

    @property
    def t(self):
        return np.array(self._tlist, dtype=float)


class _indexed_true_source:
    """``s[index]``: the components ``index`` (over the ``vshape`` axes, NumPy
    indexing, ``np.newaxis`` included) of a source with memory, sharing the
    parent's stored realisations -- evaluating the view at a new time realises
    ALL components of the parent there (reference infrastructure.py:2235-2252)."""

    def __init__(self, parent, index):
        self._parent = parent
        self._index = index if isinstance(index, tuple) else (index,)
        self.paths, self.dtype = parent.paths, parent.dtype
        self.vshape = np.empty(parent.vshape + (1,))[self._index].shape[:-1]

    def _pick(self, z, tshape):
        return z[(slice(None),)*len(tshape) + self._index]

    def __call__(self, t, dt=None):
        tshape = np.broadcast(t, 0 if dt is None else dt).shape
        return self._pick(self._parent(t, dt), tshape)

    def device_call(self, t, dt=None):
        tshape = np.broadcast(t, 0 if dt is None else dt).shape
        return self._pick(self._parent.device_call(t, dt), tshape)

    def __getitem__(self, index):
        return _indexed_true_source(self, index)

    size = property(lambda self: self._parent.size)
    t = property(lambda self: self._parent.t)
    rng = property(lambda self: self._parent.rng)


class replay_source:
    """Replay-mode source: a table of pre-drawn increments, one entry per
    integration step, e.g. the increments logged from the reference's own
    sources.  ``table[n]`` must be shaped ``vshape + (paths,)`` (host ndarray
    or CUDA tensor); ``dn`` optionally carries the Poisson counts of a compound
    source (exposed as ``dn_value`` like the reference's cpoisson_source,
    infrastructure.py:2038)."""

    def __init__(self, table, dn=None):
        self.table = table
        self.dn_table = dn
        shape = tuple(table.shape)
        self.paths, self.vshape = shape[-1], shape[1:-1]
        self.steps = shape[0]
        self._i = 0

    def __call__(self, t, dt):
        z = self.table[self._i]
        if self.dn_table is not None:
            self.dn_value = self.dn_table[self._i]
        self._i += 1
        return z


# --------------------------------------------------------------------------
# montecarlo
# --------------------------------------------------------------------------

class montecarlo:
    """Cumulated summary statistics of Monte Carlo samples (reference
    infrastructure.py:2718-3312): mean, centred moments 1..4 (centre = mean of
    the first sample, 2934), and one histogram per component whose bins are
    fixed by the first sample (2999-3004).  Samples may be host arrays, CUDA
    tensors or ``device_process`` rows; the reductions always run on the GPU
    (``sdeb_moments`` / ``sdeb_histogram``)."""

    def __init__(self, sample=None, axis=-1, bins=100, range=None, use='all',
                 dtype=None, ctype=np.int64, device=None):
        self.dtype, self.ctype = dtype, ctype
        self._paths = [0]
        self._bins, self._range, self._use = bins, range, use
        self._device = device
        self._mean = self._moments = self._counts = None
        if sample is not None:
            self.update(sample, axis=axis)

    paths = property(lambda self: self._paths[0])

    @property
    def vshape(self):
        if self._moments is None:
            raise ValueError('no sample data: vshape not defined')
        return self._moments[0].shape

    @property
    def shape(self):
        return self.vshape + (self.paths,)

    def _as_device(self, sample, axis):
        if isinstance(sample, device_process):
            sample = sample.x
        if isinstance(sample, torch.Tensor):
            if not sample.is_cuda:
                sample = sample.to(_cuda.device(self._device))
            x = sample.to(torch.float64)
        else:
            a = np.asarray(sample)
            if a.ndim == 0:
                a = a.reshape(1)
            x = _cuda.to_device(np.moveaxis(a, axis, -1), _cuda.device(self._device),
                                dtype=float)
            axis = -1
        if x.dim() == 0:
            x = x.reshape(1)
        x = x.movedim(axis, -1).contiguous()
        return x

    def update(self, sample, axis=-1):
        """Add a sample (reference infrastructure.py:2869-2922)."""
        if self._use not in ('all', 'even', 'odd'):
            raise ValueError("use must be one of 'all', 'even', 'odd', not {}"
                             .format(self._use))
        x = self._as_device(sample, axis)
        sdtype = (sample.dtype if isinstance(sample, np.ndarray) else
                  np.dtype(str(sample.dtype).replace('torch.', ''))
                  if isinstance(sample, torch.Tensor) else np.dtype(float))
        if self._use != 'all':
            # half-sum / half-difference of antithetic pairs x[k], x[K+k]
            # (reference infrastructure.py:2905-2914)
            if x.shape[-1] % 2:
                raise ValueError(
                    'the sample axis for even or odd antithetics sampling '
                    'should be of even length, but {} was found'
                    .format(x.shape[-1]))
            half = x.shape[-1]//2
            rows_in = x.reshape(-1, x.shape[-1])
            folded = _cuda.empty((rows_in.shape[0], half), x.device)
            with torch.cuda.device(x.device):
                for r0 in range(0, rows_in.shape[0], 32768):
                    r1 = min(rows_in.shape[0], r0 + 32768)
                    _lib.check(_lib.lib.sdeb_antithetic_fold(
                        _cuda.ptr(rows_in[r0:r1]), r1 - r0, half, 2*half, half,
                        1 if self._use == 'even' else -1, _cuda.ptr(folded[r0:r1]),
                        _cuda.stream_ptr(x.device)))
            x = folded.reshape(tuple(x.shape[:-1]) + (half,))
        vshape, m = tuple(x.shape[:-1]), x.shape[-1]
        rows = x.reshape(-1, m)
        first = self.paths == 0
        n = self.paths
        nrow = rows.shape[0]
        if first:
            # reference 2928-2930: a floating sample sets the dtype of the results
            dtype = ((sdtype if sdtype.kind == 'f' else float) if self.dtype is None
                     else self.dtype)
            bins = self._bins
            int_bins = isinstance(bins, (int, np.integer)) and not isinstance(bins, bool)
            if np.dtype(dtype) == np.dtype(float) and (bins is None or int_bins):
                # Two launches, ONE device-to-host copy: pass 1 (sum, min, max:
                # the centring constant = first-sample mean, 2934, and the range
                # of numpy.histogram(range=None), 2999-3004) feeds pass 2 on the
                # device, which builds centre and edges itself.
                nb = 0 if bins is None else int(bins)
                if bins is not None and nb < 1:
                    raise ValueError('`bins` must be positive, when an integer')
                mode, a, b = _lib.MC_EDGES_MINMAX, 0., 0.
                if self._range is not None and nb:
                    a, b = map(float, self._range)
                    if a > b:
                        raise ValueError('max must be larger than min in range parameter.')
                    if not (np.isfinite(a) and np.isfinite(b)):
                        raise ValueError('supplied range of [{}, {}] is not finite'.format(a, b))
                    mode = _lib.MC_EDGES_RANGE
                groups = self._alloc_groups([nb]*nrow, [True]*nrow, rows.device)
                for g in groups:
                    _cuda.mc_range(rows[g['r0']:g['r1']], m, out=g['range'])
                host = self._launch(groups, rows, m, None, True, mode, a, b)
                rs = np.concatenate([h['range'] for h in host])
                st = np.concatenate([h['stats'] for h in host])
                if nb and mode == _lib.MC_EDGES_MINMAX:
                    for i in range(nrow):
                        # a NaN anywhere in the sample makes numpy's range NaN
                        lo_i, hi_i = ((np.nan, np.nan) if np.isnan(rs[i, 0])
                                      else (rs[i, 4], rs[i, 5]))
                        if not (np.isfinite(lo_i) and np.isfinite(hi_i)):
                            raise ValueError('autodetected range of [{}, {}] is not finite'
                                             .format(lo_i, hi_i))
                center = (rs[:, 0]/m).reshape(vshape)
                self._groups = groups       # (the kernel left the centre it used in g['centre'])
                if nb:
                    self._edges, self._uniform = [], [True]*nrow
                    for h in host:
                        self._edges += [h['edges'][i] for i in range(h['edges'].shape[0])]
                    self._bins = np.empty(vshape, dtype=object)
                    for j, i in enumerate(np.ndindex(vshape)):
                        self._bins[i] = self._edges[j]
            else:
                rstats = torch.cat([_cuda.mc_range(rows[r0:r1], m)
                                    for r0, r1 in self._row_chunks(nrow)])
                rs = rstats.cpu().numpy()
                center = (rs[:, 0]/m).reshape(vshape).astype(dtype)
                lo, hi = rs[:, 4], rs[:, 5]
                cflat = np.asarray(center, dtype=float).ravel()
                centred = None
                if isinstance(bins, str):
                    # named estimators need the centred moments of the sample
                    g0 = self._alloc_groups([0]*nrow, [True]*nrow, rows.device, centre=cflat)
                    centred = np.concatenate(
                        [h['stats'] for h in self._launch(g0, rows, m, 'own')])
                if bins is not None:
                    self._setup_bins(vshape, lo, hi, rows=rows, m=m, centred=centred,
                                     integer=np.issubdtype(
                                         getattr(sample, 'dtype', np.dtype(float))
                                         if isinstance(sample, np.ndarray) else float,
                                         np.integer), centre=cflat)
                else:
                    self._groups = self._alloc_groups([0]*nrow, [True]*nrow, rows.device,
                                                      centre=cflat)
                host = self._launch(self._groups, rows, m, 'own')
                st = np.concatenate([h['stats'] for h in host])
            self._center = center
            self._moments = tuple(np.zeros(vshape, dtype=dtype) for _ in range(4))
            self._mean = np.zeros(vshape, dtype=dtype)
        else:
            if vshape != self.vshape:
                raise ValueError('sample of shape {} incompatible with the shape {} of '
                                 'the cumulated data'.format(vshape, self.vshape))
            host = self._launch(self._groups, rows, m, 'own')
            st = np.concatenate([h['stats'] for h in host])
        for k in range(4):
            mk = (st[:, k]/m).reshape(vshape)
            self._moments[k][...] = (n*self._moments[k] + m*mk)/(n + m)
        smean = self._center + (st[:, 0]/m).reshape(vshape)
        self._mean[...] = (n*self._mean + m*smean)/(n + m)
        if self._bins is not None:
            self._read_histogram(m, host)
        self._paths[0] += m

    @staticmethod
    def _row_chunks(nrow, width=32768):
        return [(r0, min(nrow, r0 + width)) for r0 in range(0, nrow, width)]

    def _alloc_groups(self, nbins, uniform, dev, edges=None, centre=None):
        """Partition the rows into runs sharing (nbins, uniform) -- one fused
        launch each.  Everything a launch reads or writes besides the sample
        lives in ONE device buffer per group (range statistics, power sums,
        centre, edges, int64 counters), so that a single device-to-host copy
        brings an update's results back."""
        groups, r0, nrow = [], 0, len(nbins)
        while r0 < nrow:
            r1 = r0 + 1
            while (r1 < nrow and r1 - r0 < 32768 and nbins[r1] == nbins[r0]
                   and uniform[r1] == uniform[r0]):
                r1 += 1
            nb, k = int(nbins[r0]), r1 - r0
            ns = _lib.NSTAT
            sizes = dict(range=k*ns, stats=k*ns, centre=k, edges=k*(nb + 1) if nb else 0,
                         counts=k*nb, outside=k if nb else 0)
            buf = _cuda.zeros((sum(sizes.values()),), dev)
            g = dict(r0=r0, r1=r1, nbins=nb, uniform=bool(uniform[r0]), buf=buf, sizes=sizes)
            at = 0
            for name, size in sizes.items():
                v = buf[at:at + size]
                at += size
                if name in ('counts', 'outside'):
                    v = v.view(torch.int64)
                g[name] = v if size else None
            g['range'], g['stats'] = g['range'].view(k, ns), g['stats'].view(k, ns)
            if nb:
                g['edges'], g['counts'] = g['edges'].view(k, nb + 1), g['counts'].view(k, nb)
                if edges is not None:
                    g['edges'].copy_(torch.from_numpy(np.stack(edges[r0:r1]).astype(float)))
            if centre is not None:
                g['centre'].copy_(torch.from_numpy(np.ascontiguousarray(centre[r0:r1], dtype=float)))
            groups.append(g)
            r0 = r1
        # per-row views (montecarlo.allreduce writes the merged counts back)
        self._counts_dev = [g['counts'][i] for g in groups if g['nbins']
                            for i in range(g['r1'] - g['r0'])]
        self._outside_dev = [g['outside'][i:i + 1] for g in groups if g['nbins']
                             for i in range(g['r1'] - g['r0'])]
        return groups

    @staticmethod
    def _launch(groups, rows, m, centre, from_range=False, mode=_lib.MC_EDGES_GIVEN,
                lo=0., hi=0.):
        """One fused moments + histogram launch per group, then ONE device-to-
        host copy of the group's buffer.  Returns per group a dict of host
        arrays (range, stats, centre, edges, counts, outside)."""
        for g in groups:
            r0, r1 = g['r0'], g['r1']
            _cuda.mc_update(
                rows[r0:r1], m, centre=g['centre'] if centre == 'own' else None,
                centre_out=None if centre == 'own' else g['centre'],
                range_stats=g['range'] if from_range else None,
                lo=lo, hi=hi, edges_mode=mode, edges=g['edges'], nbins=g['nbins'],
                uniform=g['uniform'], counts=g['counts'], outside=g['outside'],
                out=g['stats'])
        host = []
        for g in groups:
            flat = g['buf'].cpu().numpy()
            k, nb, ns = g['r1'] - g['r0'], g['nbins'], _lib.NSTAT
            h, at = {}, 0
            for name, size in g['sizes'].items():
                v = flat[at:at + size]
                at += size
                h[name] = v.view(np.int64) if name in ('counts', 'outside') else v
            h['range'], h['stats'] = h['range'].reshape(k, ns), h['stats'].reshape(k, ns)
            if nb:
                h['edges'], h['counts'] = h['edges'].reshape(k, nb + 1), h['counts'].reshape(k, nb)
            host.append(h)
        return host

    @staticmethod
    def _bin_width(name, row, n, a, b, mom, integer):
        """Bin width of numpy.histogram_bin_edges' named estimators
        (numpy/lib/_histograms_impl.py) for one component of the first sample,
        from device-side statistics: ``mom`` = centred power sums S1..S4 of the
        sdeb_moments pass, quartiles from a device sort (one-off setup work;
        the counting itself is the histogram kernel)."""
        ptp = b - a
        var = mom[1]/n - (mom[0]/n)**2
        std = np.sqrt(max(var, 0.))

        def sturges():
            return ptp/(np.log2(n) + 1.0)

        def fd():
            srt = torch.sort(row).values
            q = []
            for frac in (.75, .25):               # np.percentile, linear interpolation
                pos = frac*(n - 1)
                k = int(np.floor(pos))
                lo_v, hi_v = (float(srt[k]), float(srt[min(k + 1, n - 1)]))
                q.append(lo_v + (hi_v - lo_v)*(pos - k))
            return 2.0*(q[0] - q[1])*n**(-1.0/3.0)

        if name == 'auto':
            w = fd()
            w = min(w, sturges()) if w else sturges()
        elif name == 'fd':
            w = fd()
        elif name == 'sturges':
            w = sturges()
        elif name == 'sqrt':
            w = ptp/np.sqrt(n)
        elif name == 'rice':
            w = ptp/(2.0*n**(1.0/3))
        elif name == 'scott':
            w = (24.0*np.pi**0.5/n)**(1.0/3.0)*std
        elif name == 'doane':
            w = 0.0
            if n > 2:
                sg1 = np.sqrt(6.0*(n - 2)/((n + 1.0)*(n + 3)))
                if std > 0.0:
                    mu = mom[0]/n
                    m3 = mom[2]/n - 3*mu*mom[1]/n + 2*mu**3
                    g1 = m3/std**3
                    w = ptp/(1.0 + np.log2(n) + np.log2(1.0 + np.absolute(g1)/sg1))
        else:
            raise ValueError('{!r} is not a valid estimator for `bins`'.format(name)
                             if name != 'stone' else
                             "the 'stone' estimator is not available on the device")
        if integer and w and w < 1:
            w = 1
        return w

    def _setup_bins(self, vshape, lo, hi, rows=None, m=0, centred=None, integer=False,
                    centre=None):
        bins = self._bins
        nrow = int(np.prod(vshape, dtype=int))
        dev = rows.device if rows is not None else _cuda.device(self._device)
        self._edges, self._uniform = [], []
        if isinstance(bins, str):
            for i in range(nrow):
                if self._range is not None:
                    a, b = map(float, self._range)
                    if a > b:
                        raise ValueError('max must be larger than min in range parameter.')
                    if (a, b) != (float(lo[i]), float(hi[i])):
                        raise NotImplementedError(
                            'named bin estimators with an explicit range are not '
                            'available on the device')
                else:
                    a, b = float(lo[i]), float(hi[i])
                if not (np.isfinite(a) and np.isfinite(b)):
                    raise ValueError('autodetected range of [{}, {}] is not finite'.format(a, b))
                if a == b:
                    a, b = a - 0.5, b + 0.5
                    nb = 1
                else:
                    w = self._bin_width(bins, rows[i], m, a, b, centred[i], integer)
                    nb = int(np.ceil((b - a)/w)) if w else 1
                self._edges.append(np.linspace(a, b, nb + 1))
                self._uniform.append(True)
        elif isinstance(bins, (int, np.integer)):
            for i in range(nrow):
                if self._range is not None:
                    a, b = map(float, self._range)
                    if a > b:
                        raise ValueError('max must be larger than min in range parameter.')
                else:
                    a, b = float(lo[i]), float(hi[i])
                if not (np.isfinite(a) and np.isfinite(b)):
                    raise ValueError('autodetected range of [{}, {}] is not finite'.format(a, b))
                if a == b:                       # numpy's histogram convention
                    a, b = a - 0.5, b + 0.5
                self._edges.append(np.linspace(a, b, int(bins) + 1))
                self._uniform.append(True)
        else:
            arr = np.asarray(bins)
            if arr.dtype == object and arr.shape == vshape:
                rows = [np.asarray(arr[i], dtype=float) for i in np.ndindex(vshape)]
            elif arr.shape[:-1] == vshape:
                rows = [np.asarray(arr[i], dtype=float) for i in np.ndindex(vshape)]
            else:
                raise ValueError(
                    'shape of the bins {} not compatible with the shape {} of '
                    'sample data points'.format(arr.shape, vshape))
            for e in rows:
                if (np.diff(e) < 0).any():
                    raise ValueError('`bins` must increase monotonically, when an array')
                self._edges.append(e)
                self._uniform.append(False)
        self._groups = self._alloc_groups([len(e) - 1 for e in self._edges], self._uniform,
                                          dev, edges=self._edges, centre=centre)
        self._bins = np.empty(vshape, dtype=object)
        for j, i in enumerate(np.ndindex(vshape)):
            self._bins[i] = self._edges[j]

    def _read_histogram(self, m, host):
        """Cumulated counts of this update's device-to-host copy."""
        vshape = self.vshape
        self._counts = np.empty(vshape, dtype=object)
        self._paths_outside = np.zeros(vshape, dtype=self.ctype)
        index = list(np.ndindex(vshape))
        for g, h in zip(self._groups, host):
            c, o = h['counts'], h['outside']
            for k in range(g['r1'] - g['r0']):
                i = index[g['r0'] + k]
                self._counts[i] = c[k].astype(self.ctype)
                self._paths_outside[i] = int(o[k])
                if self._counts[i].sum() + self._paths_outside[i] != self.paths + m:
                    raise RuntimeError(
                        'total number of cumulated paths inconsistent with stored '
                        'cumulated counts')

    def allreduce(self, group=None):
        """Merge the statistics cumulated by every rank (paths sharded across
        GPUs): ONE all-reduce (SUM) of the packed vector -- path count, mean,
        the four moments re-centred on a common constant (rank 0's centre,
        shifted binomially), histogram bins and the out-of-range counts.  All
        ranks must use the same bin edges (pass explicit ``bins`` edges or an
        integer ``bins`` with a ``range``).  NCCL when the process group is
        NCCL, gloo on CPU.  Returns self (updated in place on every rank)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or self.paths == 0:
            return self
        backend = dist.get_backend(group)
        dev = (torch.device('cuda', torch.cuda.current_device())
               if backend == 'nccl' else torch.device('cpu'))
        rank = dist.get_rank(group)
        vshape = self.vshape
        # common centre and (if any) common edges: taken from rank 0
        head = [np.asarray(self._center, dtype=float).ravel()]
        has_hist = self._counts is not None
        if has_hist:
            head += [np.asarray(e, dtype=float) for e in self._edges]
        ref = torch.from_numpy(np.concatenate(head)).to(dev)
        mine = ref.clone()
        dist.broadcast(ref, src=dist.get_global_rank(group, 0) if group is not None else 0,
                       group=group)
        ref_np = ref.cpu().numpy()
        nc = int(np.prod(vshape, dtype=int))
        c0 = ref_np[:nc].reshape(vshape)
        if has_hist and not np.array_equal(ref_np[nc:], mine.cpu().numpy()[nc:]):
            raise ValueError('montecarlo.allreduce: ranks use different histogram '
                             'bins; pass explicit edges or bins=int with range=')
        # moments about c0:  E[(x-c0)^k] = sum_j C(k,j) (c-c0)^(k-j) E[(x-c)^j]
        from math import comb
        d = np.asarray(self._center, dtype=float) - c0
        m = [np.ones(vshape)] + [np.asarray(mm, dtype=float) for mm in self._moments]
        shifted = [sum(comb(k, j)*d**(k - j)*m[j] for j in range(k + 1)) for k in range(1, 5)]
        n = float(self.paths)
        parts = [np.array([n])] + [n*np.asarray(self._mean, dtype=float).ravel()] + [
            n*sk.ravel() for sk in shifted]
        if has_hist:
            parts += [np.asarray(c, dtype=float) for c in
                      (self._counts[i] for i in np.ndindex(vshape))]
            parts += [np.asarray(self._paths_outside, dtype=float).ravel()]
        buf = torch.from_numpy(np.concatenate(parts)).to(dev)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        buf = buf.cpu().numpy()
        N = buf[0]
        at = 1
        self._mean[...] = (buf[at:at + nc]/N).reshape(vshape); at += nc
        for k in range(4):
            self._moments[k][...] = (buf[at:at + nc]/N).reshape(vshape); at += nc
        self._center = c0.astype(self._center.dtype)
        cflat = np.ascontiguousarray(self._center, dtype=float).ravel()
        for g in getattr(self, '_groups', ()):
            # later updates cumulate about the common centre
            g['centre'].copy_(torch.from_numpy(cflat[g['r0']:g['r1']]))
        if has_hist:
            for j, i in enumerate(np.ndindex(vshape)):
                nb = len(self._edges[j]) - 1
                tot = np.rint(buf[at:at + nb]).astype(self.ctype); at += nb
                self._counts[i] = tot
                self._counts_dev[j].copy_(torch.from_numpy(tot.astype(np.int64)))
            out = np.rint(buf[at:at + nc]).astype(self.ctype).reshape(vshape)
            self._paths_outside = out
            for j in range(nc):
                self._outside_dev[j].fill_(int(out.ravel()[j]))
        self._paths[0] = int(round(N))
        return self

    def __getitem__(self, i):
        a = montecarlo(bins=self._bins if self.paths == 0 else None)
        a._paths = self._paths
        a.dtype, a.ctype = self.dtype, self.ctype
        if self.paths != 0:
            a._mean = self._mean[i]
            a._moments = tuple(mm[i] for mm in self._moments)
            if self._counts is not None:
                a._bins = self._bins[i]
                a._counts = self._counts[i]
                a._paths_outside = self._paths_outside[i]
        return a

    # statistics (reference infrastructure.py:3043-3076)
    def mean(self):
        return self._mean

    def var(self):
        m1, m2 = self._moments[:2]
        return 0.*m1 if self.paths < 2 else m2 - m1*m1

    def std(self):
        return np.sqrt(self.var())

    def skew(self):
        m1, m2, m3 = self._moments[:3]
        return (0.*m1 if self.paths < 2 else
                (m3 - 3*m1*m2 + 2*m1**3)/(m2 - m1*m1)**1.5)

    def kurtosis(self):
        # as in the reference (3062-3067) the value for paths >= 2 is the raw,
        # not the excess, kurtosis
        m1, m2, m3, m4 = self._moments[:4]
        return (-3.0 + 0.*m1 if self.paths < 2 else
                (m4 - 4*m1*m3 + 6*m1*m1*m2 - 3*m1**4)/(m2 - m1*m1)**2)

    def stderr(self):
        return np.nan if self.paths < 2 else np.sqrt(self.var()/(self.paths - 1))

    def __repr__(self):
        if self.paths == 0:
            return '<empty montecarlo object>'
        mean, err = np.asarray(self.mean()), np.asarray(self.stderr())
        if mean.size == 1 and err.size == 1:
            mean, err = mean.flatten()[0], err.flatten()[0]
        return repr(mean) + ' +/- ' + repr(err)

    def histogram(self):
        """(counts, bins) of the cumulated samples (reference 3090-3115)."""
        if self.paths == 0 or self._bins is None or self._counts is None:
            raise ValueError('no distribution data available')
        counts, bins = self._counts, self._bins
        if isinstance(counts, np.ndarray) and counts.dtype == object:
            if counts.size > 1:
                raise IndexError(
                    'histograms and distributions must be invoked on '
                    'single-valued montecarlo instances; use indexing to '
                    'select the value to be addressed (es. ``a[i].histogram()``)')
            counts, bins = counts.flatten()[0], bins.flatten()[0]
        return counts, bins

    def density_histogram(self):
        counts, bins = self.histogram()
        return counts/counts.sum()/np.diff(bins), bins

    @property
    def outpaths(self):
        self.histogram()
        return self._paths_outside

    def outerr(self):
        return self.outpaths/self.paths

    def _density(self, x, method, bandwidth, kind, cumulative):
        """Estimate of the sample pdf / cdf from the cumulated density histogram
        (reference infrastructure.py:3149-3278): Gaussian kernels centred on the
        bin midpoints (bandwidth x bin width), or interpolation of the
        histogram."""
        import scipy.interpolate
        import scipy.special
        dens, edges = self.density_histogram()
        x = np.asarray(x)
        width = np.diff(edges)
        mid = (edges[:-1] + edges[1:])/2
        if method == 'gaussian_kde':
            z = (x[..., np.newaxis] - mid)/(width*bandwidth)
            if cumulative:
                k = width*(scipy.special.erf(z/np.sqrt(2)) + 1)/2
            else:
                k = np.exp(-z*z/2)/np.sqrt(2*np.pi)/bandwidth
            return (k*dens).sum(axis=-1)
        if method == 'interp':
            if cumulative:
                xs = edges
                ys = np.concatenate(((0.,), (width*dens).cumsum()))
                fill = (0., 1.)
            else:
                xs = np.concatenate((edges[:1], mid, edges[-1:]))
                ys = np.concatenate(((0.,), dens, (0.,)))
                fill = 0.
            return scipy.interpolate.interp1d(
                xs, ys, kind=kind, assume_sorted=True, bounds_error=False,
                copy=False, fill_value=fill)(x)
        raise ValueError("pdf or cdf method should be 'gaussian_kde' or "
                         "'interp', not {}".format(method))

    def pdf(self, x, method='gaussian_kde', bandwidth=1., kind='linear'):
        return self._density(x, method, bandwidth, kind, False)

    def cdf(self, x, method='gaussian_kde', bandwidth=1., kind='linear'):
        return self._density(x, method, bandwidth, kind, True)

    m = property(lambda self: self.mean())
    s = property(lambda self: self.std())
    e = property(lambda self: self.stderr())
    h = property(lambda self: self.histogram())
    dh = property(lambda self: self.density_histogram())

    @property
    def stats(self):
        return dict(mean=self.mean(), stderr=self.stderr(), std=self.std(),
                    skew=self.skew(), kurtosis=self.kurtosis())
