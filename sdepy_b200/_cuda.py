"""Device plumbing: torch owns device memory and streams, libsdeb does the work.

PyTorch is used here only for allocation, H2D/D2H copies, the current-stream
handle and (in ``distributed``) NCCL; no torch operator is on the compute path.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            'sdepy_b200 needs a CUDA device (B200, sm_100a): there is no CPU '
            'fallback for the integration / statistics kernels')


def device(dev=None):
    require_cuda()
    if dev is None:
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device(dev)


def stream_ptr(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def to_device(a, dev, dtype=None):
    """Host ndarray -> device tensor through pinned memory (async H2D)."""
    a = np.ascontiguousarray(a, dtype=dtype)
    t = torch.from_numpy(a)
    if a.size:
        t = t.pin_memory()
    return t.to(dev, non_blocking=True)


_TORCH_DTYPE = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}


def upload_packed(arrays, dev):
    """Several small host arrays -> device tensors through ONE page-locked
    staging buffer and ONE async H2D copy (the per-launch tables: steps, store
    rows, parameter records, initial state, centre)."""
    arrays = [np.ascontiguousarray(a) for a in arrays]
    offs, at = [], 0
    for a in arrays:
        offs.append(at)
        at += (a.nbytes + 15) & ~15
    host = torch.empty(max(at, 16), dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    for a, o in zip(arrays, offs):
        if a.nbytes:
            hv[o:o + a.nbytes] = a.view(np.uint8).reshape(-1)
    d = host.to(dev, non_blocking=True)
    out = []
    for a, o in zip(arrays, offs):
        t = d[o:o + a.nbytes].view(_TORCH_DTYPE[a.dtype]).reshape(a.shape)   # (dtype.name is slow)
        t._sdeb_staging = host           # keep the pinned block until the copy has run
        out.append(t)
    return out


# Host copies of large results go through page-locked memory: a pageable
# `.cpu()` of a full-path slab runs at ~2 GB/s, a pinned copy at ~57 GB/s
# (torch's caching host allocator keeps the pinned block for the next call).
PINNED_OUTPUT_CAP = int(float(os.environ.get('SDEB_PINNED_OUTPUT_GIB', '8'))*2**30)


def to_host(t):
    """Device tensor -> NumPy array (owning page-locked memory when it fits
    under the cap, which SDEB_PINNED_OUTPUT_GIB sets; 0 disables)."""
    nbytes = t.numel()*t.element_size()
    if nbytes == 0 or nbytes > PINNED_OUTPUT_CAP or not t.is_cuda:
        return t.cpu().numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


class device_array(torch.Tensor):
    """CUDA tensor handed out by device-resident sources: still a tensor for
    the kernels (no copy when it is fed back as ``dw=``), and an array for NumPy
    (``np.asarray``, ``assert_array_equal``, arithmetic with ndarrays pull a
    host copy), so that code written against the reference's ndarray-returning
    sources keeps working."""

    def __array__(self, dtype=None, copy=None):
        a = to_host(self.detach().as_subclass(torch.Tensor))
        return a if dtype is None else a.astype(dtype, copy=False)

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        """NumPy ufuncs (``np.exp(z)``, ``ndarray * z``) work on host copies
        and return ndarrays."""
        args = [np.asarray(a) if isinstance(a, device_array) else a for a in inputs]
        return getattr(ufunc, method)(*args, **kw)


def _mixed(name):
    """Binary operator that stays on the device for scalars and tensors and
    hands ndarray operands to NumPy."""
    base = getattr(torch.Tensor, name)

    def op(self, other):
        if isinstance(other, np.ndarray):
            return getattr(np.asarray(self), name)(other)
        return base(self, other)
    op.__name__ = name
    return op


for _name in ('__add__', '__radd__', '__sub__', '__rsub__', '__mul__', '__rmul__',
              '__truediv__', '__rtruediv__', '__pow__', '__eq__', '__ne__',
              '__lt__', '__le__', '__gt__', '__ge__'):
    setattr(device_array, _name, _mixed(_name))
device_array.__hash__ = torch.Tensor.__hash__


def as_device_array(t):
    return t.as_subclass(device_array)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def empty(shape, dev, dtype=torch.float64):
    return torch.empty(shape, dtype=dtype, device=dev)


def zeros(shape, dev, dtype=torch.float64):
    return torch.zeros(shape, dtype=dtype, device=dev)


def moments(x2d, n_paths, centre=None):
    """Power sums / min / max over the last axis of a device tensor viewed as
    [rows, pitch] (sdeb_moments).  Returns a host array [rows, NSTAT]."""
    dev = x2d.device
    rows, pitch = x2d.shape
    out = np.empty((rows, _lib.NSTAT))
    for r0 in range(0, rows, 32768):
        r1 = min(rows, r0 + 32768)
        n = r1 - r0
        ws, ws_bytes = _workspace(n, dev)
        stats = empty((n, _lib.NSTAT), dev)
        c = None
        if centre is not None:
            c = to_device(np.asarray(centre, dtype=float).reshape(-1)[r0:r1], dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.sdeb_moments(
                ptr(x2d[r0:r1]), n, n_paths, pitch, ptr(c), ptr(stats), ptr(ws),
                ws_bytes, stream_ptr(dev)))
        out[r0:r1] = stats.cpu().numpy()
    return out


def histogram(x1d, edges, counts, outside, uniform):
    """counts/outside (device int64 tensors) += histogram of x1d on edges."""
    dev = x1d.device
    e = to_device(np.asarray(edges, dtype=float), dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.sdeb_histogram(
            ptr(x1d), x1d.numel(), ptr(e), len(edges) - 1, int(bool(uniform)),
            ptr(counts), ptr(outside), stream_ptr(dev)))


def _workspace(rows, dev):
    nbytes = _lib.lib.sdeb_moments_workspace(rows)
    return empty((nbytes // 8,), dev), nbytes


def mc_range(x2d, n, out=None):
    """Pass 1 of montecarlo's first update (sdeb_mc_range): device tensor
    [rows, NSTAT] holding sum (slot 0), min (4), max (5) of every row.  No
    host synchronisation."""
    dev = x2d.device
    rows, pitch = x2d.shape
    stats = empty((rows, _lib.NSTAT), dev) if out is None else out
    ws, ws_bytes = _workspace(rows, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.sdeb_mc_range(ptr(x2d), rows, n, pitch, ptr(stats), ptr(ws),
                                          ws_bytes, stream_ptr(dev)))
    return stats


def mc_update(x2d, n, *, centre=None, centre_out=None, range_stats=None, lo=0., hi=0.,
              edges_mode=_lib.MC_EDGES_GIVEN, edges=None, nbins=0, uniform=True,
              counts=None, outside=None, out=None):
    """Fused moments + histogram pass (sdeb_mc_update) over the rows of x2d;
    all arrays are device tensors, the result is the device tensor
    [rows, NSTAT] of centred power sums.  No host synchronisation."""
    dev = x2d.device
    rows, pitch = x2d.shape
    stats = empty((rows, _lib.NSTAT), dev) if out is None else out
    ws, ws_bytes = _workspace(rows, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.sdeb_mc_update(
            ptr(x2d), rows, n, pitch, ptr(centre), ptr(centre_out), ptr(range_stats),
            float(lo), float(hi),
            int(edges_mode), ptr(edges), int(nbins), int(bool(uniform)), ptr(stats),
            ptr(counts), ptr(outside), ptr(ws), ws_bytes, stream_ptr(dev)))
    return stats
