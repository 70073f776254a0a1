"""world_size-2 gloo tests (CPU): path sharding and the single all-reduce of
the packed statistics vector."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _montecarlo_state(sample, edges):
    """A montecarlo object in the state its device update() would leave it in
    (built with NumPy here: the CPU box has no GPU to run update())."""
    import sdepy_b200 as sd
    a = sd.montecarlo(bins=25)
    vshape, m = sample.shape[:-1], sample.shape[-1]
    a._center = sample.mean(axis=-1)
    d = sample - a._center[..., None]
    a._moments = tuple((d**k).mean(axis=-1) for k in (1, 2, 3, 4))
    a._mean = sample.mean(axis=-1)
    a._edges, a._uniform = [edges for _ in np.ndindex(vshape)], [True]*int(np.prod(vshape))
    a._counts = np.empty(vshape, dtype=object)
    a._bins = np.empty(vshape, dtype=object)
    a._paths_outside = np.zeros(vshape, dtype=np.int64)
    a._counts_dev, a._outside_dev = [], []
    for i in np.ndindex(vshape):
        c, _ = np.histogram(sample[i], bins=edges)
        a._counts[i], a._bins[i] = c.astype(np.int64), edges
        a._paths_outside[i] = m - c.sum()
        a._counts_dev.append(torch.from_numpy(c.astype(np.int64)))
        a._outside_dev.append(torch.tensor([m - c.sum()]))
    a._paths[0] = m
    return a


def _montecarlo_merge_ok(rank, world):
    rng = np.random.default_rng(5)
    x = rng.lognormal(size=(2, 3000))
    edges = np.linspace(0., 6., 26)
    lo, hi = (0, 1300) if rank == 0 else (1300, 3000)
    a = _montecarlo_state(x[:, lo:hi], edges).allreduce()
    full = _montecarlo_state(x, edges)
    ok = a.paths == 3000
    for f in ('mean', 'var', 'std', 'skew', 'kurtosis', 'stderr'):
        ok = ok and np.allclose(getattr(a, f)(), getattr(full, f)(), rtol=1e-10)
    for i in range(2):
        ok = ok and np.array_equal(a[i].histogram()[0], full[i].histogram()[0])
        ok = ok and a[i].outpaths == full[i].outpaths
    return bool(ok)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from sdepy_b200 import _lib
        from sdepy_b200.distributed import shard, allreduce_histogram, allreduce_minmax
        from sdepy_b200.integration import path_stats
        total = 1001
        off, cnt = shard(total)
        rng = np.random.default_rng(0)
        x = rng.lognormal(size=(3, 2, total))          # [rows, comps, paths]
        c = np.array([1., 1.5])
        mine = x[..., off:off + cnt]
        sums = np.zeros((3, 2, _lib.NSTAT))
        d = mine - c[None, :, None]
        for k in range(4):
            sums[..., k] = (d**(k + 1)).sum(axis=-1)
        sums[..., 4], sums[..., 5] = mine.min(axis=-1), mine.max(axis=-1)
        pay = np.maximum(mine - 1., 0.)
        sums[..., 6], sums[..., 7] = pay.sum(axis=-1), (pay**2).sum(axis=-1)
        st = path_stats(np.arange(3.), sums, c, cnt).allreduce()
        ok = (st.paths == total and
              np.allclose(np.asarray(st.pmean())[..., 0], x.mean(axis=-1), rtol=1e-13) and
              np.allclose(np.asarray(st.pvar())[..., 0], x.var(axis=-1), rtol=1e-11) and
              np.array_equal(np.asarray(st.pmin())[..., 0], x.min(axis=-1)) and
              np.array_equal(np.asarray(st.pmax())[..., 0], x.max(axis=-1)) and
              np.allclose(np.asarray(st.payoff_mean())[..., 0],
                          np.maximum(x - 1., 0.).mean(axis=-1), rtol=1e-13))
        lo, hi = allreduce_minmax(mine[0, 0].min(), mine[0, 0].max())
        ok = ok and lo == x[0, 0].min() and hi == x[0, 0].max()
        edges = np.linspace(lo, hi, 11)
        cts, _ = np.histogram(mine[0, 0], bins=edges)
        tot, out = allreduce_histogram(cts, cnt - cts.sum())
        ref, _ = np.histogram(x[0, 0], bins=edges)
        ok = ok and np.array_equal(tot, ref) and out == 0
        ok = ok and _montecarlo_merge_ok(rank, world)
        q.put((rank, off, cnt, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_shard_covers_every_path_once():
    from sdepy_b200.distributed import shard
    for total in (10, 1001, 10**8 + 3):
        for world in (1, 2, 3, 8):
            parts = [shard(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (o0, c0), (o1, _) in zip(parts, parts[1:]):
                assert o0 + c0 == o1


def test_allreduce_of_packed_statistics_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert res[0][1] == 0 and res[0][2] == 501 and res[1][1] == 501 and res[1][2] == 500
    assert all(r[3] for r in res)
