"""Path sharding across ranks (one process per GPU).

Paths are independent units: rank r of W integrates the contiguous range
[r*P/W, (r+1)*P/W) of the GLOBAL path index -- the Philox counter is the global
index, so results do not depend on W -- with no exchange while stepping.  The
only collective is ONE all-reduce of the packed statistics vector at the end
(``path_stats.allreduce`` / ``allreduce_histogram``): NCCL over NVLink when the
process group is NCCL, gloo on CPU.
"""
import numpy as np
import torch


def shard(total_paths, rank=None, world=None):
    """(path_offset, local_paths) of this rank; remainders go to the first
    ranks so that every global index is owned exactly once."""
    import torch.distributed as dist
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        else:
            rank, world = 0, 1
    base, rem = divmod(int(total_paths), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank*base + min(rank, rem)
    return offset, count


def allreduce_histogram(counts, outside, group=None):
    """Sum histogram bin counts and the out-of-range count over ranks
    (int64: exact and order-independent)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return counts, outside
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else 'cpu'
    buf = torch.from_numpy(np.concatenate((np.asarray(counts, dtype=np.int64).ravel(),
                                           [int(outside)]))).to(dev)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    buf = buf.cpu().numpy()
    return buf[:-1].reshape(np.shape(counts)), int(buf[-1])


def allreduce_minmax(lo, hi, group=None):
    """Global (min, max), e.g. to fix common histogram edges before counting
    (np.histogram's range=None rule, reference infrastructure.py:2999-3004)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return lo, hi
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else 'cpu'
    buf = torch.tensor([-float(lo), float(hi)], dtype=torch.float64, device=dev)
    dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=group)
    return -float(buf[0]), float(buf[1])
