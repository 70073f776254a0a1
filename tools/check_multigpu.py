"""Multi-GPU check (torchrun, NCCL): BASELINE config 5 sharded over the ranks
-- custom @integrate SDE, Milstein, montecarlo moments + histogram all-reduced --
must equal the same global paths integrated by one process.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_multigpu.py
"""
import os

import numpy as np
import torch
import torch.distributed as dist

import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import sdepy_b200 as sd  # noqa: E402
from sdepy_b200.distributed import shard


def main():
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    total = 20_000_001

    @sd.integrate
    def gbm(t, x, mu=.05, sigma=.2):
        return {'dt': mu*x, 'dw': sigma*x}

    kw = dict(steps=201, x0=1., method='milstein', seed=12, output='device', getinfo=False)
    off, cnt = shard(total)
    x = gbm(paths=cnt, path_offset=off, **kw)((0., 1.))
    edges = np.linspace(0., 3., 101)
    mc = sd.montecarlo(x.x[-1], bins=edges).allreduce()
    st = gbm(paths=cnt, path_offset=off, **dict(kw, output='stats'))((0., 1.)).allreduce()
    if rank == 0:
        full = gbm(paths=total, **kw)((0., 1.))
        ref = sd.montecarlo(full.x[-1], bins=edges)
        assert mc.paths == total == st.paths
        assert np.array_equal(mc.histogram()[0], ref.histogram()[0])
        assert mc.outpaths == ref.outpaths
        for f in ('mean', 'var', 'skew', 'kurtosis', 'stderr'):
            assert np.allclose(getattr(mc, f)(), getattr(ref, f)(), rtol=1e-10), f
        assert np.allclose(np.asarray(st.pmean())[-1, 0], ref.mean(), rtol=1e-12)
        assert np.allclose(np.asarray(st.pvar())[-1, 0], ref.var(), rtol=1e-9)
        print('multi-GPU check ok: world=%d, %d paths, mean %.6f +/- %.6f'
              % (dist.get_world_size(), total, float(mc.mean()), float(mc.stderr())))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
