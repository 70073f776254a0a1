"""Across-path reductions on a resident row: moments / histogram / cdf / chf kernels (GB/s)."""
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import json  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sdepy_b200 as sd  # noqa: E402
from sdepy_b200 import _cuda  # noqa: E402

HBM_GBS = 6553.6


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)*1e-3)
    return min(ts)


def main():
    n = 100_000_000
    x = torch.empty((1, n), dtype=torch.float64, device='cuda').normal_()
    res = []

    def rec(name, t, passes=1):
        res.append(dict(op=name, seconds=t, GBps=8.*n*passes/t/1e9,
                        frac_of_copy_peak=8.*n*passes/t/1e9/HBM_GBS))
    rec('moments_kernel (S1..S4, min, max) on 1e8 values', timed(lambda: _cuda.moments(x, n)))
    edges = np.linspace(-4., 4., 101)
    counts = torch.zeros(100, dtype=torch.int64, device='cuda')
    outside = torch.zeros(1, dtype=torch.int64, device='cuda')
    rec('histogram_kernel, 100 uniform bins',
        timed(lambda: _cuda.histogram(x[0], edges, counts, outside, True)))
    dp = sd.device_process(np.zeros(1), x.reshape(1, n))
    rec('device_process.pmean (two passes)', timed(dp.pmean), 2)
    rec('device_process.pvar (two passes)', timed(dp.pvar), 2)
    q = np.linspace(-3., 3., 16)
    rec('device_process.cdf, 16 thresholds', timed(lambda: dp.cdf(q)))
    rec('device_process.chf, 16 frequencies', timed(lambda: dp.chf(q)))
    rec('montecarlo(x, bins=100), whole call: range pass + fused moments/histogram pass + D2H',
        timed(lambda: sd.montecarlo(x[0], bins=100)), 2)
    mc = sd.montecarlo(x[0], bins=100)
    rec('montecarlo.update(x) on a cumulating object: one fused pass + D2H',
        timed(lambda: mc.update(x[0])), 1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
