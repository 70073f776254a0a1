#!/usr/bin/env python
"""
bench.py -- path-steps/s of the SDE hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[2], the configuration the metric is quoted
on): Heston, correlated 2-factor, European call, fp64 Euler full truncation,
1e8 paths x 252 steps PER GPU (weak scaling: paths are independent units, each
rank integrates its own contiguous range of the global path index), terminal
statistics mode (no path storage; Philox4x32-10 draws generated in-kernel).

A "step" is one pass of the hot path over one batch = one full integration of
`paths` paths over 252 steps + the fold of the per-block partials (+ for N>1
the single all-reduce of the packed statistics vector).

  value : device-timed (CUDA events on the launching stream, max over ranks),
          tables already resident in HBM when the timed region starts.
  e2e   : the same metric through the public API heston_process(...)(timeline)
          with output='stats': host-side lowering, pinned H2D copy of the step /
          parameter tables, kernel, D2H read of the statistics -- every step.
  roofline : the kernel has no HBM traffic to speak of and no tensor-core
          work; its bound is the FP64 pipe (bound="fp64").  achieved =
          ALGORITHMIC work per SURVEY.md 8(d) (100 FP64-pipe instructions per
          Heston path-step at libdevice transcendental costs) x path-steps/s,
          peak = DFMA rate measured live by sdeb_fp64_peak(); both in TFLOP/s
          (2 flop per FP64 lane-instruction).  roofline.executed reports the
          pipe utilisation of what the kernel really issues (53 FP64
          instructions per path-step; = ncu sm__pipe_fp64_cycles_active).
  cpu_baseline : the NumPy oracle port of the reference's Heston path, one
          core, on a bounded sample, same box, same run.

`--impl reference` times the reference's own CPU implementation of the path
(the oracle port: the reference is pure Python + NumPy and cannot travel to
the GPU box) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HESTON = dict(x0=100., mu=.03, sigma=1., y0=.04, theta=.04, k=2., xi=.3)
RHO = -.7
N_STEPS = 252
STRIKE, RATE = 100., .03
# ALGORITHMIC work per Heston path-step, SURVEY.md section 8(d): 34 fp64 flop
# counting each sqrt/log/sin/cos as one = ~100 FP64-pipe instructions with
# libdevice transcendental costs.  This is the per-unit figure of the roofline.
ALGO_FP64_INSTR_PER_PATH_STEP = 100
# FP64-pipe warp-instructions this kernel actually EXECUTES per path-step
# (hand-rolled log / sqrt / sincos): DFMA+DMUL+DADD+DSETP on the Philox /
# no-store path of integrate_lean_kernel<HestonSDE<1,false>>, SASS count,
# confirmed by ncu sm__inst_executed_pipe_fp64.sum / path-steps = 53.5
N64_PER_PATH_STEP = 53
# dram__bytes_read.sum + dram__bytes_write.sum of integrate_lean_kernel<Heston> from the
# ncu --set full capture in profiles/r01_ncu_integrate_lean_heston.csv (1e7 paths x 252
# steps per launch): 80.18 MB read + 22.15 MB written = the per-path int64
# `negative_y_count` diagnostic (getinfo=True, the reference's default: 8 B read per path,
# 8 B written back, part of it still in L2 when the kernel ends); the tables in and the
# per-CTA partial sums out are KBs.  10.2 B per path, once per launch -- not per step.
NCU_DRAM_BYTES_PER_PATH = (80.182528e6 + 22.148864e6)/1e7


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--paths', type=float, default=1e8, help='paths per GPU')
    ap.add_argument('--cpu-sample-paths', type=int, default=2_000_000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


# ---------------------------------------------------------------------------
# CPU side: oracle port of the reference path
# ---------------------------------------------------------------------------

def _cpu_heston(args):
    paths, seed = args
    from oracle import sde_oracle as orc
    par = {k: HESTON[k] for k in ('mu', 'sigma', 'theta', 'k', 'xi')}
    grid = np.linspace(0., 1., N_STEPS + 1)
    t0 = time.perf_counter()
    xT, _ = orc.heston_stream(par, HESTON['x0'], HESTON['y0'], RHO, grid, paths,
                              np.random.default_rng(seed))
    return time.perf_counter() - t0, float(np.maximum(xT - STRIKE, 0).mean())


def cpu_baseline(sample_paths):
    """One core, bounded sample (about 10 s of CPU work)."""
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    dt, _ = _cpu_heston((sample_paths, 1234))
    return dict(value=sample_paths*N_STEPS/dt, unit='path-steps/s', cores=1,
                kind='port',
                sample='oracle.heston_stream (NumPy restatement of sdepy '
                       'heston_process, numpy default_rng draws), %d paths x %d '
                       'steps, %.1f s' % (sample_paths, N_STEPS, dt))


def run_reference(a):
    """Reference arm: the reference's CPU implementation of the path on all
    host cores (oracle port; one worker process per core, independent
    default_rng streams).  Under torchrun only rank 0 works."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per = max(20_000, min(100_000, 4_000_000//cores))
    times = []
    with mp.get_context('spawn').Pool(cores) as pool:
        for it in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_heston, [(per, 1000*it + c) for c in range(cores)])
            if it >= a.warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    value = a.steps*cores*per*N_STEPS/total
    sample = ('%d worker processes x %d paths x %d steps per step (oracle port '
              'of sdepy heston_process)' % (cores, per, N_STEPS))
    print(json.dumps({
        'impl': 'reference', 'metric': 'path-steps/s', 'value': value,
        'unit': 'path-steps/s', 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1e3*total/a.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': config_of(int(a.paths), a.gpus),
        'cpu_baseline': dict(value=value, unit='path-steps/s', cores=cores,
                             kind='port', sample=sample),
        'e2e': dict(value=value, unit='path-steps/s', h2d_bytes_per_step=0,
                    d2h_bytes_per_step=0),
        'gpu_launches': 0}))


def config_of(paths, gpus):
    return {'workload': 'heston_process x0=100 mu=.03 sigma=1 y0=.04 theta=.04 '
                        'k=2 xi=.3 rho=-.7, European call K=100, %d paths/GPU x '
                        '%d Euler steps, fp64, terminal-stats mode, Philox4x32-10 '
                        'in-kernel draws' % (paths, N_STEPS),
            'paths_per_gpu': paths, 'n_steps': N_STEPS,
            'global_paths': paths*gpus, 'parallelism': 'paths sharded x%d' % gpus,
            'l2': 'not applicable: no HBM-resident inputs (state in registers, '
                  'draws generated in-kernel); a 512 MiB buffer is rewritten '
                  'between timed steps anyway'}


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------

class clock_sampler:
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,'
              'clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(names, r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

def run_ours(a):
    import torch
    import torch.distributed as dist
    import sdepy_b200 as sd
    from sdepy_b200 import _engine, _lib, _cuda

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    paths = int(a.paths)
    timeline = (0., 1.)
    grid = np.linspace(0., 1., N_STEPS + 1)
    payoff = ('call', STRIKE, float(np.exp(-RATE)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make(seed):
        return sd.heston_process(paths=paths, steps=grid, rho=RHO, seed=seed,
                                 output='stats', payoff=payoff, getinfo=True,
                                 path_offset=rank*paths, **HESTON)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    # ---- device-timed: tables resident, K launches ------------------------
    res = _engine.resident_stats_run(make(1), timeline)
    packed = torch.zeros(res.stats.numel() + 1, dtype=torch.float64, device=dev)

    def step_resident(i):
        st = res.launch(0x5DEECE66D + i)
        if world > 1:   # the path's only exchange: one all-reduce of the sums
            packed[:-1] = st.reshape(-1)
            dist.all_reduce(packed)
        return st

    for i in range(a.warmup):
        step_resident(i)
    barrier()
    sampler = clock_sampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(a.steps)]
    barrier()
    wall0 = time.perf_counter()
    for i in range(a.steps):
        flush.fill_(i & 0xff)                     # L2 flush, outside the event pair
        ev[i][0].record()
        st = step_resident(a.warmup + i)
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    clocks = sampler.stop() if rank == 0 else None
    sums = st.cpu().numpy()

    # ---- end to end through the public API --------------------------------
    for i in range(2):
        r = make(100 + i)(timeline)
        if world > 1:
            r = r.allreduce()
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        r = make(200 + i)(timeline)               # lowering + H2D + kernel + D2H
        if world > 1:
            r = r.allreduce()
        price = float(np.asarray(r.payoff_mean())[-1, 0])
    barrier()
    e2e_s = time.perf_counter() - t0
    # per step: steps table, store rows, parameter record, initial state, centre
    h2d = N_STEPS*16 + N_STEPS*4 + 9*8 + 2*8 + 8
    d2h = 2*_lib.NSTAT*8

    # ---- reduce timings over ranks ----------------------------------------
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        total_steps = world*paths*N_STEPS*a.steps
        value = total_steps/(dev_ms*1e-3)
        peak = _lib.f64()
        import ctypes
        _lib.check(_lib.lib.sdeb_fp64_peak(200_000, ctypes.byref(peak), _cuda.stream_ptr(dev)))
        peak_tf = peak.value*2/1e12
        per_gpu = value/world
        achieved_tf = per_gpu*ALGO_FP64_INSTR_PER_PATH_STEP*2/1e12
        executed_tf = per_gpu*N64_PER_PATH_STEP*2/1e12
        n = paths
        pay_mean = sums[-1, 0, 6]/n
        pay_se = float(np.sqrt(max(sums[-1, 0, 7]/n - pay_mean**2, 0)/(n - 1)))
        out = {
            'metric': 'path-steps/s', 'value': value, 'unit': 'path-steps/s',
            'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': dev_ms/a.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': config_of(paths, world),
            'e2e': {'value': total_steps/e2e_s, 'unit': 'path-steps/s',
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': a.steps*res.kernels_per_launch,
            'clocks': clocks,
            'roofline': {
                'bound': 'fp64', 'achieved': achieved_tf, 'peak': peak_tf,
                'unit': 'TFLOP/s', 'frac': achieved_tf/peak_tf,
                'traffic': int(NCU_DRAM_BYTES_PER_PATH*paths),
                'executed': {'fp64_instr_per_path_step': N64_PER_PATH_STEP,
                             'achieved': executed_tf, 'frac': executed_tf/peak_tf},
                'note': 'per GPU. achieved = ALGORITHMIC work (SURVEY 8d: %d FP64-pipe '
                        'instr per Heston path-step with libdevice transcendental '
                        'costs) x path-steps/s x 2 flop; peak = DFMA rate measured '
                        'live by sdeb_fp64_peak (MEASURED_PEAKS.json holds no FP64 '
                        'figure). "executed" is the FP64-pipe utilisation of the '
                        'instructions this kernel really issues (%d per path-step: '
                        'its log/sqrt/sincos are hand-rolled), = ncu '
                        'sm__pipe_fp64_cycles_active; frac can exceed 1 because '
                        'the SURVEY figure prices the transcendentals at libdevice '
                        'cost, which this kernel undercuts -- "executed" is the '
                        'utilisation to read'
                        % (ALGO_FP64_INSTR_PER_PATH_STEP, N64_PER_PATH_STEP)},
            'check': {'call_price_last_step': pay_mean, 'stderr': pay_se,
                      'closed_form': 9.2425, 'e2e_price': price},
            'wall_s_timed_region': wall,
        }
        if not a.no_cpu_baseline:
            out['cpu_baseline'] = cpu_baseline(a.cpu_sample_paths)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
