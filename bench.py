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
          pipe utilisation of what the kernel really issues (FP64-pipe
          instructions per path-step from the committed ncu counters,
          profiles/r02_lean_heston_counters.json; = ncu
          sm__pipe_fp64_cycles_active); roofline.issue the issue-slot
          utilisation (all warp instructions per warp-step over the cycles
          per warp-step measured in THIS run).
  strong : the same workload with 1e8 GLOBAL paths split over the N ranks
          (value, e2e, efficiency against this run's single-GPU time for 1e8).
  modes  : the HBM-bound full-path configurations (C2a OU, C2b HW-3f, replayed
          OU) measured in this run: stored GB/s against MEASURED_PEAKS.json.
  check  : Monte Carlo price vs closed form, and BASELINE config 5 (custom
          @integrate SDE, Milstein, montecarlo histogram + moments) sharded over
          the ranks and all-reduced vs a single-process recompute.
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
# What the kernel actually EXECUTES per path-step, and its DRAM traffic, are NOT
# constants of this file: they are read from the committed ncu counters of the
# dominant kernel (tools/ncu_counters.py over the `--set full` capture; the JSON
# names the commit and the command it was taken at).
COUNTERS_FILE = os.path.join(ROOT, 'profiles', 'r02_lean_heston_counters.json')
N_SMSP = 148*4              # B200: SM sub-partitions (issue ports)


def load_counters():
    with open(COUNTERS_FILE) as f:
        return json.load(f)


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

def measured_hbm_peak():
    """(GB/s, source): MEASURED_PEAKS.json (driver-written), else the profiling
    recipe's fallback."""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs'
    except Exception:
        return 6553.6, 'fallback of B200_PROFILING.md (no MEASURED_PEAKS.json)'


def full_path_modes(sd, torch, dev, rank=0, world=1, dist=None):
    """The HBM-bound output mode: full paths stored time-major [steps, vars,
    paths] in HBM (output='device').  Kernel time = CUDA events around
    sdeb_integrate on the launching stream (best of 3 after one warm-up).  At
    N > 1 every rank runs the three kernel cases on its own shard of paths at
    the same time (barrier before each) and the figures are the aggregate over
    the ranks at the MAX of their times; the host-facing rows run on rank 0."""
    from sdepy_b200 import _lib
    peak, src = measured_hbm_peak()
    events = []
    real = _lib.lib.sdeb_integrate

    def timed_integrate(p, stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = real(p, stream)
        e1.record()
        events.append((e0, e1))
        return rc

    def kernel_seconds(fn):
        best = None
        for it in range(4):
            del events[:]
            out = fn()
            torch.cuda.synchronize(dev)
            t = sum(e0.elapsed_time(e1) for e0, e1 in events)*1e-3
            del out
            if it and (best is None or t < best):
                best = t
        return best

    def hw_theta(t):
        return np.array(((.02 + .001*t,), (0.,), (0.,)))

    def hw_corr(t):
        c01, c02, c12 = .3*np.cos(t), -.2 + .05*t, .1
        return np.array(((1, c01, c02), (c01, 1, c12), (c02, c12, 1)))

    out = {}
    _lib.lib.sdeb_integrate = timed_integrate
    try:
        p, n = 1_000_000, 500
        tl = np.linspace(0., 5., n + 1)
        cases = [
            ('C2a_ou_tdep_philox', p, n, 0, lambda: sd.ornstein_uhlenbeck_process(
                x0=.1, theta=lambda s: .2 + .1*s, k=1., sigma=.3, paths=p, seed=2,
                output='device')(tl)),
            ('C2b_hw3f_tdep_corr_philox', p, n, 0, lambda: sd.hull_white_process(
                factors=3, x0=((.01,), (0.,), (0.,)), theta=hw_theta,
                k=((.1,), (.5,), (1.,)), sigma=((.01,), (.008,), (.005,)), corr=hw_corr,
                paths=p, seed=3, output='device')(tl))]
        pr, nr = 4_000_000, 250
        g = torch.Generator(device=dev)
        g.manual_seed(0)
        dW = torch.randn((nr, pr), dtype=torch.float64, device=dev, generator=g)*np.sqrt(1/nr)
        tlr = np.linspace(0., 1., nr + 1)
        cases.append(('replay_ou', pr, nr, pr*nr, lambda: sd.ornstein_uhlenbeck_process(
            x0=.1, theta=.2, k=1., sigma=.3, paths=pr, dw=sd.replay_source(dW),
            output='device')(tlr)))
        for name, paths, steps, read, fn in cases:
            if world > 1:
                dist.barrier()
            t = kernel_seconds(fn)
            torch.cuda.empty_cache()
            if world > 1:
                tt_ = torch.tensor([t], dtype=torch.float64, device=dev)
                dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
                t = float(tt_[0])
            nbytes = 8.*(paths*(steps + 1) + read)
            out[name] = {'paths_per_gpu': paths, 'steps': steps, 'kernel_s': t,
                         'path_steps_per_s': world*paths*steps/t,
                         'algorithmic_bytes_per_gpu': nbytes, 'GBps': world*nbytes/t/1e9,
                         'GBps_per_gpu': nbytes/t/1e9,
                         'frac_of_hbm_peak': nbytes/t/1e9/peak}
        del dW
        if rank != 0:
            return None
        # the drop-in default: the same C2a run returned as a HOST process
        # (output='process'): lowering + kernel + pinned D2H of the 4 GB slab
        run = lambda: sd.ornstein_uhlenbeck_process(
            x0=.1, theta=lambda s: .2 + .1*s, k=1., sigma=.3, paths=p, seed=2)(tl)
        run()                                   # pins the host buffer
        walls = []
        for _ in range(3):
            t0 = time.perf_counter()
            xh = run()
            walls.append(time.perf_counter() - t0)
            del xh
        slab = 8.*p*(n + 1)
        out['C2a_e2e_process'] = {
            'paths': p, 'steps': n, 'wall_s': min(walls),
            'path_steps_per_s': p*n/min(walls), 'slab_bytes': slab,
            'd2h_GBps_whole_call': slab/min(walls)/1e9,
            'note': "output='process' (host ndarray subclass, the reference's return type): "
                    'the call is the device-to-host copy of the slab over PCIe (~55 GB/s '
                    'pinned); the kernel is ~2 % of it, so there is nothing to overlap'}
        # price of the full-resolution draw option on the headline workload
        # (96 instead of 64 random bits per pair of normals; NVRTC-compiled)
        grid = np.linspace(0., 1., N_STEPS + 1)
        res = {}
        for tag in ('fast', 'full'):
            P = sd.heston_process(paths=10_000_000, steps=grid, rho=RHO, seed=1, output='stats',
                                  draws=tag, getinfo=False, **HESTON)
            res[tag] = 10_000_000*N_STEPS/kernel_seconds(lambda: P((0., 1.)))
        out['heston_draws_full'] = {
            'paths': 10_000_000, 'steps': N_STEPS, 'path_steps_per_s_fast': res['fast'],
            'path_steps_per_s_full': res['full'], 'full_over_fast': res['full']/res['fast'],
            'note': "draws='full': 52-bit radius uniform + 32-bit angle, one Philox block per "
                    'pair (default: 32 + 32 bits, two pairs per block)'}
    finally:
        _lib.lib.sdeb_integrate = real
    out['hbm_peak_GBps'] = peak
    out['hbm_peak_source'] = src
    out['n_gpus'] = world
    out['note'] = ('full-path output mode, stored rows [steps+1, paths] fp64 in HBM; '
                   'algorithmic bytes = 8 B per stored value (+ 8 B per replayed increment); '
                   'the Philox rows are FP64-issue bound (draws), replay_ou is the HBM-bound one; '
                   'at N > 1 all ranks run at once: GBps and path_steps_per_s are aggregates '
                   'at the max of the ranks\' times, frac_of_hbm_peak is per GPU')
    return out


def c5_allreduce_check(sd, torch, dist, rank, world, dev):
    """BASELINE config 5, scaled: custom @integrate GBM, Milstein, 2e7 GLOBAL
    paths x 2000 steps sharded over the ranks; histogram edges fixed by a first
    min/max all-reduce (np.histogram's range=None rule, reference
    infrastructure.py:2999-3004), then montecarlo(...).allreduce().  Rank 0
    recomputes everything in one process and compares: counts and edges
    array_equal, moments rtol 1e-10."""
    from sdepy_b200.distributed import shard, allreduce_minmax
    total, nsteps = 20_000_003, 2000

    @sd.integrate
    def gbm(t, x, mu=.05, sigma=.2):
        return {'dt': mu*x, 'dw': sigma*x}

    kw = dict(steps=nsteps + 1, x0=1., method='milstein', seed=12, output='device',
              getinfo=False)
    off, cnt = shard(total, rank, world)
    x = gbm(paths=cnt, path_offset=off, **kw)((0., 1.))
    lo, hi = float(np.asarray(x.pmin())[-1, 0]), float(np.asarray(x.pmax())[-1, 0])
    lo, hi = allreduce_minmax(lo, hi)
    mc = sd.montecarlo(x.x[-1], bins=100, range=(lo, hi))
    if world > 1:
        mc.allreduce()
    else:
        # one rank: cumulate the same paths in two chunks instead
        half = cnt//2
        mc = sd.montecarlo(x.x[-1][:half], bins=100, range=(lo, hi))
        mc.update(x.x[-1][half:])
    res = {'global_paths': total, 'steps': nsteps, 'ranks': world}
    if rank == 0:
        full = x if world == 1 else gbm(paths=total, **kw)((0., 1.))
        ref = sd.montecarlo(full.x[-1], bins=100)
        ok = mc.paths == total == ref.paths
        ok = ok and np.array_equal(mc.histogram()[1], ref.histogram()[1])
        ok = ok and np.array_equal(mc.histogram()[0], ref.histogram()[0])
        ok = ok and int(mc.outpaths) == int(ref.outpaths) == 0
        for f in ('mean', 'var', 'skew', 'kurtosis', 'stderr'):
            ok = ok and bool(np.allclose(getattr(mc, f)(), getattr(ref, f)(), rtol=1e-10, atol=0))
        res.update(ok=bool(ok), mean=float(mc.mean()), stderr=float(mc.stderr()),
                   expected_mean=float(np.exp(.05)),
                   histogram_total=int(mc.histogram()[0].sum()))
    return res


def run_ours(a):
    import torch
    import torch.distributed as dist
    import sdepy_b200 as sd
    from sdepy_b200 import _engine, _lib, _cuda
    from sdepy_b200.distributed import shard

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

    def make(seed, npaths, offset):
        return sd.heston_process(paths=npaths, steps=grid, rho=RHO, seed=seed,
                                 output='stats', payoff=payoff, getinfo=True,
                                 path_offset=offset, **HESTON)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def measure(npaths, offset, sample_clocks):
        """(device ms over K steps, e2e seconds over K steps, clocks, last sums,
        last e2e price, kernels per step) for `npaths` local paths."""
        # ---- device-timed: tables resident, K launches --------------------
        res = _engine.resident_stats_run(make(1, npaths, offset), timeline)
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
        if sample_clocks and rank == 0:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(a.steps)]
        barrier()
        for i in range(a.steps):
            flush.fill_(i & 0xff)                 # L2 flush, outside the event pair
            ev[i][0].record()
            st = step_resident(a.warmup + i)
            ev[i][1].record()
        barrier()
        dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        clocks = sampler.stop() if sample_clocks and rank == 0 else None
        sums = st.cpu().numpy()
        # ---- end to end through the public API ----------------------------
        for i in range(max(2, a.warmup)):
            r = make(100 + i, npaths, offset)(timeline)
            if world > 1:
                r = r.allreduce()
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            r = make(200 + i, npaths, offset)(timeline)   # lowering + H2D + kernel + D2H
            if world > 1:
                r = r.allreduce()
            price = float(np.asarray(r.payoff_mean())[-1, 0])
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), clocks, sums, price, res.kernels_per_launch

    wall0 = time.perf_counter()
    dev_ms, e2e_s, clocks, sums, price, kpl = measure(paths, rank*paths, True)
    wall = time.perf_counter() - wall0

    # ---- strong scaling: the same 1e8 GLOBAL paths split over the ranks ----
    strong = None
    if world > 1:
        off, cnt = shard(paths, rank, world)
        s_ms, s_e2e, _, s_sums, s_price, _ = measure(cnt, off, False)
        strong = {
            'global_paths': paths, 'paths_per_gpu': paths//world,
            'value': paths*N_STEPS*a.steps/(s_ms*1e-3), 'unit': 'path-steps/s',
            'ms_per_step': s_ms/a.steps,
            'e2e': paths*N_STEPS*a.steps/s_e2e, 'e2e_ms_per_step': 1e3*s_e2e/a.steps,
            # this run's own single-GPU time for 1e8 paths is the weak line's step
            'efficiency': dev_ms/(world*s_ms), 'e2e_efficiency': e2e_s/(world*s_e2e),
            'e2e_call_price': s_price,
            'note': 'same workload, %d GLOBAL paths sharded over %d ranks (one all-reduce '
                    'of the sums per step); efficiency = T(1 GPU, %d paths, this run: the '
                    'weak line) / (N x T(N GPUs))' % (paths, world, paths)}

    # ---- the other configurations, under the same driver run ---------------
    modes = full_path_modes(sd, torch, dev, rank, world, dist)
    barrier()
    c5 = c5_allreduce_check(sd, torch, dist, rank, world, dev)
    # per step: steps table, store rows, parameter record, initial state, centre
    h2d = N_STEPS*16 + N_STEPS*4 + 9*8 + 2*8 + 8
    d2h = 2*_lib.NSTAT*8

    if rank == 0:
        total_steps = world*paths*N_STEPS*a.steps
        value = total_steps/(dev_ms*1e-3)
        peak = _lib.f64()
        import ctypes
        _lib.check(_lib.lib.sdeb_fp64_peak(200_000, ctypes.byref(peak), _cuda.stream_ptr(dev)))
        peak_tf = peak.value*2/1e12
        per_gpu = value/world
        cnt_ = load_counters()
        n64 = cnt_['fp64_pipe_instr_per_path_step']
        achieved_tf = per_gpu*ALGO_FP64_INSTR_PER_PATH_STEP*2/1e12
        executed_tf = per_gpu*n64*2/1e12
        # issue-slot view: cycles one sub-partition spends per warp-step in THIS
        # run (SM clock sampled under load) against the warp instructions it issues
        sm_hz = 1e6*((clocks or {}).get('sm_mhz') or cnt_['sm_mhz'])
        cyc = N_SMSP*sm_hz*32/per_gpu
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
            'gpu_launches': a.steps*kpl,
            'clocks': clocks,
            'roofline': {
                'bound': 'fp64', 'achieved': achieved_tf, 'peak': peak_tf,
                'unit': 'TFLOP/s', 'frac': achieved_tf/peak_tf,
                'traffic': int(cnt_['dram_bytes_per_path']*paths),
                'executed': {'fp64_instr_per_path_step': n64,
                             'achieved': executed_tf, 'frac': executed_tf/peak_tf},
                'issue': {'warp_instr_per_warp_step': cnt_['warp_instr_per_path_step'],
                          'cycles_per_warp_step': cyc,
                          'frac': cnt_['warp_instr_per_path_step']/cyc,
                          'fp64_pipe_frac': 2*n64/cyc,
                          'floor_cycles': max(cnt_['warp_instr_per_path_step'], 2*n64)},
                'counters': {k: cnt_[k] for k in ('commit', 'source', 'command',
                                                  'ncu_fp64_pipe_pct', 'ncu_issue_active_pct')},
                'note': 'per GPU. achieved = ALGORITHMIC work (SURVEY 8d: %d FP64-pipe '
                        'instr per Heston path-step with libdevice transcendental '
                        'costs) x path-steps/s x 2 flop; peak = DFMA rate measured '
                        'live by sdeb_fp64_peak (MEASURED_PEAKS.json holds no FP64 '
                        'figure); frac exceeds 1 because the kernel\'s hand-rolled '
                        'log/sqrt/sincos undercut the libdevice pricing. "executed" = '
                        'FP64-pipe utilisation of the instructions really issued (%.1f '
                        'per path-step, ncu counters of the named commit; = ncu '
                        'sm__pipe_fp64_cycles_active). "issue" = warp instructions per '
                        'warp-step / cycles a sub-partition spends per warp-step in this '
                        'run (= ncu smsp__issue_active): the fraction to read, < 1 by '
                        'construction; floor_cycles = max(issue slots, 2 x FP64 instr)'
                        % (ALGO_FP64_INSTR_PER_PATH_STEP, n64)},
            'check': {'call_price_last_step': pay_mean, 'stderr': pay_se,
                      'closed_form': 9.2425, 'e2e_price': price,
                      'c5_allreduce_ok': c5.get('ok'), 'c5': c5},
            'modes': modes,
            'wall_s_timed_region': wall,
        }
        if strong is not None:
            out['strong'] = strong
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
