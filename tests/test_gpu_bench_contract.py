"""bench.py end to end on one GPU (reduced path count): the JSON line carries
every key of the measurement contract and its built-in checks pass."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--paths', '2e6',
                          '--steps', '2', '--warmup', '3', '--cpu-sample-paths', '20000'],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
              'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'e2e',
              'gpu_launches', 'clocks', 'roofline', 'cpu_baseline', 'check', 'modes'):
        assert k in d, k
    assert d['metric'] == d['unit'] == 'path-steps/s' and d['dtype'] == 'f64'
    assert d['n_gpus'] == 1 and d['steps'] == 2 and d['warmup'] == 3 and d['gpu_launches'] == 4
    assert d['value'] > 1e10 and 0 < d['e2e']['value'] <= d['value']*1.05
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    r = d['roofline']
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic', 'executed', 'issue'):
        assert k in r, k
    assert 0 < r['issue']['frac'] < 1 and 0 < r['executed']['frac'] < 1
    assert r['counters']['commit'] and abs(r['frac'] - r['achieved']/r['peak']) < 1e-12
    c = d['check']
    assert c['c5_allreduce_ok'] is True
    assert abs(c['call_price_last_step'] - c['closed_form']) < 5*c['stderr'] + 1e-2
    m = d['modes']
    assert m['replay_ou']['frac_of_hbm_peak'] > .5 and m['C2a_ou_tdep_philox']['GBps'] > 500
    assert 0 < m['heston_draws_full']['full_over_fast'] < 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] == 1


def test_bench_reference_arm_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith('{')][-1])
    assert d['impl'] == 'reference' and d['gpu_launches'] == 0
    assert d['e2e']['value'] == d['value'] == d['cpu_baseline']['value'] > 1e6
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['cpu_baseline']['kind'] == 'port'
