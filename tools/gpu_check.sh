#!/bin/bash
# full GPU test-suite + mode / reduction benchmarks (one B200)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
PYTHONPATH=. python tools/bench_reductions.py > gpurun_out/reductions.json 2> gpurun_out/reductions.err; tail -3 gpurun_out/reductions.err
python -c "
import json
for r in json.load(open('gpurun_out/reductions.json')): print('%-85s %.3f ms %7.0f GB/s %.2f'%(r['op'],r['seconds']*1e3,r['GBps'],r['frac_of_copy_peak']))"
[ "$1" = modes ] && PYTHONPATH=. python tools/bench_modes.py > gpurun_out/modes.json 2> gpurun_out/modes.err
