#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -3 gpurun_out/bench_1gpu.err
PYTHONPATH=. python tools/profile_e2e.py 1e6 > gpurun_out/profile_e2e.txt 2>&1
head -12 gpurun_out/profile_e2e.txt
PYTHONPATH=. python tools/bench_reductions.py > gpurun_out/reductions.json 2> gpurun_out/reductions.err
