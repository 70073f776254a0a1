#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -3 gpurun_out/bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bench_1e7.csv \
    python bench.py --paths 1e7 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:integrate_lean_kernel -s 1 -c 1 -f -o gpurun_out/prof_lean \
    python bench.py --paths 1e7 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_lean.ncu-rep --page raw --csv > gpurun_out/ncu_lean_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_lean.ncu-rep --page source --csv > gpurun_out/ncu_lean_source.csv 2>/dev/null
PYTHONPATH=. python tools/bench_modes.py > gpurun_out/modes.json 2> gpurun_out/modes.err
