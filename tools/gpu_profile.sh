#!/bin/bash
# One-B200 evidence run of a round: tests, bench line, launch list, ncu captures
# (lean Heston kernel: raw + source pages; stream / montecarlo kernels: raw pages), mode tables.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -3 gpurun_out/bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bench_1e7.csv \
    python bench.py --paths 1e7 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:integrate_lean_kernel -s 1 -c 1 -f -o gpurun_out/prof_lean \
    python bench.py --paths 1e7 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_lean.ncu-rep --page raw --csv > gpurun_out/ncu_lean_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_lean.ncu-rep --page source --csv > gpurun_out/ncu_lean_source.csv 2>/dev/null
for m in replay_ou c2a c2b mc; do
  ncu --set full --clock-control none --import-source on -k regex:'mc_update|moments|stream' -s $([ $m = mc ] && echo 2 || echo 1) -c $([ $m = mc ] && echo 2 || echo 1) -f -o gpurun_out/prof_$m \
      python tools/run_mode.py $m > gpurun_out/ncu_$m.log 2>&1
  ncu -i gpurun_out/prof_$m.ncu-rep --page raw --csv > gpurun_out/ncu_${m}_raw.csv 2>/dev/null
done
PYTHONPATH=. python tools/bench_modes.py > gpurun_out/modes.json 2> gpurun_out/modes.err
PYTHONPATH=. python tools/bench_reductions.py > gpurun_out/reductions.json 2> gpurun_out/reductions.err
