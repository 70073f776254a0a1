#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -5 gpurun_out/bench_1gpu.err
PYTHONPATH=. python tools/bench_reductions.py > gpurun_out/reductions.json 2> gpurun_out/reductions.err; tail -3 gpurun_out/reductions.err
PYTHONPATH=. python tools/bench_modes.py > gpurun_out/modes.json 2> gpurun_out/modes.err
for m in replay_ou c2a c2b mc; do
  ncu --set full --clock-control none --import-source on -k regex:'integrate|mc_update|moments|stream' -s $([ $m = mc ] && echo 2 || echo 1) -c $([ $m = mc ] && echo 2 || echo 1) -f -o gpurun_out/prof_$m \
      python tools/run_mode.py $m > gpurun_out/ncu_$m.log 2>&1
  ncu -i gpurun_out/prof_$m.ncu-rep --page raw --csv > gpurun_out/ncu_${m}_raw.csv 2>/dev/null
done
