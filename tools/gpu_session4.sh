#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
PYTHONPATH=. python tools/bench_reductions.py > gpurun_out/reductions.json 2> gpurun_out/reductions.err; tail -3 gpurun_out/reductions.err
python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -5 gpurun_out/bench_1gpu.err
for m in mc c2b; do
  ncu --set full --clock-control none --import-source on -k regex:'integrate|mc_update|moments|stream' -s $([ $m = mc ] && echo 2 || echo 1) -c $([ $m = mc ] && echo 2 || echo 1) -f -o gpurun_out/prof_$m \
      python tools/run_mode.py $m > gpurun_out/ncu_$m.log 2>&1
  ncu -i gpurun_out/prof_$m.ncu-rep --page raw --csv > gpurun_out/ncu_${m}_raw.csv 2>/dev/null
done
