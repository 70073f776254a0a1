#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_base.json').read().strip().splitlines()[-1])
print('base', '%.4e'%d['value'], '%.3f ms'%d['ms_per_step'], d['check']['call_price_last_step'])"
bash tools/run_variants.sh gpurun_out/variants nofix skew350 both ppt3 ppt4
