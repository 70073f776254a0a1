#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
PYTHONPATH=. python tools/bench_modes.py > gpurun_out/modes.json 2> gpurun_out/modes.err
tail -3 gpurun_out/modes.err | cut -c1-200
