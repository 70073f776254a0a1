#!/bin/bash
# full GPU test-suite + mode benchmarks (one B200)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
PYTHONPATH=. python tools/bench_modes.py > gpurun_out/modes.json 2> gpurun_out/modes.err
grep "C4 merton replay" gpurun_out/modes.err | cut -c1-400; tail -2 gpurun_out/modes.err | cut -c1-200
