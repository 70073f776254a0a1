#!/bin/bash
# compute-sanitizer passes over the integration kernels (one B200): memcheck on the
# smoke run and the stream / parity / edge tests, racecheck on the stream / parity tests
# (shared-memory staging, cp.async rings, the aligned generator tables and the gap in
# front of them).  Every command is bounded by its own timeout.
mkdir -p gpurun_out
S="compute-sanitizer --error-exitcode 9"
run() { name=$1; shift; timeout "$T" "$@" > gpurun_out/san_$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/san_summary.log; grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/san_$name.log | tail -3 | tee -a gpurun_out/san_summary.log; }
: > gpurun_out/san_summary.log
T=200 run mem_smoke $S --tool memcheck python -c "import __graft_entry__ as g; g.smoke()"
T=420 run mem_stream_parity $S --tool memcheck python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py -q -x
T=300 run mem_edge_jit $S --tool memcheck python -m pytest tests/test_gpu_edge.py tests/test_gpu_jit.py -q -x
T=420 run race_stream_parity $S --tool racecheck python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py -q -x
T=420 run mem_stats_dropin $S --tool memcheck python -m pytest tests/test_gpu_stats.py tests/test_gpu_dropin.py -q -x
T=420 run race_stats_edge $S --tool racecheck python -m pytest tests/test_gpu_stats.py tests/test_gpu_edge.py -q -x
T=300 run sync_stream_parity $S --tool synccheck python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py -q -x
