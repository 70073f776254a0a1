#!/bin/bash
# usage: tools/run_variants.sh <outdir> <variant>...   (on the GPU box: times bench.py with each variant library)
out=$1; shift
mkdir -p $out
for v in "$@"; do
  SDEB_LIB=gpurun_variants/libsdeb_$v.so python bench.py --no-cpu-baseline --steps 5 --warmup 3 > $out/bench_$v.json 2> $out/bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open('$out/bench_$v.json').read().strip().splitlines()[-1])
    print('$v', '%.4e'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e %.4e'%d['e2e']['value'], d['clocks']['sm_mhz'], d['check']['call_price_last_step'])
except Exception as e:
    print('$v', 'FAILED', e); print(open('$out/bench_$v.err').read()[-1500:])
PY
done
