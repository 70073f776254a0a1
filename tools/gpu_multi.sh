#!/bin/bash
# usage: tools/gpu_multi.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -c 2500 gpurun_out/bench_${N}gpu.json; tail -5 gpurun_out/bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    tools/check_multigpu.py > gpurun_out/check_multigpu_${N}.log 2>&1
tail -3 gpurun_out/check_multigpu_${N}.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${N}gpu_reference.json 2> gpurun_out/bench_${N}gpu_reference.err
tail -c 600 gpurun_out/bench_${N}gpu_reference.json
