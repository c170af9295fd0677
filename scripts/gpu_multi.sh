#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/gpus.txt
echo "== sharded parity"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -12 | tee gpurun_out/mgpu_check.log
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -3 gpurun_out/bench_n$N.err
echo "== bench N=1 (same box)"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n1_samebox.json
