#!/bin/bash
# Sharded prover: per-round trace of the resident kernels on N GPUs.  -> gpurun_out/
set -u
mkdir -p gpurun_out
N=${1:-2}
SCB_PERSIST_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/trace_n$N.err | tee gpurun_out/trace_n$N.json
grep "persist m" gpurun_out/trace_n$N.err | tail -4
