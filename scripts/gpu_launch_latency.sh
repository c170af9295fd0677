#!/bin/bash
set -u
mkdir -p gpurun_out
cd scripts && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/launch_latency launch_latency.cu 2>&1 | tail -3
for m in 0 1 2; do /tmp/launch_latency $m; done | tee ../gpurun_out/r2x_launch_latency.jsonl
