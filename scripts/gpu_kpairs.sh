#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/kbench_pairs.py 28 3 2>&1 | tee gpurun_out/kbench_pairs.txt
timeout 1200 python -m pytest tests/test_gpu_pairs.py -m gpu -x -q 2>&1 | tail -4
SCB_PERSIST_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_pairs.err | cut -c1-250
grep -E "pairs m" gpurun_out/bench_pairs.err | tail -2
if [ "${NCU:-0}" = "1" ]; then
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_grid_sp|k_pair_pass_sp" -c 8 -o gpurun_out/ncu_pairs -f python scripts/kbench_pairs.py 28 3 > gpurun_out/ncu_pairs.log 2>&1
tail -3 gpurun_out/ncu_pairs.log
fi
