#!/bin/bash
# Two rounds per pass: parity tests + A/B timing + trace.  -> gpurun_out/
set -u
mkdir -p gpurun_out
echo "== pairs parity"
timeout 1200 python -m pytest tests/test_gpu_pairs.py tests/test_gpu_parity.py tests/test_gpu_packed.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_pairs.log
echo "== bench pairs (default)"
SCB_PERSIST_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_pairs.err | tee gpurun_out/bench_pairs.json
grep -E "pairs m|persist m" gpurun_out/bench_pairs.err | tail -2
tail -3 gpurun_out/bench_pairs.err | grep -v "pairs m"
echo "== bench one round per pass (SCB_PAIRS=0)"
SCB_PAIRS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>>gpurun_out/bench_pairs.err | tee gpurun_out/bench_nopairs.json
