#!/bin/bash
# Per-round device time / host turn-around of the grid-wide resident kernel (SCB_PERSIST_TRACE=1).  -> gpurun_out/
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py -m gpu -x -q -k "resident or tail or large or transcript or packed" 2>&1 | tail -5
SCB_PERSIST_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/trace_sp.err | tee gpurun_out/trace_sp.json
grep persist gpurun_out/trace_sp.err | tail -2
SCB_PERSIST_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --modulus 18446744069414584321 2>gpurun_out/trace_g1.err | tee gpurun_out/trace_g1.json
grep persist gpurun_out/trace_g1.err | tail -2
