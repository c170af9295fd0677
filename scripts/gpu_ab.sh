#!/bin/bash
# Same-box A/B of the pair kernels' load paths.  -> gpurun_out/
set -u
mkdir -p gpurun_out
for st in 0 1 0 1; do
  echo "== SCB_PAIR_STAGE=$st"
  SCB_PAIR_STAGE=$st timeout 120 python scripts/kbench_pairs.py 28 3 2>&1 | head -4
  SCB_PAIR_STAGE=$st SCB_PERSIST_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'])"
  grep -E "pairs m" gpurun_out/ab.err | tail -1
done
