#!/bin/bash
# Grid-wide resident kernel: parity + A/B timing against one launch per round.  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
echo "== persist parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resident or tail or large or transcript" 2>&1 | tail -8 | tee gpurun_out/pytest_persist.log
echo "== bench persistent (default)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_p.err | tee gpurun_out/bench_persist.json
tail -3 gpurun_out/bench_p.err
echo "== bench per-round launches + single-CTA tail (SCB_PERSIST_VARS=0)"
SCB_PERSIST_VARS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>>gpurun_out/bench_p.err | tee gpurun_out/bench_nopersist.json
echo "== goldilocks persistent vs not"
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --modulus 18446744069414584321 2>>gpurun_out/bench_p.err | tee gpurun_out/bench_gold_persist.json
SCB_PERSIST_VARS=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --modulus 18446744069414584321 2>>gpurun_out/bench_p.err | tee gpurun_out/bench_gold_nopersist.json
