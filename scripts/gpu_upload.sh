#!/bin/bash
# Packed upload (scb_poly_product_from_host): parity tests, then the e2e bench with the upload lanes switched three ways.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/upload_nproc.txt; lscpu | grep -E "Model name|Socket|^CPU\(s\)|NUMA node\(s\)" >> gpurun_out/upload_nproc.txt
echo "== upload parity"
timeout 300 python -m pytest tests/test_gpu_upload.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_upload.log
echo "== bench (default: both lanes)"
timeout 240 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_upload.err | tee gpurun_out/bench_upload.json | cut -c1-300
echo "== bench, plain copies (SCB_HOST_PACK=0)"
SCB_HOST_PACK=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_upload.err | tee gpurun_out/bench_upload_plain.json | cut -c1-200
echo "== bench, host lane only (SCB_HOST_PACK_RAW=0)"
SCB_HOST_PACK_RAW=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_upload.err | tee gpurun_out/bench_upload_noraw.json | cut -c1-200
echo "== bench, 8 pack threads"
SCB_HOST_PACK_THREADS=8 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_upload.err | tee gpurun_out/bench_upload_t8.json | cut -c1-200
for f in gpurun_out/bench_upload*.json; do python - "$f" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"]), "e2e", json.dumps(d["e2e"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
P
done
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== packed / pairs parity"
timeout 400 python -m pytest tests/test_gpu_packed.py tests/test_gpu_pairs.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_packed_pairs.log
tail -5 gpurun_out/bench_upload.err
