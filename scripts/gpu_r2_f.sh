#!/bin/bash
# round 2, GPU call F: GKR multi-round passes (tests + bench), plain-upload timing, upload sweep, reference-shaped sweep
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gkr.py tests/test_gpu_fullsize.py -k "gkr" -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2f_pytest_gkr.log
tail -3 gpurun_out/r2f_pytest_gkr.log
timeout 600 python scripts/bench_gkr.py > gpurun_out/r2f_gkr.json 2> gpurun_out/r2f_gkr.err
SCB_GKR_MULTI=0 timeout 600 python scripts/bench_gkr.py > gpurun_out/r2f_gkr_nomulti.json 2>/dev/null
timeout 300 python scripts/kbench_plain_upload.py > gpurun_out/r2f_plain_upload.jsonl 2>&1
cat gpurun_out/r2f_plain_upload.jsonl
timeout 900 python scripts/kbench_upload.py > gpurun_out/r2f_upload_sweep.jsonl 2> gpurun_out/r2f_upload_sweep.err
timeout 900 python scripts/bench_mm_sweep.py > gpurun_out/r2f_mm_sweep.jsonl 2> gpurun_out/r2f_mm_sweep.err
tail -2 gpurun_out/r2f_mm_sweep.jsonl
