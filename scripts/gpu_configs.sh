#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== configs"
timeout 1200 python scripts/bench_configs.py 2>gpurun_out/configs.err | tee gpurun_out/configs.jsonl
tail -5 gpurun_out/configs.err
