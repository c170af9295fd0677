#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
