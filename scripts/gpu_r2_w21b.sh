#!/bin/bash
# round 2, third session: bilinear two-variable fold in the triple pair pass -- parity (A/B test over the five settings) and timing
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pairs.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2h_pytest_pairs.log
tail -3 gpurun_out/r2h_pytest_pairs.log
timeout 300 python scripts/kbench_w21.py 28 > gpurun_out/r2h_w21_ab.jsonl 2> gpurun_out/r2h_w21_ab.err
cat gpurun_out/r2h_w21_ab.jsonl; tail -3 gpurun_out/r2h_w21_ab.err
