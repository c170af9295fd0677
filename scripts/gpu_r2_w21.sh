#!/bin/bash
# round 2, third session: option pair_w21 -- parity (A/B tests) and timing on the headline workload
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pairs.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2w_pytest_pairs.log
tail -3 gpurun_out/r2w_pytest_pairs.log
timeout 300 python scripts/kbench_w21.py 28 > gpurun_out/r2w_w21_ab.jsonl 2> gpurun_out/r2w_w21_ab.err
cat gpurun_out/r2w_w21_ab.jsonl; tail -3 gpurun_out/r2w_w21_ab.err
timeout 120 python scripts/kbench_w21.py 26 > gpurun_out/r2w_w21_ab_v26.jsonl 2>&1
g++ -O3 -march=native -pthread scripts/host_pack_bench.cpp -o /tmp/host_pack_bench && timeout 200 /tmp/host_pack_bench 16 28 > gpurun_out/r2w_host_pack_bench.jsonl 2>&1
