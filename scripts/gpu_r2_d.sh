#!/bin/bash
# round 2, GPU call D: g29 kernel tests + A/B timing of the three 4-limb kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_g4.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2d_pytest_g4.log
tail -3 gpurun_out/r2d_pytest_g4.log
BLS=52435875175126190479447740508185965837690552500527637822603658699938581184513
for k in 2 1; do
  SCB_G4_KERNEL=$k timeout 600 python bench.py --modulus $BLS --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2d_bench_bls_k$k.json 2> gpurun_out/r2d_bench_bls_k$k.err
done
for bps in 1; do
  SCB_G4_KERNEL=2 SCB_BPS=$bps timeout 600 python bench.py --modulus $BLS --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2d_bench_bls_k2_bps$bps.json 2>/dev/null
done
