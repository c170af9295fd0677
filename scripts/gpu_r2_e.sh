#!/bin/bash
# round 2, GPU call E: 4-limb kernels, variants (kernel generation x resident CTAs per SM)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_g4.py tests/test_gpu_trait_path.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2e_pytest_g4.log
tail -3 gpurun_out/r2e_pytest_g4.log
BLS=52435875175126190479447740508185965837690552500527637822603658699938581184513
for k in 1 2; do for b in 2 3; do
  SCB_G4_KERNEL=$k SCB_G4_BLOCKS=$b timeout 600 python bench.py --modulus $BLS --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_bls_k${k}_b${b}.json 2> gpurun_out/r2e_bls_k${k}_b${b}.err
done; done
