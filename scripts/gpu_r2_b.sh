#!/bin/bash
# round 2, GPU call B: g4 kernel tests + timing, remaining tests after the first call's failure, ncu launch list of MLE eval
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_g4.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest_g4.log
tail -3 gpurun_out/r2b_pytest_g4.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_g4.py --deselect tests/test_gpu_parity.py 2>&1 | tail -15 > gpurun_out/r2b_pytest_rest.log
tail -3 gpurun_out/r2b_pytest_rest.log
timeout 600 python bench.py --modulus 52435875175126190479447740508185965837690552500527637822603658699938581184513 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench_bls.json 2> gpurun_out/r2b_bench_bls.err
tail -c 300 gpurun_out/r2b_bench_bls.err
SCB_G4_KERNEL=0 timeout 600 python bench.py --modulus 52435875175126190479447740508185965837690552500527637822603658699938581184513 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench_bls_old.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_mle_launches.csv python scripts/kbench_mle.py > gpurun_out/r2b_kbench_mle_ncu.txt 2>&1
