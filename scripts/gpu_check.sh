#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, smoke, bench.  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu" 
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
echo "== bench BLS12-381 Fr (4 limbs), 2^28"
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --modulus 52435875175126190479447740508185965837690552500527637822603658699938581184513 2>>gpurun_out/bench.err | tee gpurun_out/bench_bls.json
echo "== bench Goldilocks (1 limb, 64-bit), 2^28"
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --modulus 18446744069414584321 2>>gpurun_out/bench.err | tee gpurun_out/bench_goldilocks.json
