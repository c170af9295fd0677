#!/bin/bash
# round 2: fourth-generation 4-limb kernels (wide accumulators, p = 1 mod 2^32 variant, fixed-multiplier folds) -- parity, then the BLS12-381 Fr proof
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_g4.py tests/test_gpu_trait_path.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2w_pytest_g4.log
tail -4 gpurun_out/r2w_pytest_g4.log
BLS=52435875175126190479447740508185965837690552500527637822603658699938581184513
for cfg in "3 1" "3 0" "1 1"; do set -- $cfg
  SCB_G4_KERNEL=$1 SCB_G4_P0ONE=$2 timeout 600 python bench.py --modulus $BLS --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2w_bls_k$1_p$2.json 2> gpurun_out/r2w_bls_k$1_p$2.err
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/r2w_bls_k%s_p%s.json"%(sys.argv[1],sys.argv[2])))
    print("g4_kernel",sys.argv[1],"p0one",sys.argv[2],"ms/proof",round(d["ms_per_step"],2),"verified",d.get("verified"),"kernel_ms",round(d["roofline"]["kernel_ms"],2))
except Exception as e: print("failed",e)
PY
done
