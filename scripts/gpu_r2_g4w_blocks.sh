#!/bin/bash
# round 2: resident CTAs per SM of the fourth-generation 4-limb kernels (option g4_blocks4), p = 1 variant on / off
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_g4.py -m gpu -x -q 2>&1 | tail -3
BLS=52435875175126190479447740508185965837690552500527637822603658699938581184513
for cfg in "0 1" "0 0" "2 1"; do set -- $cfg
SCB_G4_BLOCKS4=$1 SCB_G4_P0ONE=$2 timeout 600 python bench.py --modulus $BLS --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2w_bls_blocks$1_p$2.json 2> gpurun_out/r2w_bls_blocks$1_p$2.err
python - $1 $2 <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2w_bls_blocks%s_p%s.json"%(sys.argv[1],sys.argv[2])))
print("g4_blocks4",sys.argv[1],"p0one",sys.argv[2],"ms/proof",round(d["ms_per_step"],2),"verified",d.get("verified"),"kernel_ms",round(d["roofline"]["kernel_ms"],2))
PY
done
