#!/bin/bash
set -u
mkdir -p gpurun_out
BLS=52435875175126190479447740508185965837690552500527637822603658699938581184513
SCB_TAIL_VARS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fold_round" -c 1 -f -o gpurun_out/prof_bls python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --vars 24 --modulus $BLS > gpurun_out/prof_bls.log 2>&1
ncu -i gpurun_out/prof_bls.ncu-rep --page raw --csv > gpurun_out/prof_bls_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_bls.ncu-rep --page source --csv > gpurun_out/prof_bls_src.csv 2>/dev/null
rm -f gpurun_out/prof_bls.ncu-rep
ls -la gpurun_out | tail -5
