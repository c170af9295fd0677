#!/bin/bash
# round 2: ncu --set full of the fourth-generation 4-limb kernels (2^24-entry tables, BLS12-381 Fr)
set -u
mkdir -p gpurun_out
BLS=52435875175126190479447740508185965837690552500527637822603658699938581184513
SCB_TAIL_VARS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"g4w" -c 2 -f -o gpurun_out/prof_g4w python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --vars 24 --modulus $BLS > gpurun_out/prof_g4w.log 2>&1
ncu -i gpurun_out/prof_g4w.ncu-rep --page raw --csv > gpurun_out/prof_g4w_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_g4w.ncu-rep --page source --csv > gpurun_out/prof_g4w_src.csv 2>/dev/null
rm -f gpurun_out/prof_g4w.ncu-rep
ls -la gpurun_out/prof_g4w*
