#!/bin/bash
# round 2: ncu --set full of the one-launch MLE evaluation at 2^24 entries (one-limb field)
set -u
mkdir -p gpurun_out
KB_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_mle_eval_fused" -c 3 -f -o gpurun_out/prof_mle python scripts/kbench_mle.py > gpurun_out/prof_mle.log 2>&1
ncu -i gpurun_out/prof_mle.ncu-rep --page raw --csv > gpurun_out/prof_mle_raw.csv 2>/dev/null
rm -f gpurun_out/prof_mle.ncu-rep
ls -la gpurun_out/prof_mle*
