#!/bin/bash
# compute-sanitizer over the 4-limb kernels of every generation (scripts/sanitize_run.py --only-g4)
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_run.py --only-g4 > gpurun_out/r2_sanitize_g4_${tool}.log 2>&1
  echo "g4 $tool rc=$?" | tee -a gpurun_out/r2_sanitize_g4_summary.txt
  tail -3 gpurun_out/r2_sanitize_g4_${tool}.log
done
