#!/bin/bash
# round 2, third session: compute-sanitizer over the 21-bit-triple kernels
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
rm -f gpurun_out/r2d_sanitize_summary.txt
for tool in memcheck racecheck synccheck initcheck; do
  timeout 400 $CS --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_w21.py > gpurun_out/r2d_sanitize_w21_${tool}.log 2>&1
  echo "w21 $tool rc=$?" | tee -a gpurun_out/r2d_sanitize_summary.txt
  tail -2 gpurun_out/r2d_sanitize_w21_${tool}.log
done
