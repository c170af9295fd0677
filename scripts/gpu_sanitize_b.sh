#!/bin/bash
# round 2 (second session): compute-sanitizer over the kernels changed in this session -- the 4-limb kernels of every generation
# (--only-g4) and the default workload (product cases of six fields, both MLE evaluation forms incl. g4_mle.cuh and the pipelined one-limb kernel)
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
rm -f gpurun_out/r2b_sanitize_summary.txt
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_run.py --only-g4 > gpurun_out/r2b_sanitize_g4_${tool}.log 2>&1
  echo "g4 $tool rc=$?" | tee -a gpurun_out/r2b_sanitize_summary.txt
  tail -2 gpurun_out/r2b_sanitize_g4_${tool}.log
done
for tool in memcheck racecheck synccheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_run.py > gpurun_out/r2b_sanitize_all_${tool}.log 2>&1
  echo "all $tool rc=$?" | tee -a gpurun_out/r2b_sanitize_summary.txt
  tail -2 gpurun_out/r2b_sanitize_all_${tool}.log
done
