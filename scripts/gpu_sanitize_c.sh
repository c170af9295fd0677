#!/bin/bash
# round 2 (second session, after the GKR work): compute-sanitizer over the default workload (incl. the GKR scatter / tail kernels and the
# single-CTA finish) and the GKR GPU tests themselves under memcheck + racecheck
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
rm -f gpurun_out/r2c_sanitize_summary.txt
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_run.py > gpurun_out/r2c_sanitize_all_${tool}.log 2>&1
  echo "all $tool rc=$?" | tee -a gpurun_out/r2c_sanitize_summary.txt
  tail -2 gpurun_out/r2c_sanitize_all_${tool}.log
done
for tool in memcheck racecheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 --error-exitcode 9 python -m pytest tests/test_gpu_gkr.py -m gpu -x -q -k "not full" > gpurun_out/r2c_sanitize_gkrtests_${tool}.log 2>&1
  echo "gkr tests $tool rc=$?" | tee -a gpurun_out/r2c_sanitize_summary.txt
  tail -3 gpurun_out/r2c_sanitize_gkrtests_${tool}.log
done
