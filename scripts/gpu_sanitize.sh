#!/bin/bash
# compute-sanitizer over one small instance of every kernel family (scripts/sanitize_run.py); logs -> gpurun_out/
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_run.py --variants > gpurun_out/r2_sanitize_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r2_sanitize_summary.txt
  tail -4 gpurun_out/r2_sanitize_${tool}.log
done
for tool in memcheck racecheck synccheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_run.py --resident > gpurun_out/r2_sanitize_${tool}_resident.log 2>&1
  echo "$tool resident rc=$?" | tee -a gpurun_out/r2_sanitize_summary.txt
  tail -4 gpurun_out/r2_sanitize_${tool}_resident.log
done
