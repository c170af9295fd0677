#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU): launch list of a short bench run + full capture of the top kernels
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fold_round_sp -c 3 -f -o gpurun_out/prof_fold_round_sp $BENCH > gpurun_out/prof1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_round_evals -c 1 -f -o gpurun_out/prof_round_evals $BENCH > gpurun_out/prof2.log 2>&1
ls -la gpurun_out
