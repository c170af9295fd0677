#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU): launch list of a short bench run + full capture of the top kernels.
# SCB_TAIL_VARS=0: ncu serialises kernel and host, which the resident tail kernel's mailbox handshake cannot survive
# (it would give up after 250 ms and fall back); with 0 every round is an ordinary launch and shows up in the list.
set -u
mkdir -p gpurun_out
export SCB_TAIL_VARS=0
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fold_round_sp -c 3 -f -o gpurun_out/prof_fold_round_sp $BENCH > gpurun_out/prof1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_round_evals -c 1 -f -o gpurun_out/prof_round_evals $BENCH > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/launches_bench.log
