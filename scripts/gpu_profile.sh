#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU): launch list of a short bench run + full captures of the top kernels.
# ncu serialises kernel and host, which the resident kernels' mailbox handshake cannot survive, so the profiled runs use
# SCB_PAIR_RESIDENT=0 / SCB_TAIL_VARS... : every pass is an ordinary launch of the same pass body (k_pair_pass_sp instead of
# a pass of k_persist_pairs_sp) and shows up in the list.
set -u
mkdir -p gpurun_out
export SCB_PAIR_RESIDENT=0
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_pass_sp -c 2 -f -o gpurun_out/prof_pair_pass_sp $BENCH > gpurun_out/prof1.log 2>&1
ncu --set full --clock-control none -k regex:k_grid_sp -c 1 -f -o gpurun_out/prof_grid_sp $BENCH > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/launches_bench.log
# the reports embed the whole cubin (~50 MB each): keep the raw-page CSVs, drop the reports (gpurun_out/ is capped at 64 MiB)
for r in prof_pair_pass_sp prof_grid_sp; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out/
