#!/bin/bash
# round 2: ncu --set full of the GKR gate-list kernels (one launch each) at width 2^20
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --cache-control none -k regex:"k_gkr_phase1|k_gkr_phase2|k_gkr_wiring_eval|k_pqs_multi" -s 4 -c 6 -f -o gpurun_out/prof_gkr python scripts/bench_gkr.py --depth 2 > gpurun_out/prof_gkr.log 2>&1
ncu -i gpurun_out/prof_gkr.ncu-rep --page raw --csv > gpurun_out/prof_gkr_raw.csv 2>/dev/null
rm -f gpurun_out/prof_gkr.ncu-rep
