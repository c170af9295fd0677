#!/bin/bash
# round 2: warm per-kernel durations of the GKR prover (ncu without cache flushes) next to the wall time of the same run
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/r2x_launches_gkr_warm.csv python scripts/bench_gkr.py --depth 4 > gpurun_out/r2x_gkr_under_ncu.log 2>&1
tail -2 gpurun_out/r2x_gkr_under_ncu.log | cut -c1-300
timeout 300 python scripts/bench_gkr.py --depth 4 2>/dev/null | cut -c1-900
