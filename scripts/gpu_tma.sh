#!/bin/bash
set -u
mkdir -p gpurun_out
SCB_GRID_TMA=1 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_grid_sp_tma" -c 2 -o gpurun_out/ncu_tma -f python scripts/kbench_pairs.py 28 3 > gpurun_out/ncu_tma.log 2>&1
tail -3 gpurun_out/ncu_tma.log
