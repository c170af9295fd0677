#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/kbench_pairs.py 28 3 2>&1 | tee gpurun_out/kbench_pairs.txt
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_grid_sp|k_pair_pass_sp" -c 4 -o gpurun_out/ncu_pairs -f python scripts/kbench_pairs.py 28 3 > gpurun_out/ncu_pairs.log 2>&1
tail -3 gpurun_out/ncu_pairs.log
