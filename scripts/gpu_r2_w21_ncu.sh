#!/bin/bash
# round 2, third session: ncu launch list of the bench command (pair passes as ordinary launches: ncu serialises kernel and host,
# so the resident kernels cannot run under it) and --set full captures of the two 21-bit-triple kernels
set -u
mkdir -p gpurun_out
P=gpurun_out/r2f
NCU="ncu --clock-control none"
SCB_PAIR_RESIDENT=0 timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file ${P}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fields > ${P}_bench_under_ncu.log 2>&1
SCB_PAIR_RESIDENT=0 timeout 600 $NCU --set full --import-source on -k regex:"k_grid_sp_pf_w21|k_pair_pass_sp_w21" -c 4 -o ${P}_w21 python scripts/kbench_w21_once.py > ${P}_ncu_w21.log 2>&1
tail -3 ${P}_ncu_w21.log
ncu -i ${P}_w21.ncu-rep --page raw --csv > ${P}_ncu_w21_raw.csv 2>/dev/null
rm -f ${P}_w21.ncu-rep
grep -o "k_[a-z0-9_]*" ${P}_launches_bench.csv | sort | uniq -c
