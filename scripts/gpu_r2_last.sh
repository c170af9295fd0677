#!/bin/bash
# round 2, third session: default bench.py once more (10 e2e steps) and the ncu launch list + --set full captures with the final library
set -u
mkdir -p gpurun_out
P=gpurun_out/r2j
timeout 600 python bench.py > ${P}_bench_default.json 2> ${P}_bench_default.err
tail -c 300 ${P}_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2j_bench_default.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "verified", "steps", "warmup")}, "e2e", d["e2e"]["ms_per_step"], d["e2e"]["step_ms_rank0"])
PY
NCU="ncu --clock-control none"
SCB_PAIR_RESIDENT=0 timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file ${P}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fields > ${P}_bench_under_ncu.log 2>&1
SCB_PAIR_RESIDENT=0 timeout 600 $NCU --set full --import-source on -k regex:"k_pair_pass_sp_w21" -c 2 -o ${P}_w21 python scripts/kbench_w21_once.py > ${P}_ncu_w21.log 2>&1
ncu -i ${P}_w21.ncu-rep --page raw --csv > ${P}_ncu_w21_raw.csv 2>/dev/null
rm -f ${P}_w21.ncu-rep
grep -o "k_[a-z0-9_]*" ${P}_launches_bench.csv | sort | uniq -c
