#!/bin/bash
# round 2, third session: the round-end sequence with the final library (21-bit triples, prefetching pack threads), then the
# ncu launch list of the bench command and --set full captures of the two new kernels
set -u
mkdir -p gpurun_out
P=gpurun_out/r2f
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12 > ${P}_pytest.log
tail -3 ${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -1 ${P}_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > ${P}_bench_ref.json 2> ${P}_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err
tail -c 300 ${P}_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2f_bench.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "verified")}, "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("upload"), "roof", d["roofline"]["kernel"][:40], d["roofline"]["frac"])
PY
timeout 600 python scripts/kbench_upload_pf.py > ${P}_upload_pf.jsonl 2> ${P}_upload_pf.err; cat ${P}_upload_pf.jsonl
timeout 600 python scripts/bench_gkr.py > ${P}_gkr.json 2> ${P}_gkr.err
timeout 900 python scripts/bench_configs.py > ${P}_configs.jsonl 2> ${P}_configs.err
NCU="ncu --clock-control none"
SCB_PAIR_RESIDENT=0 SCB_TAIL_VARS=0 timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file ${P}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fields > ${P}_bench_under_ncu.log 2>&1
SCB_PAIR_RESIDENT=0 SCB_TAIL_VARS=0 timeout 600 $NCU --set full --import-source on -k regex:"k_grid_sp_pf_w21|k_pair_pass_sp_w21" -c 2 -o ${P}_w21 python scripts/kbench_w21_once.py > ${P}_ncu_w21.log 2>&1
ncu -i ${P}_w21.ncu-rep --page raw --csv > ${P}_ncu_w21_raw.csv 2>/dev/null
rm -f ${P}_w21.ncu-rep
