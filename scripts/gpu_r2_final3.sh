#!/bin/bash
# round 2, third session: the round-end sequence with the final library (21-bit triples + bilinear fold, stand-alone first pair
# pass from 2^24 entries, prefetching pack threads)
set -u
mkdir -p gpurun_out
P=gpurun_out/r2i
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12 > ${P}_pytest.log
tail -3 ${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -1 ${P}_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > ${P}_bench_ref.json 2> ${P}_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err
tail -c 300 ${P}_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2i_bench.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "verified")}, "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("upload"), "roof", d["roofline"]["kernel"][:40], d["roofline"]["frac"])
PY
timeout 600 python scripts/bench_gkr.py > ${P}_gkr.json 2> ${P}_gkr.err
timeout 900 python scripts/bench_configs.py > ${P}_configs.jsonl 2> ${P}_configs.err
