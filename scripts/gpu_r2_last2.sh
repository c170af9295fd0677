#!/bin/bash
# round 2, third session: whole -m gpu suite, smoke and the default bench.py line with the last library build
set -u
mkdir -p gpurun_out
P=gpurun_out/r2k
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12 > ${P}_pytest.log
tail -3 ${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -1 ${P}_smoke.log
timeout 600 python bench.py > ${P}_bench_default.json 2> ${P}_bench_default.err
tail -c 300 ${P}_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2k_bench_default.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "verified", "steps", "warmup")}, "e2e", d["e2e"]["ms_per_step"], d["e2e"]["step_ms_rank0"]["timed_ms"], "trait", d["e2e"]["trait_only"]["ms_per_step"])
PY
