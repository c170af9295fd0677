#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for cfg in "1 8" "1 12" "1 16" "2 8" "4 4" "4 3" "4 5" "4 8"; do set -- $cfg; SCB_QP32=$1 SCB_BPS32=$2 python scripts/kbench.py --iters 20 >> gpurun_out/sweep.jsonl 2>&1; done
for v in 27 26 25; do python scripts/kbench.py --iters 20 --vars $v >> gpurun_out/sweep.jsonl 2>&1; done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['vars'], d['env'], {k: round(v['GBs']) for k, v in d.items() if isinstance(v, dict) and 'GBs' in v})
PY
