#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for cfg in "1 2" "1 3" "1 4" "1 5" "2 2" "2 3" "2 4" "2 5" "1 4" "1 3"; do
  set -- $cfg
  SCB_UNROLL=$1 SCB_BPS=$2 python scripts/kbench.py --iters 20 >> gpurun_out/sweep.jsonl 2>&1
done
for v in 27 26 24 22 20; do SCB_BPS=4 python scripts/kbench.py --iters 20 --vars $v >> gpurun_out/sweep.jsonl 2>&1; done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['env'], d['vars'], "round %.0f GB/s  fold %.0f GB/s (min-time %.0f)" % (d['round_evals_GBs'], d['fold_round_GBs'], d['fold_round_GBs']*d['fold_round_ms']/d['fold_round_min_ms']))
PY
