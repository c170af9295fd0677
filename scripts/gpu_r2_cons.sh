#!/bin/bash
# strong scaling on N GPUs: sweep of the consolidation threshold
set -u
N=${1:-8}
mkdir -p gpurun_out
for cat in 12 14 16 18 20; do
  SCB_CONSOLIDATE_AUTO=$cat timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$((cat % 10)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --scaling strong --no-strong > gpurun_out/r2_cons_${N}_$cat.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_cons_${N}_$cat.json").read().strip().splitlines()[-1])
print("consolidate_at", $cat, "ms", round(d["ms_per_step"], 4), "Gelem/s", round(d["value"] / 1e3, 1), "verified", d["verified"], "launches/proof", d["gpu_launches"] / d["steps"])
PY
done
