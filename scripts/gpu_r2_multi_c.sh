#!/bin/bash
# round 2: sharded 4-limb proof on N GPUs (the library picks a consolidation point that fits the peer window)
set -u
N=${1:-8}
mkdir -p gpurun_out
BLS=52435875175126190479447740508185965837690552500527637822603658699938581184513
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --modulus $BLS --vars 26 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2n_bench_bls_$N.json 2> gpurun_out/r2n_bench_bls_$N.err
grep -n "ScbError" gpurun_out/r2n_bench_bls_$N.err | head -3
python - <<PY
import json
for f in ("gpurun_out/r2n_bench_bls_$N.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "verified", "sharded_equals_single", "scaling")})
        print("  strong:", {k: (d.get("strong_scaling") or {}).get(k) for k in ("value", "ms_per_step", "verified")})
    except Exception as e:
        print(f, "failed", e)
PY
