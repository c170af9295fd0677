#!/bin/bash
# round 2 (second session), multi-GPU call: sharded parity (mgpu_check, 4-limb cases through the new kernels) + sharded BLS proof + bench
# usage: gpurun --gpus N -- 'bash scripts/gpu_r2_multi_b.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/mgpu_check.py > gpurun_out/r2n_mgpu_check_$N.log 2>&1
tail -4 gpurun_out/r2n_mgpu_check_$N.log
BLS=52435875175126190479447740508185965837690552500527637822603658699938581184513
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --modulus $BLS --vars 26 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2n_bench_bls_$N.json 2> gpurun_out/r2n_bench_bls_$N.err
tail -c 300 gpurun_out/r2n_bench_bls_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_$N.json 2> gpurun_out/r2n_bench_$N.err
tail -c 300 gpurun_out/r2n_bench_$N.err
python - <<PY
import json
for f in ("gpurun_out/r2n_bench_bls_$N.json", "gpurun_out/r2n_bench_$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "verified", "sharded_equals_single", "scaling")})
        print("  strong:", {k: (d.get("strong_scaling") or {}).get(k) for k in ("value", "ms_per_step", "verified")})
        print("  e2e:", {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "verified")} if d.get("e2e") else None)
    except Exception as e:
        print(f, "failed", e)
PY
