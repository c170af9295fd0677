#!/bin/bash
# round 2, third session: sharded parity (mgpu_check) + bench (weak headline + strong-scaling record) on N GPUs with the final
# library (21-bit triples, prefetching pack threads).  usage: gpurun --gpus N -- 'bash scripts/gpu_r2_multi_d.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/mgpu_check.py > gpurun_out/r2g_mgpu_check_$N.log 2>&1
tail -3 gpurun_out/r2g_mgpu_check_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_$N.json 2> gpurun_out/r2g_bench_$N.err
tail -c 400 gpurun_out/r2g_bench_$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2g_bench_$N.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "verified", "sharded_equals_single", "scaling")})
print("strong:", {k: (d.get("strong_scaling") or {}).get(k) for k in ("value", "ms_per_step", "verified")})
print("e2e:", {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "verified")} if d.get("e2e") else None)
print("roofline:", d["roofline"]["kernel"][:50], d["roofline"]["frac"], d["roofline"].get("w21_triples"))
PY
