#!/bin/bash
# round 2, multi-GPU call: sharded parity (mgpu_check) + bench (weak headline + strong-scaling record) on N GPUs
# usage: gpurun --gpus N -- 'bash scripts/gpu_r2_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2m_ngpus_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/mgpu_check.py > gpurun_out/r2m_mgpu_check_$N.log 2>&1
tail -4 gpurun_out/r2m_mgpu_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_$N.json 2> gpurun_out/r2m_bench_$N.err
tail -c 400 gpurun_out/r2m_bench_$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2m_bench_$N.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "verified", "sharded_equals_single", "scaling")})
print("strong:", d.get("strong_scaling"))
print("e2e:", {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "verified")} if d.get("e2e") else None)
PY
SCB_PERSIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --scaling strong --no-strong 2> gpurun_out/r2m_trace_strong_$N.err > gpurun_out/r2m_bench_strong_$N.json
grep -E "^\[pairs|^\[persist" gpurun_out/r2m_trace_strong_$N.err | tail -4
timeout 600 python bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2m_ref_$N.json 2>/dev/null
