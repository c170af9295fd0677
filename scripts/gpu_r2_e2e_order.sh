#!/bin/bash
# round 2: does running the CPU reference arm first (the driver's order) slow the e2e leg of the GPU arm?
set -u
mkdir -p gpurun_out
nproc > gpurun_out/r2y_host.txt; free -g >> gpurun_out/r2y_host.txt; numactl -H >> gpurun_out/r2y_host.txt 2>&1; cat /sys/kernel/mm/transparent_hugepage/enabled >> gpurun_out/r2y_host.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2y_bench_ref.json 2> gpurun_out/r2y_bench_ref.err
free -g >> gpurun_out/r2y_host.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-trait-leg > gpurun_out/r2y_bench_after_ref.json 2> gpurun_out/r2y_bench_after_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-trait-leg > gpurun_out/r2y_bench_again.json 2> gpurun_out/r2y_bench_again.err
python - <<'PY'
import json
for f in ["gpurun_out/r2y_bench_after_ref.json", "gpurun_out/r2y_bench_again.json"]:
    d = json.load(open(f)); e = d["e2e"]
    print(f, e["ms_per_step"], e["step_ms_rank0"], e["pageable_host_tables"]["step_ms_rank0"], e["upload"])
PY
