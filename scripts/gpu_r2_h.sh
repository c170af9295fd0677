#!/bin/bash
# round 2, GPU call H: GKR / triangle after the loads + grid fixes, r02 launch list of the bench path
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gkr.py tests/test_gpu_triangle.py tests/test_gpu_fullsize.py -k "gkr or triangle" -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2h_pytest.log
tail -3 gpurun_out/r2h_pytest.log
timeout 600 python scripts/bench_gkr.py > gpurun_out/r2h_gkr.json 2> gpurun_out/r2h_gkr.err
timeout 900 python scripts/bench_configs.py > gpurun_out/r2h_configs.jsonl 2> gpurun_out/r2h_configs.err
SCB_PAIR_RESIDENT=0 timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2h_launches_bench_pairs.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fields > gpurun_out/r2h_bench_under_ncu.log 2>&1
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2h_launches_gkr.csv python scripts/bench_gkr.py --depth 2 > /dev/null 2>&1
