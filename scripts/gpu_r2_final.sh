#!/bin/bash
# round 2: what the driver runs at round end (whole -m gpu suite, smoke, both bench arms) + GKR bench
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/r2u_pytest.log
tail -4 gpurun_out/r2u_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; tail -1 gpurun_out/r2u_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2u_bench_ref.json 2> gpurun_out/r2u_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
tail -c 300 gpurun_out/r2u_bench.err
timeout 600 python scripts/bench_gkr.py > gpurun_out/r2u_gkr.json 2> gpurun_out/r2u_gkr.err
timeout 900 python scripts/bench_configs.py > gpurun_out/r2u_configs.jsonl 2> gpurun_out/r2u_configs.err
tail -c 300 gpurun_out/r2u_configs.err
