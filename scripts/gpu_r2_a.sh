#!/bin/bash
# round 2, GPU call A: the whole -m gpu suite, the default bench (both arms), the MLE kernel bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt; free -g >> gpurun_out/r2a_nproc.txt
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python scripts/kbench_mle.py > gpurun_out/r2a_kbench_mle.jsonl 2> gpurun_out/r2a_kbench_mle.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
tail -2 gpurun_out/r2a_smoke.log
