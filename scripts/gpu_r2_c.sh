#!/bin/bash
# round 2, GPU call C: carry-chain issue rates, triangle/trait tests, configs bench, ncu of the g4 kernel, sanitizers
set -u
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/imad_chain scripts/imad_chain.cu && /tmp/imad_chain > gpurun_out/r2c_imad_chain.jsonl 2>&1
cat gpurun_out/r2c_imad_chain.jsonl
timeout 900 python -m pytest tests/test_gpu_triangle.py tests/test_gpu_trait_path.py tests/test_gpu_gkr.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
timeout 900 python scripts/bench_configs.py > gpurun_out/r2c_configs.jsonl 2> gpurun_out/r2c_configs.err
tail -c 400 gpurun_out/r2c_configs.err
cat > /tmp/g4prof.py <<'PY'
import sys; sys.path.insert(0, '.')
import thaler_study_b200 as T
p = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
F = T.Field(p)
g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, 24, 0xB200 + k) for k in range(3)])
claim = T.evals_to_univariate(F, T.KIND_PRODUCT, g.round_evals()).evaluate(12345)
for _ in range(2):
    g.fix_and_round_evals(12345, claim=claim)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fold_round_g4 -c 1 -o gpurun_out/r2c_g4 python /tmp/g4prof.py > gpurun_out/r2c_ncu_g4.log 2>&1
ncu -i gpurun_out/r2c_g4.ncu-rep --page raw --csv > gpurun_out/r2c_ncu_g4_raw.csv 2>/dev/null
bash scripts/gpu_sanitize.sh
