#!/bin/bash
# round 2 profiles: ncu launch list of the bench command (pair passes as ordinary launches: ncu serialises kernel and
# host, so the resident kernels cannot run under it), and --set full captures of the kernels added in round 2
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
SCB_PAIR_RESIDENT=0 SCB_TAIL_VARS=0 timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2p_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fields > gpurun_out/r2p_bench_under_ncu.log 2>&1
cat > /tmp/prof_new.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
import thaler_study_b200 as T
F = T.Field(1572869)
m = T.DenseMultilinearExtension.synthetic(F, 24, 21)
r = list(range(3, 27))
for _ in range(2):
    m.evaluate_be(r)
F28 = T.Field(268435361)
rng = np.random.default_rng(4)
n = 1024
up = np.triu(rng.integers(0, 2, size=(n, n), dtype=np.int64), 1)
adj = up + up.T
g = T.TriangleG.new_adj_matrix(F28, 20, adj.reshape(-1).astype(bool).tolist())
T.generate_transcript(T.Prover(g))
PY
SCB_TAIL_VARS=0 timeout 900 $NCU --set full --import-source on -k regex:"k_mle_eval_fused|k_field_matmul_sp" -c 3 -o gpurun_out/r2p_new_kernels python /tmp/prof_new.py > gpurun_out/r2p_ncu_new.log 2>&1
ncu -i gpurun_out/r2p_new_kernels.ncu-rep --page raw --csv > gpurun_out/r2p_ncu_new_raw.csv 2>/dev/null
SCB_TAIL_VARS=0 timeout 600 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r2p_launches_triangle_mle.csv python /tmp/prof_new.py > /dev/null 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2p_launches_gkr.csv python scripts/bench_gkr.py --depth 2 > /dev/null 2>&1
for v in 25 26 27; do timeout 300 python bench.py --vars $v --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-fields > gpurun_out/r2p_bench_1gpu_v$v.json 2>/dev/null; done
rm -f gpurun_out/r2p_new_kernels.ncu-rep gpurun_out/r2c_g4.ncu-rep
