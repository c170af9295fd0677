#!/bin/bash
# The resident kernels cannot make progress under a serialising profiler: they must give up (250 ms) and the host must
# finish the proof with ordinary launches, bytes unchanged.  Runs smoke() and a transcript comparison under ncu.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
cat > /tmp/fb.py <<'PY'
import sys; sys.path.insert(0, '.')
import thaler_study_b200 as T
for p, v, K in ((1572869, 18, 3), (1572869, 17, 2), (0xFFFFFFFF00000001, 16, 3), (1572869, 9, 3)):
    F = T.Field(p)
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 70 + k) for k in range(K)])
    t = T.generate_transcript(T.Prover(g))
    print(p, v, K, len(t), __import__('hashlib').sha256(b''.join(t)).hexdigest()[:16], T.verify_transcript(t, T.Verifier(v, g)))
PY
echo "== plain"; timeout 300 python /tmp/fb.py
echo "== under ncu"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fb_launches.csv python /tmp/fb.py 2>&1 | tail -6
grep -c "k_" gpurun_out/fb_launches.csv
