"""MLE evaluation (SURVEY 8 a8/a10, configs[1]): device time at 2^24 and 2^28 entries, one launch (option mle_fused = 1,
default) against round 1's two launches (mle_fused = 0).  The timed span is the raw C-ABI call (scb_mle_evaluate_be
with the point already in Montgomery limbs): first kernel start .. result on the host.  Prints one JSON object per
case with the roofline of the call (algorithmic bytes = the table, read once).  Run plain for CUDA-event
spans, under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import thaler_study_b200 as T  # noqa: E402
from thaler_study_b200._lib import check, lib, u64p  # noqa: E402

reps = int(os.environ.get("KB_REPS", "30"))
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
for p, vs in ((1572869, (24, 28)), (0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001, (24,)), (0xFFFFFFFF00000001, (24,))):
    F = T.Field(p)
    for v in vs:
        m = T.DenseMultilinearExtension.synthetic(F, v, 21)
        rng = np.random.default_rng(1)
        r = [int(x) % p for x in rng.integers(0, 2**62, size=v)]
        pt = F.to_mont(r)
        out = np.zeros((1, F.n), dtype=np.uint64)
        results = {}
        for fused in (0, 1):
            T.set_option("mle_fused", fused)
            want = m.evaluate_be(r)
            evs, walls = [], []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                a.record()
                check(lib.scb_mle_evaluate_be(m._h, pt.ctypes.data_as(u64p), v, out.ctypes.data_as(u64p)))
                b.record()
                torch.cuda.synchronize()
                walls.append((time.perf_counter() - t0) * 1e3)
                evs.append(a.elapsed_time(b))
                assert F.from_mont(out)[0] == want
            results[fused] = (want, sorted(evs)[len(evs) // 2], sorted(walls)[len(walls) // 2])
        assert results[0][0] == results[1][0]
        nbytes = (1 << v) * 8 * F.n
        for fused in (0, 1):
            _, ev, wall = results[fused]
            print(json.dumps({"field_bits": F.bits, "vars": v, "launches": 1 if fused else 2, "algorithmic_bytes": nbytes, "event_span_ms": round(ev, 4),
                              "call_wall_ms": round(wall, 4),
                              "roofline": {"bound": "hbm", "achieved": round(nbytes / ev / 1e6, 1), "peak": PEAK, "unit": "GB/s", "frac": round(nbytes / ev / 1e6 / PEAK, 3)}}), flush=True)
        del m
T.reset_options()
