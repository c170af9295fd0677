"""MLE evaluation (SURVEY 8 a8/a10, configs[1]): device time of eq-table build + dot at 2^24 and 2^28 entries.
Run plain for CUDA-event spans, under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thaler_study_b200 as T  # noqa: E402

reps = int(os.environ.get("KB_REPS", "20"))
for p, vs in ((1572869, (24, 28)), (0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001, (24,))):
    F = T.Field(p)
    for v in vs:
        m = T.DenseMultilinearExtension.synthetic(F, v, 21)
        rng = np.random.default_rng(1)
        r = [int(x) % p for x in rng.integers(0, 2**62, size=v)]
        want = m.evaluate_be(r)
        evs, walls = [], []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a.record()
            got = m.evaluate_be(r)
            b.record()
            torch.cuda.synchronize()
            walls.append((time.perf_counter() - t0) * 1e3)
            evs.append(a.elapsed_time(b))
            assert got == want
        nbytes = (1 << v) * 8 * F.n
        ev = sorted(evs)[len(evs) // 2]
        print(json.dumps({"field_bits": F.bits, "vars": v, "bytes": nbytes, "event_span_ms": round(ev, 4), "GBs": round(nbytes / ev / 1e6, 1),
                          "wall_ms": round(sorted(walls)[len(walls) // 2], 4)}), flush=True)
        del m
