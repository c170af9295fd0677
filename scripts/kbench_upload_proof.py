"""Where does the proof spend its time on a handle made by the packed upload (uint32 tables from the start) compared
with the caller's 8-byte tables?  One GPU; prints JSON lines."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thaler_study_b200 as T  # noqa: E402
from thaler_study_b200._lib import check, lib  # noqa: E402

v = int(os.environ.get("KB_VARS", "28"))
K, p = 3, 1572869
F = T.Field(p)
host = []
for k in range(K):
    m = T.DenseMultilinearExtension.synthetic(F, v, 900 + k)
    d = torch.empty([1 << v, 1], dtype=torch.int64, device="cuda")
    check(lib.scb_mle_copy_to_device(m._h, d.data_ptr()))
    h = torch.empty([1 << v, 1], dtype=torch.int64, pin_memory=True)
    h.copy_(d)
    host.append(h.numpy().view(np.uint64))
    del d, m
torch.cuda.synchronize()


def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
        del r
    return [round(t, 3) for t in ts[2:]]


plain = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 900 + k) for k in range(K)])
packed = T.ProductMLE.from_host_tables(F, v, host)
want = T.generate_transcript(T.Prover(plain))
assert T.generate_transcript(T.Prover(packed)) == want
for name, g in (("plain 8-byte tables", plain), ("packed upload handle", packed)):
    print(json.dumps({"handle": name, "Prover::new ms": timed(lambda: T.Prover(g)),
                      "new + transcript ms": timed(lambda: T.generate_transcript(T.Prover(g))),
                      "grid_evals ms": timed(lambda: g.grid_evals()),
                      "pair_pass ms": timed(lambda: g.pair_pass(5, 7))}), flush=True)
print(json.dumps({"upload + drop ms": timed(lambda: T.ProductMLE.from_host_tables(F, v, host))}))
print(json.dumps({"upload + proof ms": timed(lambda: T.generate_transcript(T.Prover(T.ProductMLE.from_host_tables(F, v, host))))}))
