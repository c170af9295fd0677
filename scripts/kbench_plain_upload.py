"""Times the plain (8-byte) upload path per call: scb_mle_from_host from pinned and from pageable memory, with the
option local_ranks = 2 (what a sharded run sets, which keeps the narrowing upload off)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thaler_study_b200 as T
from thaler_study_b200._lib import check, lib

v = int(os.environ.get("KB_VARS", "28"))
F = T.Field(1572869)
m = T.DenseMultilinearExtension.synthetic(F, v, 900)
d = torch.empty([1 << v, 1], dtype=torch.int64, device="cuda")
check(lib.scb_mle_copy_to_device(m._h, d.data_ptr()))
h = torch.empty([1 << v, 1], dtype=torch.int64, pin_memory=True)
h.copy_(d)
pinned = h.numpy().view(np.uint64)
pageable = np.array(pinned, copy=True)
del d
for lr in (1, 2):
    T.set_option("local_ranks", lr)
    for name, tab in (("pinned", pinned), ("pageable", pageable)):
        ts = []
        for i in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x = T.DenseMultilinearExtension.from_evaluations_vec(F, v, tab)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
            del x
        print(json.dumps({"local_ranks": lr, "memory": name, "vars": v, "ms": [round(t, 1) for t in ts]}), flush=True)
T.reset_options()
