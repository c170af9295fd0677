"""Sweep of the packed upload's switches on one GPU: time scb_poly_product_from_host alone (three pinned 2^v-entry host
tables -> packed handle), per chunk size / pack threads / lanes / store kind.  Prints one JSON line per setting."""
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
ref = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 900 + k) for k in range(K)]).round_evals()


def run(tag, reps=3, **env):
    keys = ["SCB_HOST_PACK", "SCB_HOST_PACK_THREADS", "SCB_HOST_PACK_CHUNK_LOG2", "SCB_HOST_PACK_RAW", "SCB_HOST_PACK_NT", "SCB_HOST_PACK_WIRE"]
    for k_ in keys:
        os.environ.pop(k_, None)
    for k_, val in env.items():
        os.environ[k_] = str(val)
    ts = []
    for i in range(reps + 1):
        t0 = time.perf_counter()
        g = T.ProductMLE.from_host_tables(F, v, host)
        ts.append((time.perf_counter() - t0) * 1e3)
        if i == 0:
            assert g.round_evals() == ref, tag
        del g
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    check(lib.scb_host_pack_stats(C.byref(a), C.byref(b), C.byref(c)))
    print(json.dumps({"tag": tag, **env, "ms_min": round(min(ts[1:]), 2), "ms_all": [round(t, 1) for t in ts[1:]],
                      "chunks_host": a.value, "chunks_device": b.value, "h2d_GB": round(c.value / 1e9, 3)}), flush=True)


run("plain", SCB_HOST_PACK=0)
run("defaults")
for wire in (21, 32):
    for nt in (1, 0):
        for cl in (19, 20, 21, 22):
            for th in (8, 12, 16):
                run("sweep", SCB_HOST_PACK_WIRE=wire, SCB_HOST_PACK_CHUNK_LOG2=cl, SCB_HOST_PACK_THREADS=th, SCB_HOST_PACK_NT=nt)
for wire in (21, 32):
    for cl in (20, 21):
        run("host lane only", SCB_HOST_PACK_WIRE=wire, SCB_HOST_PACK_CHUNK_LOG2=cl, SCB_HOST_PACK_RAW=0)
