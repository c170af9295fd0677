"""Narrowing upload (scb_poly_product_from_host alone, three pinned 2^v-entry host tables -> packed handle) against the
software-prefetch distance of the pack threads (option host_pack_prefetch, csrc/host/hostpack.hpp).  One JSON line per setting."""
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


def run(tag, reps=3, **opts):
    T.reset_options()
    for k_, val in opts.items():
        T.set_option(k_, val)
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
    print(json.dumps({"tag": tag, **opts, "ms_min": round(min(ts[1:]), 2), "ms_all": [round(t, 1) for t in ts[1:]],
                      "chunks_host": a.value, "chunks_device": b.value, "h2d_GB": round(c.value / 1e9, 3)}), flush=True)


run("plain", host_pack=0)
for rep in range(2):
    for pf in (0, 1024, 2048, 4096, 8192, 16384):
        run("prefetch", host_pack_prefetch=pf)
for pf in (4096, 8192):
    for slots in (2, 6):
        run("prefetch x raw lane depth", host_pack_prefetch=pf, host_pack_raw_slots=slots)
    run("prefetch + nt stores", host_pack_prefetch=pf, host_pack_nt=1)
