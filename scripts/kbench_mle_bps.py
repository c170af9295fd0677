import ctypes as C, time, sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
import thaler_study_b200 as T
from thaler_study_b200._lib import check, lib, u64p
F = T.Field(1572869)
v = 24
m = T.DenseMultilinearExtension.synthetic(F, v, 21)
pt = F.to_mont([i * 7 + 1 for i in range(v)])
out = np.zeros((1, F.n), dtype=np.uint64)
for bps in (0, 1, 2, 3, 4, 6, 8):
    T.set_option("bps", bps)
    for _ in range(20): check(lib.scb_mle_evaluate_be(m._h, pt.ctypes.data_as(u64p), v, out.ctypes.data_as(u64p)))
    ts = []
    for _ in range(300):
        t0 = time.perf_counter()
        lib.scb_mle_evaluate_be(m._h, pt.ctypes.data_as(u64p), v, out.ctypes.data_as(u64p))
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print("bps", bps, "median call us", round(ts[150] * 1e6, 1), "min", round(ts[0] * 1e6, 1))
