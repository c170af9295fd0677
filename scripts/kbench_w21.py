"""A/B of option pair_w21 (21-bit triples written by the grid pass, read by the first pair pass; csrc/pairs.cuh) on the
headline workload: per-proof time and the library's own CUDA-event times of the grid pass and the first pair pass."""
import ctypes as C
import json
import sys, os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import thaler_study_b200 as T

v = int(sys.argv[1]) if len(sys.argv) > 1 else 28
T.options_from_env()  # e.g. SCB_PAIR_FIRST_ALONE=24: the threshold from which the first pair pass is a launch of its own
steps = 10
F = T.Field(1572869)
g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 0xB200 + k) for k in range(3)])
ref = None
for mode in ((0, 1, 2, 3, 4, 0, 2, 3) if len(sys.argv) < 3 else tuple(int(x) for x in sys.argv[2].split(','))):
    T.set_option("pair_w21", mode)
    for _ in range(3):
        tr = T.generate_transcript(T.Prover(g))
    T.lib.scb_resident_stats_reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        tr = T.generate_transcript(T.Prover(g))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    n, gms, wg, wp = C.c_uint64(), C.c_double(), C.c_uint64(), C.c_uint64()
    T.lib.scb_grid_pass_stats(C.byref(n), C.byref(gms), C.byref(wg), C.byref(wp))
    pn, pms = C.c_uint64(), C.c_double()
    T.lib.scb_pair_pass_stats(C.byref(pn), C.byref(pms))
    blob = b"".join(tr)
    ref = ref or blob
    print(json.dumps({"pair_w21": mode, "pair_first_alone": T.get_option("pair_first_alone"), "vars": v, "ms_per_proof": round(ms, 4), "Gelem_s": round((1 << v) / ms / 1e6, 2),
                      "grid_pass_ms": round(gms.value / max(n.value, 1), 4), "first_pair_pass_ms": round(pms.value / max(pn.value, 1), 4),
                      "w21_grid_launches": wg.value, "w21_pair_launches": wp.value, "same_transcript": blob == ref,
                      "verified": bool(T.verify_transcript(tr, T.Verifier(v, g)))}), flush=True)
