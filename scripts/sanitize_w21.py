"""compute-sanitizer workload for the 21-bit-triple kernels (k_grid_sp_pf_w21, k_pair_pass_sp_w21 in its four variants) on small
tables: every setting must print the transcript of the 8-byte path (checked here), with the resident kernels off (the sanitizer
serialises kernel and host) so that every pass is an ordinary launch."""
import os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thaler_study_b200 as T

T.set_option("pair_resident", 0)
F = T.Field(1572869)
n = 0
for v in (12, 13, 15, 16):
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 300 + k) for k in range(3)])
    want = None
    for mode in (0, 1, 2, 3, 4):
        T.set_option("pair_w21", mode)
        tr = b"".join(T.generate_transcript(T.Prover(g)))
        want = want or tr
        assert tr == want, (v, mode)
        n += 1
import ctypes as C

a, b = C.c_uint64(), C.c_uint64()
T.lib.scb_grid_pass_stats(None, None, C.byref(a), C.byref(b))
assert a.value == 16 and b.value == 16, (a.value, b.value)
print("w21 sanitize workload ok:", n, "proofs,", a.value, "triple grid passes,", b.value, "triple pair passes")
