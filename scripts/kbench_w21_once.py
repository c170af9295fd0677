"""Two proofs of the headline workload (for ncu captures of the 21-bit-triple kernels)."""
import os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thaler_study_b200 as T

T.options_from_env()
F = T.Field(1572869)
g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, 28, 0xB200 + k) for k in range(3)])
for _ in range(2):
    tr = T.generate_transcript(T.Prover(g))
assert T.verify_transcript(tr, T.Verifier(28, g))
print("ok")
