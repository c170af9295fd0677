"""The drop-in through the TRAIT ONLY: exactly the calls an unmodified sum_check_protocol::Prover<F, GpuPoly> and
fiat_shamir::generate_transcript issue (sum-check-protocol/src/lib.rs:88-112, fiat-shamir/src/lib.rs:75-98) --
Prover::new = to_evaluations() + host sum, per round fix_variables then to_univariate -- must give the same bytes as
the library's fast path (scb_prover_new + scb_fs_generate_transcript, what rust/sumcheck-b200's GpuProver binds) and
as the Python oracle.  bench.py times the same sequence at 2^28 (`e2e.trait_only`)."""
import random

import pytest

from oracle import pyoracle as O

import thaler_study_b200 as T

pytestmark = pytest.mark.gpu
FIELDS = [O.FP5, O.FP389, O.FP1572869, O.Field(0xFFFFFFFF00000001), O.BLS12_381_FR]


def trait_only_transcript(g):
    F = g.F
    c_1 = sum(g.to_evaluations()) % F.p                       # Prover::new :89
    msgs, sofar = [], b""
    for j in range(g.num_vars()):
        if j > 0:
            g = g.fix_variables([F.hash_to_field(sofar)])     # self.g = self.g.fix_variables(&[r_prev]) :108
        m = (c_1.to_bytes(F.ser_bytes, "little") if j == 0 else b"") + g.to_univariate().serialize_uncompressed()  # :111
        msgs.append(m)
        sofar += m
    return c_1, msgs


@pytest.mark.parametrize("OF", FIELDS, ids=lambda F: f"p{F.bits}")
def test_trait_only_sequence_equals_fast_path_and_oracle(OF):
    F = T.Field(OF.p)
    rnd = random.Random(OF.bits)
    for K, v in ((1, 5), (2, 9), (3, 12), (3, 17)):
        if K >= OF.p:
            continue
        if v <= 12:
            vals = [[rnd.randrange(OF.p) for _ in range(1 << v)] for _ in range(K)]
            g = T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, v, t) for t in vals])
        else:
            vals = None
            g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 300 + k) for k in range(K)])
        c_1, slow = trait_only_transcript(g)
        prover = T.Prover(g)
        assert prover.c_1() == c_1
        fast = T.generate_transcript(prover)
        assert slow == fast, (K, v)
        assert T.verify_transcript(slow, T.Verifier(v, g))
        if vals is not None:
            og = O.ProductMLE(OF, [O.DenseMLE(OF, v, t) for t in vals])
            assert slow == O.generate_transcript(OF, O.Prover(og))


def test_trait_only_matmul_g():
    OF, F = O.FP389, T.Field(389)
    rnd = random.Random(2)
    v = 8
    a, b = [rnd.randrange(389) for _ in range(1 << v)], [rnd.randrange(389) for _ in range(1 << v)]
    g = T.MatMulG.from_tables(T.DenseMultilinearExtension.from_evaluations_vec(F, v, a), T.DenseMultilinearExtension.from_evaluations_vec(F, v, b))
    _, slow = trait_only_transcript(g)
    assert slow == T.generate_transcript(T.Prover(g))
    assert slow == O.generate_transcript(OF, O.Prover(O.MatMulG(OF, O.DenseMLE(OF, v, a), O.DenseMLE(OF, v, b))))
