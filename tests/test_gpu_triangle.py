"""triangle_counting::G in matrix form (csrc/tri.cuh): M = f2 x f1 computed once per proof by the tiled field matmul,
then folded along x, against round 1's per-round n^3 kernels (option tri_tiled = 0) and the oracle."""
import random

import numpy as np
import pytest

from oracle import pyoracle as O

import thaler_study_b200 as T

pytestmark = pytest.mark.gpu
FIELDS = [O.FP389, O.FP1572869, O.Field(268435361), O.Field(0xFFFFFFFF00000001), O.BLS12_381_FR]


def random_graph(n, rnd):
    up = np.triu(rnd.integers(0, 2, size=(n, n), dtype=np.int64), 1)
    return up + up.T


@pytest.mark.parametrize("OF", FIELDS, ids=lambda F: f"p{F.bits}")
def test_matrix_form_equals_per_round_kernels(OF):
    F = T.Field(OF.p)
    rnd = np.random.default_rng(OF.bits)
    prnd = random.Random(OF.bits)
    for bits in (1, 2, 3, 5, 6, 7):
        n = 1 << bits
        adj = random_graph(n, rnd)
        flat = adj.reshape(-1).astype(bool).tolist()
        outs = []
        pt = [prnd.randrange(OF.p) for _ in range(3 * bits)]
        for tiled in (1, 0):
            T.set_option("tri_tiled", tiled)
            g = T.TriangleG.new_adj_matrix(F, 2 * bits, flat)
            prover = T.Prover(g)
            c_1 = prover.c_1()
            tr = T.generate_transcript(prover)
            assert T.verify_transcript(tr, T.Verifier(3 * bits, g))
            partial = []
            for k in (1, bits - 1, bits, bits + 1, 2 * bits):
                if 1 <= k < 3 * bits:
                    gk = g.fix_variables(pt[:k])
                    partial.append((gk.sum(), gk.round_evals(), gk.num_vars()))
            outs.append((c_1, tr, partial, g.evaluate(pt)))
        assert outs[0] == outs[1], bits
        tri6 = int(((adj @ adj) * adj).sum())
        assert outs[0][0] == tri6 % OF.p
    T.reset_options()


def test_matmul_kernel_with_rectangular_and_small_shapes():
    """fix_variables leaves f1 over (x', y) with fewer x bits than y bits: the matmul behind a freshly created handle is
    square, but sums and messages of partially folded handles go through the folded M -- compare with the oracle."""
    OF, F = O.FP1572869, T.Field(1572869)
    rnd = np.random.default_rng(5)
    prnd = random.Random(5)
    bits = 3
    adj = random_graph(1 << bits, rnd)
    flat = adj.reshape(-1).astype(bool).tolist()
    og = O.TriangleG.new_adj_matrix(OF, 2 * bits, flat)
    g = T.TriangleG.new_adj_matrix(F, 2 * bits, flat)
    assert T.generate_transcript(T.Prover(g)) == O.generate_transcript(OF, O.Prover(og))
    pt = [prnd.randrange(OF.p) for _ in range(3 * bits)]
    for k in range(1, 3 * bits):
        gk, ok = g.fix_variables(pt[:k]), og.fix_variables(pt[:k])
        assert gk.sum() == sum(ok.to_evaluations()) % OF.p
        assert gk.to_univariate().coeffs == ok.to_univariate().coeffs
