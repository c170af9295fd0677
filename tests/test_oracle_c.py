"""Cross-check the plain-C oracle (oracle/oracle.c) against the Python big-int
oracle (oracle/pyoracle.py), which is itself pinned to the reference's KATs.
CPU only."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import pyoracle as O
from oracle.coracle import CField, _ptr, lib

FIELDS = [O.FP5, O.FP389, O.FP1572869, O.Field((1 << 61) - 1), O.Field(0xFFFFFFFF00000001), O.BLS12_381_FR]


@pytest.mark.parametrize("F", FIELDS, ids=lambda F: f"p{F.bits}")
def test_montgomery_arithmetic(F):
    cf = CField(F.p)
    rnd = random.Random(F.p & 0xFFFF)
    vals = [0, 1, F.p - 1, F.p // 2] + [rnd.randrange(F.p) for _ in range(60)]
    m = cf.to_mont(vals)
    assert cf.from_mont(m) == vals
    assert cf.unpack(m) == [F.to_mont(v) for v in vals]  # ark's in-memory representation
    ops = (
        ("orc_add", lambda x, y: (x + y) % F.p),
        ("orc_sub", lambda x, y: (x - y) % F.p),
        ("orc_mul", lambda x, y: x * y % F.p),
    )
    for i in range(len(vals) - 1):
        a, b = m[i : i + 1].copy(), m[i + 1 : i + 2].copy()
        for name, op in ops:
            out = np.empty_like(a)
            getattr(lib(), name)(C.byref(cf.f), _ptr(out), _ptr(a), _ptr(b))
            assert cf.from_mont(out) == [op(vals[i], vals[i + 1])], name


@pytest.mark.parametrize("F", FIELDS, ids=lambda F: f"p{F.bits}")
def test_fold_sum_round_evals_prove(F):
    cf = CField(F.p)
    rnd = random.Random(101)
    for K in (1, 2, 3):
        v = 5
        vals = [[rnd.randrange(F.p) for _ in range(1 << v)] for _ in range(K)]
        tabs = [cf.to_mont(t) for t in vals]
        g = O.ProductMLE(F, [O.DenseMLE(F, v, t) for t in vals])
        assert cf.from_mont(cf.product_sum(tabs)) == [O.Prover(g).c_1()]
        assert cf.from_mont(cf.product_sum(tabs, threads=3)) == [O.Prover(g).c_1()]
        assert cf.from_mont(cf.product_round_evals(tabs, K + 1)) == g.round_evals()
        r = rnd.randrange(F.p)
        folded = cf.fix_variable(tabs[0], cf.to_mont([r]))
        assert cf.from_mont(folded) == g.tables[0].fix_variables([r]).evals
        assert cf.from_mont(cf.fix_variable(tabs[0], cf.to_mont([r]), threads=2)) == cf.from_mont(folded)
        ch = [rnd.randrange(F.p) for _ in range(v - 1)]
        for threads in (1, 4):
            c1, ev = cf.product_prove(tabs, cf.to_mont(ch), K + 1, threads=threads)
            assert cf.from_mont(c1) == [O.Prover(g).c_1()]
            gg = g
            for j in range(v):
                if j:
                    gg = gg.fix_variables([ch[j - 1]])
                assert cf.from_mont(ev[j]) == gg.round_evals()


def test_matmul_k2_formula_is_the_reference_one():
    # matrix-multiplication/src/lib.rs:110-122 literal vs the generalised pass
    F = O.FP1572869
    cf = CField(F.p)
    rnd = random.Random(5)
    a = [rnd.randrange(F.p) for _ in range(64)]
    b = [rnd.randrange(F.p) for _ in range(64)]
    g = O.MatMulG(F, O.DenseMLE(F, 6, a), O.DenseMLE(F, 6, b))
    assert cf.from_mont(cf.product_round_evals([cf.to_mont(a), cf.to_mont(b)], 3)) == g.round_evals()


@pytest.mark.parametrize("F", [O.FP5, O.FP1572869, O.BLS12_381_FR], ids=lambda F: f"p{F.bits}")
def test_mle_eval(F):
    cf = CField(F.p)
    rnd = random.Random(7)
    for v in (1, 2, 6):
        ev = [rnd.randrange(F.p) for _ in range(1 << v)]
        r = [rnd.randrange(F.p) for _ in range(v)]
        want = O.vsbw_multilinear_from_evaluations(F, ev, r)
        assert cf.from_mont(cf.mle_vsbw(cf.to_mont(ev), cf.to_mont(r))) == [want]
        assert cf.from_mont(cf.mle_cti(cf.to_mont(ev), cf.to_mont(r))) == [want]
        assert cf.from_mont(cf.mle_evaluate_le(cf.to_mont(ev), cf.to_mont(r))) == [O.DenseMLE(F, v, ev).evaluate(r)]


def test_triangle_and_gkrw_against_python():
    F = O.FP1572869
    cf = CField(F.p)
    rnd = random.Random(9)
    n = 3
    m = [[0] * 8 for _ in range(8)]
    for i in range(8):
        for j in range(i + 1, 8):
            m[i][j] = m[j][i] = rnd.randrange(2)
    g = O.TriangleG.new_adj_matrix(F, 2 * n, [bool(x) for x in sum(m, [])])
    pts = O.domain4_elements(F) + [0, 1, 2]
    for j in range(3 * n):
        f1, f2, f3 = (cf.to_mont(t.evals) for t in (g.f_a_1, g.f_a_2, g.f_a_3))
        xn, yn, zn = g.x_vars_num(), g.y_vars_num(), g.z_vars_num()
        assert cf.from_mont(cf.triangle_sum(f1, f2, f3, xn, yn, zn)) == [sum(g.to_evaluations()) % F.p]
        for e in pts:
            want = sum(g.fix_variables([e]).to_evaluations()) % F.p
            assert cf.from_mont(cf.triangle_round_eval_at(f1, f2, f3, xn, yn, zn, cf.to_mont([e]))) == [want]
        g = g.fix_variables([rnd.randrange(F.p)])
    # GKR W with random dense tables
    k = 2
    w = O.GkrW(
        F,
        O.DenseMLE(F, 2 * k, [rnd.randrange(F.p) for _ in range(1 << (2 * k))]),
        O.DenseMLE(F, 2 * k, [rnd.randrange(F.p) for _ in range(1 << (2 * k))]),
        O.DenseMLE(F, k, [rnd.randrange(F.p) for _ in range(1 << k)]),
        O.DenseMLE(F, k, [rnd.randrange(F.p) for _ in range(1 << k)]),
    )
    for j in range(2 * k):
        a, mu, wb, wc = (cf.to_mont(t.evals) for t in (w.add_i, w.mul_i, w.w_b, w.w_c))
        bn, cn = w.w_b.num_vars, w.w_c.num_vars
        assert cf.from_mont(cf.gkrw_sum(a, mu, wb, wc, bn, cn)) == [sum(w.to_evaluations()) % F.p]
        for e in pts:
            want = sum(w.fix_variables([e]).to_evaluations()) % F.p
            assert cf.from_mont(cf.gkrw_round_eval_at(a, mu, wb, wc, bn, cn, cf.to_mont([e]))) == [want]
        w = w.fix_variables([rnd.randrange(F.p)])


def test_synth_stream_matches_python():
    for F in (O.FP1572869, O.BLS12_381_FR, O.Field((1 << 61) - 1)):
        cf = CField(F.p)
        got = cf.unpack(cf.synth(0xB200, 5, 40))
        assert got == [O.synth_element(F, 0xB200, 5 + i) for i in range(40)]
        assert all(x < F.p for x in got)
