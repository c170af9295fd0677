"""Pin the Python oracle against every known-answer vector the reference's own
tests hold for the hot path, and replay
the reference's self-consistency tests (table S) on it.  CPU only."""
import random

import pytest

from oracle import pyoracle as O
from oracle.pyoracle import FP5, FP389, FP1572869


class PyRng:
    """RngF stand-in (sum-check-protocol/src/lib.rs:13-21)."""

    def __init__(self, F, seed=0):
        self.F, self.r = F, random.Random(seed)

    def draw(self):
        return self.r.randrange(self.F.p)


def run_protocol(F, g, rng):
    """The loop every reference test uses, e.g. sum-check-protocol/src/lib.rs:440-458."""
    prover = O.Prover(g)
    n = g.num_vars()
    verifier = O.Verifier(n, g)
    verifier.set_c_1(prover.c_1())
    r_j = 1
    final = None
    for j in range(n):
        g_j = prover.round(r_j, j)
        kind, val = verifier.round(g_j, rng)
        if kind == "JthRound":
            r_j = val
        else:
            final = val
    return final


# ---- multilinear-extensions/src/lib.rs:76-120 -------------------------------
def test_mle_example_from_book():
    evals = [1, 2, 1, 4]
    expected = [
        [1, 2, 3, 4, 0],
        [1, 4, 2, 0, 3],
        [1, 1, 1, 1, 1],
        [1, 3, 0, 2, 4],
        [1, 0, 4, 3, 2],
    ]
    for fn in (O.cti_multilinear_from_evaluations, O.vsbw_multilinear_from_evaluations):
        for i in range(5):
            assert [fn(FP5, evals, [i, j]) for j in range(5)] == expected[i]


def test_mle_big_endian_vs_ark_little_endian():
    rnd = random.Random(1)
    F = FP1572869
    for v in range(1, 7):
        evals = [rnd.randrange(F.p) for _ in range(1 << v)]
        r = [rnd.randrange(F.p) for _ in range(v)]
        a = O.vsbw_multilinear_from_evaluations(F, evals, r)
        b = O.cti_multilinear_from_evaluations(F, evals, r)
        c = O.DenseMLE(F, v, evals).evaluate(list(reversed(r)))
        assert a == b == c


# ---- sum-check-protocol/src/lib.rs:383-416 ----------------------------------
def test_sumcheck_basic_test_fix_variables():
    poly = O.SparseMVPoly(FP5, 2, [(2, [(0, 1), (1, 1)]), (3, [(0, 2), (1, 2)])])
    res = poly.fix_variables([2])
    expected = O.SparseMVPoly(FP5, 1, [(4, [(0, 1)]), (2, [(0, 2)])])
    assert res.nv == expected.nv and sorted(res.terms) == sorted(expected.terms)


# ---- sum-check-protocol/src/lib.rs:418-459 ----------------------------------
def test_sumcheck_test_from_book():
    g = O.SparseMVPoly(FP5, 3, [(2, [(0, 3)]), (1, [(0, 1), (2, 1)]), (1, [(1, 1), (2, 1)])])
    assert O.Prover(g).c_1() == 12 % 5
    assert run_protocol(FP5, g, PyRng(FP5, 3)) is True


def rand_mv_poly(F, l, d, rnd):
    terms = [(rnd.randrange(F.p), [])]
    for _ in range(rnd.randrange(1, 40)):
        term = [(i, rnd.randrange(1, d + 1)) for i in range(l) if rnd.random() < 0.5]
        terms.append((rnd.randrange(F.p), term))
    return O.SparseMVPoly(F, l, terms)


# ---- sum-check-protocol/src/lib.rs:494-521 ----------------------------------
def test_sumcheck_protocol_test():
    rnd = random.Random(7)
    for n in range(2, 8):
        g = rand_mv_poly(FP5, n, 3, rnd)
        assert run_protocol(FP5, g, PyRng(FP5, n)) is True


# ---- matrix-multiplication/src/lib.rs:245-303, 315-374 ----------------------
def u32_to_boolean_vec(v, bits):
    return [(v >> i) & 1 for i in range(bits)]


def matmul(F, a, b, n):
    return [[sum(a[i][k] * b[k][j] for k in range(n)) % F.p for j in range(n)] for i in range(n)]


def test_matmul_example_from_book():
    a = [[0, 1], [2, 0]]
    b = [[1, 0], [0, 4]]
    assert matmul(FP5, a, b, 2) == [[0, 4], [2, 0]]  # :202-243
    for i in range(2):
        for j in range(2):
            point = u32_to_boolean_vec(i, 1) + u32_to_boolean_vec(j, 1)
            g = O.MatMulG.new(FP5, 1, sum(a, []), sum(b, []), point)
            assert O.Prover(g).c_1() == matmul(FP5, a, b, 2)[i][j]
            # n = 1: the reference's Verifier takes the "first round" branch
            # (sum-check-protocol/src/lib.rs:284-297) and never reaches FinalRound.
            assert run_protocol(FP5, g, PyRng(FP5, i * 2 + j)) is None


@pytest.mark.parametrize("F", [FP5, FP1572869])
def test_matmul_randomized_test(F):
    rnd = random.Random(11)
    for p in range(2, 5):
        n = 1 << p
        a = [[rnd.randrange(F.p) for _ in range(n)] for _ in range(n)]
        b = [[rnd.randrange(F.p) for _ in range(n)] for _ in range(n)]
        c = matmul(F, a, b, n)
        for i in range(n):
            for j in range(n):
                point = u32_to_boolean_vec(i, p) + u32_to_boolean_vec(j, p)
                g = O.MatMulG.new(F, p, sum(a, []), sum(b, []), point)
                prover = O.Prover(g)
                assert prover.c_1() == c[i][j]  # :339-340
                resu = sum(g.evaluate(u32_to_boolean_vec(x, p)) for x in range(n)) % F.p
                assert prover.c_1() == resu  # :342-352
                if (i + j) % 5 == 0:
                    assert run_protocol(F, g, PyRng(F, i * n + j)) is True


# ---- triangle-counting/src/lib.rs:224-266, 268-318 --------------------------
def adj_matrix(n, rnd):
    m = [[False] * n for _ in range(n)]
    for i in range(n):
        for j in range(i + 1, n):
            m[i][j] = m[j][i] = rnd.random() < 0.5
    return m


def triangle_count(m):
    n = len(m)
    return sum(1 for x in range(n) for y in range(n) for z in range(n) if m[x][y] and m[y][z] and m[x][z]) // 6


def test_triangle_simple_matrix():
    adj = [
        [False, True, True, False],
        [True, False, True, False],
        [True, True, False, False],
        [False, False, False, False],
    ]
    g = O.TriangleG.new_adj_matrix(FP389, len(adj), sum(adj, []))
    assert O.Prover(g).c_1() == 6
    assert run_protocol(FP389, g, PyRng(FP389, 5)) is True


def test_triangle_randomized_test():
    rnd = random.Random(13)
    F = FP1572869
    for i in range(1, 5):
        n = 1 << i
        m = adj_matrix(n, rnd)
        g = O.TriangleG.new_adj_matrix(F, 2 * i, sum(m, []))
        assert O.Prover(g).c_1() == 6 * triangle_count(m)  # :294-300
        assert run_protocol(F, g, PyRng(F, i)) is True


def test_triangle_round_degree_is_two():
    # every variable occurs in only two of the three factors
    rnd = random.Random(17)
    F = FP1572869
    m = adj_matrix(8, rnd)
    g = O.TriangleG.new_adj_matrix(F, 6, sum(m, []))
    prover = O.Prover(g)
    r = 1
    for j in range(g.num_vars()):
        poly = prover.round(r, j)
        assert all(d <= 2 for d, _ in poly.coeffs)
        r = rnd.randrange(F.p)


# ---- gkr-protocol ------------------------------------------------------------
def test_gkr_restrict_poly():  # gkr-protocol/src/lib.rs:506-548
    poly = O.restrict_poly(FP389, [2, 4], [3, 2], O.DenseMLE(FP389, 2, [0, 0, 2, 5]))
    assert poly.to_dense() == [32, 385, 383]


def test_gkr_circuit_from_book():  # gkr-protocol/src/circuit.rs:258-284
    c = O.circuit_from_book()
    assert c.evaluate(O.Field(1 << 61), [3, 2, 3, 1]) == [[36, 6], [9, 4, 6, 1], [3, 2, 3, 1]]
    for a in range(4):
        for b in range(4):
            for cc in range(4):
                exp = ((a in (0, 1)) and a == b and a == cc) or (a == 2 and b == 1 and cc == 2) or (a == b == cc == 3)
                assert c.mul_i(1, a, b, cc) == exp


def run_gkr(F, circuit, inp, expected_outputs, rng):
    """gkr-protocol/src/lib.rs:574-623."""
    prover = O.GkrProver(F, circuit, inp)
    begin = prover.start_protocol()
    assert begin == ("Begin", expected_outputs)
    verifier = O.GkrVerifier(F, circuit)
    kind, r_i = verifier.receive_prover_msg(begin, rng)
    assert kind == "R"
    for i in range(len(circuit.layers)):
        msg = prover.start_round(i, r_i)
        num_vars = 2 * circuit.num_vars_at(i + 1)
        verifier.receive_prover_msg(msg, rng)
        for j in range(num_vars - 1):
            vmsg = verifier.receive_prover_msg(prover.round_msg(j), rng)
            prover.receive_verifier_msg(vmsg)
        prover.receive_verifier_msg(verifier.final_random_point(rng))
        kind, r_i = verifier.receive_prover_msg(prover.round_msg(num_vars - 1), rng)
        assert kind == "R"
    return verifier.check_input(inp)


def test_gkr_protocol_test_from_book():
    assert run_gkr(FP389, O.circuit_from_book(), [3, 2, 3, 1], [36, 6], PyRng(FP389, 19))


def test_gkr_three_layer_protocol_test():
    assert run_gkr(FP389, O.three_layer_circuit(), [0, 1] * 4, [2, 2], PyRng(FP389, 23))


def test_gkr_round_degree_is_two():
    F = FP389
    circuit = O.circuit_from_book()
    prover = O.GkrProver(F, circuit, [3, 2, 3, 1])
    prover.start_round(0, [7])
    rnd = random.Random(3)
    r = 1
    for j in range(4):
        poly = prover.prover.round(r, j)
        assert all(d <= 2 for d, _ in poly.coeffs)
        r = rnd.randrange(F.p)


# ---- fiat-shamir/src/lib.rs:219-236 -------------------------------------------
def test_fiat_shamir_it_works():
    rnd = random.Random(29)
    for n in range(2, 8):
        g = rand_mv_poly(FP5, n, 3, rnd)
        transcript = O.generate_transcript(FP5, O.Prover(g))
        assert len(transcript) == n
        assert O.verify_transcript(FP5, transcript, O.Verifier(n, g))


def test_fiat_shamir_on_product_and_matmul():
    rnd = random.Random(31)
    for F in (FP5, FP389, FP1572869, O.BLS12_381_FR):
        for v in (2, 5):
            tabs = [O.DenseMLE(F, v, [rnd.randrange(F.p) for _ in range(1 << v)]) for _ in range(3)]
            g = O.ProductMLE(F, tabs)
            assert O.verify_transcript(F, O.generate_transcript(F, O.Prover(g)), O.Verifier(v, g))
            g2 = O.MatMulG(F, tabs[0], tabs[1])
            assert O.verify_transcript(F, O.generate_transcript(F, O.Prover(g2)), O.Verifier(v, g2))


# ---- SparsePolynomial conventions (SURVEY section 7 hard part 2) ----------------
def test_interpolate_quadratic_explicit_zero_term():
    F = FP5
    # g(0) = 0, g(1) != 0 : poly_1 is zero => result = poly_2 (+ poly_3) keeps the explicit (0,0)
    p = O.interpolate_quadratic_poly(F, [(0, 0), (1, 3), (2, 0)])
    assert p.coeffs[0] == (0, 0)
    assert [p.evaluate(x) for x in range(3)] == [0, 3, 0]
    # all-zero message -> empty polynomial
    assert O.interpolate_quadratic_poly(F, [(0, 0), (1, 0), (2, 0)]).coeffs == []
    # generic: agrees with dense Lagrange
    rnd = random.Random(5)
    for F in (FP5, FP389, FP1572869):
        for _ in range(50):
            ys = [rnd.randrange(F.p) for _ in range(3)]
            q = O.interpolate_quadratic_poly(F, list(zip(range(3), ys)))
            assert [q.evaluate(x) for x in range(3)] == ys
            assert q.to_dense() == O.lagrange_to_coeffs(F, ys)


def test_product_mle_k2_matches_matmul_g_values():
    rnd = random.Random(37)
    F = FP1572869
    tabs = [O.DenseMLE(F, 6, [rnd.randrange(F.p) for _ in range(64)]) for _ in range(2)]
    a, b = O.ProductMLE(F, tabs), O.MatMulG(F, tabs[0], tabs[1])
    assert a.round_evals() == b.round_evals()
    assert a.to_univariate().to_dense() == b.to_univariate().to_dense()


def test_two_adicity_of_reference_fields():
    assert [F.two_adicity() for F in (FP5, FP389, FP1572869)] == [2, 2, 2]
    assert O.BLS12_381_FR.two_adicity() == 32


def test_hash_to_field_shape():
    # not pinned by the reference (no vector); pins the restated shape only
    assert (FP5.bits + 128 + 7) // 8 == 17
    r = O.hash_to_field(FP5, b"")
    assert 0 <= r < 5
    assert O.hash_to_field(O.BLS12_381_FR, b"abc") < O.BLS12_381_FR.p
