"""GPU parity tests: the CUDA engine (through the C ABI) against the oracles on the same inputs.
Bit-exact everywhere: field elements are compared as canonical integers, transcripts as bytes."""
import random

import numpy as np
import pytest

from oracle import pyoracle as O
from oracle.coracle import CField

import thaler_study_b200 as T

pytestmark = pytest.mark.gpu

FIELDS = [O.FP5, O.FP389, O.FP1572869, O.Field((1 << 61) - 1), O.Field(0xFFFFFFFF00000001), O.BLS12_381_FR]
fid = lambda F: f"p{F.bits}"
# multiplicative generators of the MontConfigs (2 in every reference field; 7 for ark_ed_on_bls12_381::Fq)
GENERATOR = {O.BLS12_381_FR.p: 7}


class PyRng:
    def __init__(self, F, seed=0):
        self.F, self.r = F, random.Random(seed)

    def draw(self):
        return self.r.randrange(self.F.p)


def rand_table(OF, v, rnd):
    return [rnd.randrange(OF.p) for _ in range(1 << v)]


def run_protocol(g, rng, c_1_expected=None):
    """The loop of every reference test (e.g. matrix-multiplication/src/lib.rs:354-370) on device types."""
    prover = T.Prover(g)
    c_1 = prover.c_1()
    if c_1_expected is not None:
        assert c_1 == c_1_expected
    n = g.num_vars()
    verifier = T.Verifier(n, g)
    verifier.set_c_1(c_1)
    r_j, final = 1, None
    for j in range(n):
        g_j = prover.round(r_j, j)
        kind, val = verifier.round(g_j, rng)
        if kind == "JthRound":
            r_j = val
        else:
            final = val
    return final


# ----------------------------------------------------------------------------- a4 / a2 / a5 kernels
@pytest.mark.parametrize("OF", FIELDS, ids=fid)
def test_fold_sum_round_evals_vs_c_oracle(OF):
    F, cf = T.Field(OF.p), CField(OF.p)
    rnd = random.Random(OF.p & 0xFFFF)
    for v in (1, 2, 3, 5, 9, 13):
        seeds = [rnd.randrange(1 << 20) for _ in range(4)]
        tabs_c = [cf.synth(s, 0, 1 << v) for s in seeds]
        tabs_g = [T.DenseMultilinearExtension.synthetic(F, v, s) for s in seeds]
        for tc, tg in zip(tabs_c, tabs_g):
            assert np.array_equal(tg.to_evaluations_mont(), tc)  # same synthetic stream on device and host
        r = rnd.randrange(OF.p)
        folded = tabs_g[0].fix_variables([r])
        assert np.array_equal(folded.to_evaluations_mont(), cf.fix_variable(tabs_c[0], cf.to_mont([r])))
        for K in (1, 2, 3, 4):
            if K >= OF.p:
                continue
            g = T.ProductMLE.new(tabs_g[:K])
            assert g.num_vars() == v and g.n_points == K + 1
            assert g.sum() == cf.from_mont(cf.product_sum(tabs_c[:K]))[0]
            assert g.round_evals() == cf.from_mont(cf.product_round_evals(tabs_c[:K], K + 1))
            if v >= 2:
                g2, ev = g.fix_and_round_evals(r)
                f_c = [cf.fix_variable(t, cf.to_mont([r])) for t in tabs_c[:K]]
                assert ev == cf.from_mont(cf.product_round_evals(f_c, K + 1))
                for k in range(K):
                    assert np.array_equal(g2.table(k).to_evaluations_mont(), f_c[k])
                assert g2.num_vars() == v - 1
                # unfused path gives the same polynomial object
                g3 = g.fix_variables([r])
                assert g3.round_evals() == ev


@pytest.mark.parametrize("OF", [O.FP1572869, O.Field(0xFFFFFFFF00000001), O.BLS12_381_FR], ids=fid)
def test_full_prover_vs_c_oracle_v16(OF):
    F, cf = T.Field(OF.p), CField(OF.p)
    rnd = random.Random(41)
    v, K = 16, 3
    seeds = [7, 8, 9]
    tabs_c = [cf.synth(s, 0, 1 << v) for s in seeds]
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, s) for s in seeds])
    ch = [rnd.randrange(OF.p) for _ in range(v - 1)]
    c1, ev = cf.product_prove(tabs_c, cf.to_mont(ch), K + 1, threads=4)
    assert g.sum() == cf.from_mont(c1)[0]
    assert g.round_evals() == cf.from_mont(ev[0])
    for j in range(1, v):
        g, got = g.fix_and_round_evals(ch[j - 1])
        assert got == cf.from_mont(ev[j]), j


# ----------------------------------------------------------------------------- transcripts, byte for byte
def oracle_and_device_polys(OF, F, kind, v, rnd):
    if kind == "product1" or kind == "product2" or kind == "product3" or kind == "product4":
        K = int(kind[-1])
        vals = [rand_table(OF, v, rnd) for _ in range(K)]
        og = O.ProductMLE(OF, [O.DenseMLE(OF, v, t) for t in vals])
        dg = T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, v, t) for t in vals])
    elif kind == "matmul":
        a, b = rand_table(OF, v, rnd), rand_table(OF, v, rnd)
        og = O.MatMulG(OF, O.DenseMLE(OF, v, a), O.DenseMLE(OF, v, b))
        dg = T.MatMulG.from_tables(
            T.DenseMultilinearExtension.from_evaluations_vec(F, v, a), T.DenseMultilinearExtension.from_evaluations_vec(F, v, b)
        )
    else:
        raise ValueError(kind)
    return og, dg


@pytest.mark.parametrize("OF", FIELDS, ids=fid)
@pytest.mark.parametrize("kind", ["product1", "product2", "product3", "product4", "matmul"])
def test_fiat_shamir_transcript_bytes(OF, kind):
    if kind.startswith("product") and int(kind[-1]) >= OF.p:
        pytest.skip("degree >= characteristic")
    F = T.Field(OF.p)
    rnd = random.Random((OF.p * 31 + len(kind) * 7 + ord(kind[-1])) & 0xFFFFFF)
    for v in (1, 2, 3, 6, 9):
        og, dg = oracle_and_device_polys(OF, F, kind, v, rnd)
        want = O.generate_transcript(OF, O.Prover(og))
        got = T.generate_transcript(T.Prover(dg))
        assert got == want, (kind, v)
        assert T.verify_transcript(got, T.Verifier(v, dg))
        assert O.verify_transcript(OF, got, O.Verifier(v, og))
        if v >= 2 and OF.bits > 20:
            # (the reference's last-round branch checks only g_v(r_v) == g(r), :298-310, so over a tiny field a
            # tampered coefficient is accepted whenever r_v^deg vanishes; only test where that cannot happen)
            bad = list(got)
            last = bytearray(bad[-1])
            last[-1] ^= 1
            bad[-1] = bytes(last)
            try:
                ok = T.verify_transcript(bad, T.Verifier(v, dg))
            except T.ScbError:
                ok = False  # non-canonical coefficient -> codec error, or claim mismatch
            assert not ok


def test_zero_tables_give_empty_polynomials():
    # common over F5: all-zero messages must serialise as empty coefficient lists
    OF = O.FP5
    F = T.Field(5)
    v = 4
    zero = [0] * (1 << v)
    a = rand_table(OF, v, random.Random(1))
    for mk_o, mk_d in (
        (lambda x, y: O.MatMulG(OF, O.DenseMLE(OF, v, x), O.DenseMLE(OF, v, y)),
         lambda x, y: T.MatMulG.from_tables(T.DenseMultilinearExtension.from_evaluations_vec(F, v, x), T.DenseMultilinearExtension.from_evaluations_vec(F, v, y))),
        (lambda x, y: O.ProductMLE(OF, [O.DenseMLE(OF, v, x), O.DenseMLE(OF, v, y)]),
         lambda x, y: T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, v, x), T.DenseMultilinearExtension.from_evaluations_vec(F, v, y)])),
    ):
        og, dg = mk_o(a, zero), mk_d(a, zero)
        assert T.generate_transcript(T.Prover(dg)) == O.generate_transcript(OF, O.Prover(og))


# ----------------------------------------------------------------------------- reference test replays
def u32_to_boolean_vec(v, bits):
    return [(v >> i) & 1 for i in range(bits)]


def matmul(p, a, b, n):
    return [[sum(a[i][k] * b[k][j] for k in range(n)) % p for j in range(n)] for i in range(n)]


def test_matmul_example_from_book():
    # matrix-multiplication/src/lib.rs:245-303
    F = T.Field(5)
    a, b = [[0, 1], [2, 0]], [[1, 0], [0, 4]]
    c = matmul(5, a, b, 2)
    assert c == [[0, 4], [2, 0]]
    for i in range(2):
        for j in range(2):
            point = u32_to_boolean_vec(i, 1) + u32_to_boolean_vec(j, 1)
            g = T.MatMulG.new(F, 1, sum(a, []), sum(b, []), point)
            run_protocol(g, PyRng(O.FP5, i * 2 + j), c_1_expected=c[i][j])


@pytest.mark.parametrize("OF", [O.FP5, O.FP1572869, O.BLS12_381_FR], ids=fid)
def test_matmul_randomized_test(OF):
    # matrix-multiplication/src/lib.rs:315-374
    F = T.Field(OF.p)
    rnd = random.Random(11)
    for p in range(2, 5):
        n = 1 << p
        a = [[rnd.randrange(OF.p) for _ in range(n)] for _ in range(n)]
        b = [[rnd.randrange(OF.p) for _ in range(n)] for _ in range(n)]
        c = matmul(OF.p, a, b, n)
        for i in range(0, n, max(1, n // 4)):
            for j in range(0, n, max(1, n // 4)):
                point = u32_to_boolean_vec(i, p) + u32_to_boolean_vec(j, p)
                g = T.MatMulG.new(F, p, sum(a, []), sum(b, []), point)
                og = O.MatMulG.new(OF, p, sum(a, []), sum(b, []), point)
                assert g.table(0).to_evaluations() == og.f_a.evals  # relabel + fix layout (:81-86)
                assert g.table(1).to_evaluations() == og.f_b.evals
                resu = sum(g.evaluate(u32_to_boolean_vec(x, p)) for x in range(n)) % OF.p
                assert resu == c[i][j]  # :342-352
                assert run_protocol(g, PyRng(OF, i * n + j), c_1_expected=c[i][j]) is True
        # random (non-Boolean) point: parity with the oracle's G::new and transcript
        point = [rnd.randrange(OF.p) for _ in range(2 * p)]
        g = T.MatMulG.new(F, p, sum(a, []), sum(b, []), point)
        og = O.MatMulG.new(OF, p, sum(a, []), sum(b, []), point)
        assert T.generate_transcript(T.Prover(g)) == O.generate_transcript(OF, O.Prover(og))


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
    # triangle-counting/src/lib.rs:224-266
    F = T.Field(389)
    adj = [[False, True, True, False], [True, False, True, False], [True, True, False, False], [False, False, False, False]]
    g = T.TriangleG.new_adj_matrix(F, len(adj), sum(adj, []))
    assert g.num_vars() == 6
    assert run_protocol(g, PyRng(O.FP389, 5), c_1_expected=6) is True


@pytest.mark.parametrize("OF", [O.FP1572869, O.Field((1 << 61) - 1), O.BLS12_381_FR], ids=fid)
def test_triangle_randomized_test(OF):
    # triangle-counting/src/lib.rs:268-318 (n = 2..32) + round-by-round parity with the oracle
    F = T.Field(OF.p)
    rnd = random.Random(13)
    for i in range(1, 6):
        n = 1 << i
        m = adj_matrix(n, rnd)
        flat = sum(m, [])
        g = T.TriangleG.new_adj_matrix(F, 2 * i, flat)
        assert run_protocol(g, PyRng(OF, i), c_1_expected=6 * triangle_count(m) % OF.p) is True
        if i <= 3:
            og = O.TriangleG.new_adj_matrix(OF, 2 * i, flat, generator=GENERATOR.get(OF.p, 2))
            if OF.two_adicity() >= 2:  # the reference's size-4 FFT domain exists (else its to_univariate panics)
                assert T.generate_transcript(T.Prover(g)) == O.generate_transcript(OF, O.Prover(og))
            assert g.to_evaluations() == og.to_evaluations()
            pt = [rnd.randrange(OF.p) for _ in range(3 * i)]
            assert g.evaluate(pt) == og.evaluate(pt)
            for k in range(1, 3 * i):
                gk, ok = g.fix_variables(pt[:k]), og.fix_variables(pt[:k])
                assert gk.num_vars() == ok.num_vars()
                assert gk.sum() == sum(ok.to_evaluations()) % OF.p
                assert gk.to_evaluations() == ok.to_evaluations()


@pytest.mark.parametrize("OF", [O.FP389, O.FP1572869, O.BLS12_381_FR], ids=fid)
def test_gkr_w_vs_oracle(OF):
    F = T.Field(OF.p)
    rnd = random.Random(17)
    for k in (1, 2, 3):
        tabs = [rand_table(OF, 2 * k, rnd), rand_table(OF, 2 * k, rnd), rand_table(OF, k, rnd), rand_table(OF, k, rnd)]
        ow = O.GkrW(OF, O.DenseMLE(OF, 2 * k, tabs[0]), O.DenseMLE(OF, 2 * k, tabs[1]), O.DenseMLE(OF, k, tabs[2]), O.DenseMLE(OF, k, tabs[3]), generator=GENERATOR.get(OF.p, 2))
        dw = T.GkrW.new(*[T.DenseMultilinearExtension.from_evaluations_vec(F, nv, t) for nv, t in zip((2 * k, 2 * k, k, k), tabs)])
        assert dw.num_vars() == 2 * k
        assert dw.to_evaluations() == ow.to_evaluations()
        assert dw.sum() == sum(ow.to_evaluations()) % OF.p
        pt = [rnd.randrange(OF.p) for _ in range(2 * k)]
        assert dw.evaluate(pt) == ow.evaluate(pt)
        assert T.generate_transcript(T.Prover(dw)) == O.generate_transcript(OF, O.Prover(ow))
        assert run_protocol(dw, PyRng(OF, k)) is (True if 2 * k > 1 else None)


def test_gkr_w_from_book_circuit():
    # the W polynomials the reference's GKR prover builds for circuit_from_book (gkr-protocol/src/lib.rs:373-436)
    OF, F = O.FP389, T.Field(389)
    rnd = random.Random(19)
    circuit = O.circuit_from_book()
    layers = circuit.evaluate(OF, [3, 2, 3, 1])
    assert layers[0] == [36, 6]
    for i in range(2):
        r_i = [rnd.randrange(OF.p) for _ in range(circuit.num_vars_at(i))]
        add_i, mul_i = circuit.wiring_tables(OF, i)
        d_add = T.DenseMultilinearExtension.from_evaluations_vec(F, add_i.num_vars, add_i.evals).fix_variables(r_i)
        d_mul = T.DenseMultilinearExtension.from_evaluations_vec(F, mul_i.num_vars, mul_i.evals).fix_variables(r_i)
        kn = circuit.num_vars_at(i + 1)
        d_w = T.DenseMultilinearExtension.from_evaluations_vec(F, kn, layers[i + 1])
        dw = T.GkrW.new(d_add, d_mul, d_w, d_w.clone())
        o_w = O.DenseMLE(OF, kn, layers[i + 1])
        ow = O.GkrW(OF, add_i.fix_variables(r_i), mul_i.fix_variables(r_i), o_w, o_w.clone())
        op, dp = O.Prover(ow), T.Prover(dw)
        assert dp.c_1() == op.c_1()
        r = 1
        for j in range(2 * kn):
            a, b = dp.round(r, j), op.round(r, j)
            assert a.coeffs == b.coeffs
            r = rnd.randrange(OF.p)


# ----------------------------------------------------------------------------- a8 / a9 / a10 MLE evaluation
def test_mle_example_from_book():
    # multilinear-extensions/src/lib.rs:76-120
    F = T.Field(5)
    expected = [[1, 2, 3, 4, 0], [1, 4, 2, 0, 3], [1, 1, 1, 1, 1], [1, 3, 0, 2, 4], [1, 0, 4, 3, 2]]
    for fn in (T.cti_multilinear_from_evaluations, T.vsbw_multilinear_from_evaluations):
        for i in range(5):
            assert [fn(F, [1, 2, 1, 4], [i, j]) for j in range(5)] == expected[i]


@pytest.mark.parametrize("OF", FIELDS, ids=fid)
def test_mle_eval_vs_oracle(OF):
    F, cf = T.Field(OF.p), CField(OF.p)
    rnd = random.Random(23)
    for v in (0, 1, 2, 3, 7, 11, 12, 13, 15):
        tab_c = cf.synth(v + 100, 0, 1 << v)
        m = T.DenseMultilinearExtension.synthetic(F, v, v + 100)
        r = [rnd.randrange(OF.p) for _ in range(v)]
        want_be = cf.from_mont(cf.mle_vsbw(tab_c, cf.to_mont(r) if v else np.zeros((0, cf.n), dtype=np.uint64)))[0] if v else cf.from_mont(tab_c)[0]
        assert m.evaluate_be(r) == want_be
        assert m.evaluate(list(reversed(r))) == want_be
        if v:
            assert T.vsbw_multilinear_from_evaluations(F, tab_c, r) == want_be
            assert m.evaluate(r) == cf.from_mont(cf.mle_evaluate_le(tab_c, cf.to_mont(r)))[0]
        if 0 < v <= 7:
            assert want_be == O.vsbw_multilinear_from_evaluations(OF, cf.from_mont(tab_c), r)


# ----------------------------------------------------------------------------- relabel
def test_relabel_vs_oracle():
    OF, F = O.FP1572869, T.Field(1572869)
    rnd = random.Random(29)
    t = rand_table(OF, 8, rnd)
    om, dm = O.DenseMLE(OF, 8, t), T.DenseMultilinearExtension.from_evaluations_vec(F, 8, t)
    for a, b, k in ((0, 4, 4), (0, 2, 2), (1, 5, 3), (4, 0, 4), (3, 3, 2), (0, 7, 1)):
        assert dm.relabel(a, b, k).to_evaluations() == om.relabel(a, b, k).evals
    with pytest.raises(T.ScbError):
        dm.relabel(0, 2, 4)


# ----------------------------------------------------------------------------- error behaviour
def test_error_conventions():
    F = T.Field(1572869)
    m = T.DenseMultilinearExtension.from_evaluations_vec(F, 3, list(range(8)))
    g = T.ProductMLE.new([m, m])
    assert g.evaluate([1, 2]) is None  # dimension mismatch -> None (sum-check-protocol/src/lib.rs:124-126)
    with pytest.raises(T.ScbError):
        m.fix_variables([1, 2, 3, 4])  # [ARK] "invalid size of partial point"
    with pytest.raises(ValueError):
        T.DenseMultilinearExtension.from_evaluations_vec(F, 3, list(range(7)))
    v = T.Verifier(3, None, F)
    v.set_c_1(5)
    with pytest.raises(T.ProverClaimMismatch):
        v.round(T.SparsePolynomial(F, [(0, 1)]), PyRng(O.FP1572869))  # g(0)+g(1) = 2 != 5
    # no oracle access on the last round -> NoPolySet
    p = T.Prover(g)
    v = T.Verifier(3, None, F)
    v.set_c_1(p.c_1())
    rng = PyRng(O.FP1572869, 1)
    r = 1
    with pytest.raises(T.NoPolySet):
        for j in range(3):
            kind, r = v.round(p.round(r, j), rng)


@pytest.mark.parametrize("n", [1, 2, 5])
def test_strict_verifier_rejects_a_false_claim_the_reference_logic_accepts(n):
    """The reference's last-round branch (sum-check-protocol/src/lib.rs:298-310) never links g_n to g_{n-1}: a prover
    who claims c_1 + 1, shifts g_j by 2^-j for j < n and then sends the honest g_n is accepted.  The default (option
    strict_verifier = 1) rejects it; honest runs are accepted either way with the same bytes."""
    OF, F = O.FP1572869, T.Field(1572869)
    rnd = random.Random(77 + n)
    t = [rand_table(OF, n, rnd) for _ in range(2)]
    g = T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, n, x) for x in t])
    half = pow(2, -1, OF.p)

    def run(strict, cheat):
        T.set_option("strict_verifier", 1 if strict else 0)
        prover = T.Prover(g)
        ver = T.Verifier(n, g)
        ver.set_c_1((prover.c_1() + (1 if cheat else 0)) % OF.p)
        rng, r, out = PyRng(OF, 3), 1, None
        for j in range(n):
            g_j = prover.round(r, j)
            if cheat and (j < n - 1 or n == 1):
                # g_j' = g_j + 2^-(j+1): g_1'(0) + g_1'(1) = c_1 + 1 and g_j'(0) + g_j'(1) = g_{j-1}'(r_{j-1}), so every
                # check the reference makes before its last round passes; the last message is the honest g_n
                d = dict(g_j.coeffs)
                d[0] = (d.get(0, 0) + pow(half, j + 1, OF.p)) % OF.p
                g_j = T.SparsePolynomial(F, sorted(d.items()))
            kind, val = ver.round(g_j, rng)
            if kind == "JthRound":
                r = val
            else:
                out = val
        return out

    assert run(True, False) is True and run(False, False) is (True if n > 1 else None)
    if n == 1:
        assert run(True, True) is False          # the only round evaluates the oracle: g_1'(r) != g(r)
        assert run(False, True) is None          # reference: the first-round branch returns before any oracle check
    else:
        with pytest.raises(T.ProverClaimMismatch):
            run(True, True)
        assert run(False, True) is True          # the reference's literal logic accepts the false claim


def test_truncated_transcript_is_not_accepted():
    OF, F = O.FP389, T.Field(389)
    rnd = random.Random(5)
    g = T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, 4, rand_table(OF, 4, rnd)) for _ in range(2)])
    tr = T.generate_transcript(T.Prover(g))
    assert T.verify_transcript(tr, T.Verifier(4, g))
    assert not T.verify_transcript(tr[:3], T.Verifier(4, g))
    T.set_option("strict_verifier", 0)  # fiat-shamir/src/lib.rs:131-141 loops over whatever it is given
    assert T.verify_transcript(tr[:3], T.Verifier(4, g))


# ----------------------------------------------------------------------------- size-independent properties at scale
@pytest.mark.parametrize("OF,v,K", [(O.FP1572869, 22, 3), (O.FP1572869, 26, 3), (O.Field(268435399), 25, 4), (O.FP389, 23, 2), (O.BLS12_381_FR, 18, 3),
                                    (O.Field((1 << 61) - 1), 20, 2)], ids=lambda x: str(getattr(x, "bits", x)))
def test_large_prover_verifier_invariants(OF, v, K):
    """At sizes the oracle would take too long for: the verifier's checks g_j(0)+g_j(1) = g_{j-1}(r_{j-1}) and
    g_v(r_v) = g(r) (sum-check-protocol/src/lib.rs:286-291,302-307,316-323) with the final oracle evaluated by the
    independent eq-table kernel, plus c_1 against a host-side sum of the D2H'd product table."""
    F = T.Field(OF.p)
    tabs = [T.DenseMultilinearExtension.synthetic(F, v, 1000 + k) for k in range(K)]
    g = T.ProductMLE.new(tabs)
    assert run_protocol(g, PyRng(OF, v)) is True
    transcript = T.generate_transcript(T.Prover(g))
    assert len(transcript) == v
    assert T.verify_transcript(transcript, T.Verifier(v, g))
    if OF.n_limbs == 1 and v <= 22:  # a 2^v-entry Python list
        ev = g.to_evaluations()
        assert T.Prover(g).c_1() == sum(ev) % OF.p


# ----------------------------------------------------------------------------- persistent tail kernel
def test_tail_kernel_matches_per_round_launches():
    """The resident tail kernel (mailbox in mapped memory) must give the same transcript bytes as one launch per
    round, for every threshold; SCB_TAIL_VARS is read once per process, hence the subprocesses."""
    import os
    import subprocess
    import sys

    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import thaler_study_b200 as T; T.options_from_env()\n"
        "for p, v, K in ((1572869, 13, 3), (5, 9, 2), (0xFFFFFFFF00000001, 11, 4), (%d, 10, 3)):\n"
        "    F = T.Field(p)\n"
        "    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 50 + k) for k in range(K)])\n"
        "    print(b''.join(T.generate_transcript(T.Prover(g))).hex())\n"
        "    a = T.DenseMultilinearExtension.synthetic(F, v, 7); b = T.DenseMultilinearExtension.synthetic(F, v, 8)\n"
        "    print(b''.join(T.generate_transcript(T.Prover(T.MatMulG.from_tables(a, b)))).hex())\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), O.BLS12_381_FR.p)
    outs = []
    for tv in ("0", "2", "5", "14", "24"):
        env = dict(os.environ, SCB_TAIL_VARS=tv)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300))
    for o in outs:
        assert o.returncode == 0, o.stderr[-2000:]
    assert len(set(o.stdout for o in outs)) == 1
    assert len(outs[0].stdout.split()) == 8
    # and the SCB_TAIL_VARS=0 output is the oracle's transcript
    OF = O.FP1572869
    cf = CField(OF.p)
    tabs = [cf.from_mont(cf.synth(50 + k, 0, 1 << 13)) for k in range(3)]
    want = O.generate_transcript(OF, O.Prover(O.ProductMLE(OF, [O.DenseMLE(OF, 13, t) for t in tabs])))
    assert outs[0].stdout.split()[0] == b"".join(want).hex()


def test_grid_resident_kernel_matches_per_round_launches():
    """The grid-wide resident kernel (persist.cuh: all rounds in one cooperative launch, ticket/flag barrier between
    rounds) must give the same transcript bytes as one launch per round.  SCB_PERSIST_VARS=4 forces it onto tiny
    tables too (single active CTA, the u32 quad-pair path running out of pairs)."""
    import os
    import subprocess
    import sys

    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import thaler_study_b200 as T; T.options_from_env()\n"
        "for p, v, K in ((1572869, 19, 3), (1572869, 16, 4), (1572869, 6, 1), (5, 15, 2), (0xFFFFFFFF00000001, 17, 3), (%d, 15, 2)):\n"
        "    F = T.Field(p)\n"
        "    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 90 + k) for k in range(K)])\n"
        "    print(b''.join(T.generate_transcript(T.Prover(g))).hex())\n"
        "    a = T.DenseMultilinearExtension.synthetic(F, v, 7); b = T.DenseMultilinearExtension.synthetic(F, v, 8)\n"
        "    print(b''.join(T.generate_transcript(T.Prover(T.MatMulG.from_tables(a, b)))).hex())\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), O.BLS12_381_FR.p)
    outs = []
    for env_add in ({"SCB_TAIL_VARS": "0"}, {"SCB_PERSIST_VARS": "0"}, {"SCB_PERSIST_VARS": "4"}, {}, {"SCB_PERSIST_VARS": "17"}):
        env = dict(os.environ, **env_add)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600))
    for o in outs:
        assert o.returncode == 0, o.stderr[-2000:]
    assert len(outs[0].stdout.split()) == 12
    assert len(set(o.stdout for o in outs)) == 1


@pytest.mark.parametrize("p", [1572869, 389, 0xFFFFFFFF00000001, O.BLS12_381_FR.p], ids=lambda p: f"p{p.bit_length()}")
def test_evaluate_many_matches_single_evaluations_and_the_oracle(p):
    """scb_mle_evaluate_many (restrict_poly's k + 1 evaluations along a line, gkr-protocol/src/lib.rs:291-321): the
    row-wise kernel (one-limb fields, 2^8 .. 2^20 entries), the staged one and the single-evaluation fallback against
    the C oracle's evaluation, 1 .. 21 points."""
    from oracle.coracle import CField

    F, cf = T.Field(p), CField(p)
    rnd = random.Random(p % 977)
    for v in (3, 8, 9, 13, 17):
        m = T.DenseMultilinearExtension.synthetic(F, v, 5 + v)
        tab = cf.synth(5 + v, 0, 1 << v)
        for n_pts in (1, 7, 8, 9, 21):
            pts = [[rnd.randrange(p) for _ in range(v)] for _ in range(n_pts)]
            want = [cf.from_mont(cf.mle_evaluate_le(tab, cf.to_mont(pt)))[0] for pt in pts]
            for rows in (1, 0):
                T.set_option("mle_rows_multi", rows)
                assert m.evaluate_many(pts) == want, (v, n_pts, rows)
    T.reset_options()
