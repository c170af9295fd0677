"""GPU tests of the gate-list GKR layer prover (csrc/gkr.cuh) against the Python oracle, which follows the reference's
DENSE formulation (gkr-protocol/src/lib.rs:373-456, round_polynomial.rs): every message must be the same polynomial."""
import random

import numpy as np
import pytest

from oracle import pyoracle as O

import thaler_study_b200 as T
from thaler_study_b200.gkr import ADD, MUL, Circuit, GkrProver, GkrVerifier

pytestmark = pytest.mark.gpu
GENERATOR = {O.BLS12_381_FR.p: 7, 0xFFFFFFFF00000001: 7}  # multiplicative generators (2 for the reference fields)


class PyRng:
    def __init__(self, F, seed=0):
        self.F, self.r = F, random.Random(seed)

    def draw(self):
        return self.r.randrange(self.F.p)


def to_oracle_circuit(layers, num_inputs):
    return O.Circuit([[(O.ADD if t == ADD else O.MUL, ins) for t, ins in l] for l in layers], num_inputs)


BOOK = [[(MUL, (0, 1)), (MUL, (2, 3))], [(MUL, (0, 0)), (MUL, (1, 1)), (MUL, (1, 2)), (MUL, (3, 3))]]   # circuit.rs:215-253
THREE = [[(ADD, (0, 1)), (ADD, (2, 3))], [(ADD, (0, 1)), (ADD, (2, 3)), (ADD, (4, 5)), (ADD, (6, 7))]]  # lib.rs:488-504


def run_gkr(F, circuit, inp, expected_outputs, rng):
    """gkr-protocol/src/lib.rs:574-623 on the device types."""
    prover = GkrProver(circuit, inp)
    begin = prover.start_protocol()
    if expected_outputs is not None:
        assert begin == ("Begin", expected_outputs)
    verifier = GkrVerifier(circuit)
    kind, r_i = verifier.receive_prover_msg(begin, rng)
    assert kind == "R"
    for i in range(circuit.layers_len()):
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


def test_protocol_test_from_book():  # lib.rs:550-624
    F = T.Field(389)
    c = Circuit(F, BOOK, 4)
    assert [c.num_vars_at(i) for i in range(3)] == [1, 2, 2]
    p = GkrProver(c, [3, 2, 3, 1])
    assert [p.layer(i).to_evaluations() for i in range(3)] == [[36, 6], [9, 4, 6, 1], [3, 2, 3, 1]]  # circuit.rs:263-267
    assert run_gkr(F, c, [3, 2, 3, 1], [36, 6], PyRng(O.FP389, 19))


def test_three_layer_protocol_test():  # lib.rs:626-703
    F = T.Field(389)
    assert run_gkr(F, Circuit(F, THREE, 8), [0, 1] * 4, [2, 2], PyRng(O.FP389, 23))


def random_layers(rnd, sizes, num_inputs):
    layers = []
    for i, s in enumerate(sizes):
        nxt = sizes[i + 1] if i + 1 < len(sizes) else num_inputs
        layers.append([(rnd.choice((ADD, MUL)), (rnd.randrange(nxt), rnd.randrange(nxt))) for _ in range(s)])
    return layers


@pytest.mark.parametrize("OF", [O.FP389, O.FP1572869, O.Field(0xFFFFFFFF00000001), O.BLS12_381_FR], ids=lambda F: f"p{F.bits}")
def test_messages_equal_dense_reference_formulation(OF):
    """Same challenges into the oracle's dense-table prover and the gate-list prover: identical c_1 and identical round
    polynomials for every layer.  q is the same POLYNOMIAL; its term list can differ from the reference's in one
    documented way: restrict_poly (gkr-protocol/src/lib.rs:291-321) multiplies sparse
    polynomials term by term and can keep explicit zero-coefficient terms, the engine sends the unique interpolant
    with zero terms dropped.  The comparison below is therefore on the non-zero terms, and
    test_restrict_poly_zero_term_deviation pins a case where the two term lists really differ."""
    F = T.Field(OF.p)
    rnd = random.Random(OF.bits)
    gen = GENERATOR.get(OF.p, 2)
    for sizes, n_in in (([2, 4], 4), ([4, 8, 4], 8), ([2, 2, 8], 4), ([8, 8], 16)):
        layers = random_layers(rnd, sizes, n_in)
        inp = [rnd.randrange(OF.p) for _ in range(n_in)]
        oc, dc = to_oracle_circuit(layers, n_in), Circuit(F, layers, n_in)
        op, dp = O.GkrProver(OF, oc, inp, generator=gen), GkrProver(dc, inp)
        assert dp.start_protocol() == op.start_protocol()
        two_adic_ok = OF.two_adicity() >= 2
        for i in range(len(sizes)):
            r_i = [rnd.randrange(OF.p) for _ in range(oc.num_vars_at(i))]
            dmsg = dp.start_round(i, r_i)
            num_vars = 2 * oc.num_vars_at(i + 1)
            if two_adic_ok:
                omsg = op.start_round(i, r_i)
                assert dmsg == omsg, (sizes, i)
            else:  # the reference's size-4 FFT domain does not exist: compare against the dense W summed directly
                add_i, mul_i = oc.wiring_tables(OF, i)
                w = O.DenseMLE(OF, oc.num_vars_at(i + 1), op.layers[i + 1])
                ow = O.GkrW(OF, add_i.fix_variables(r_i), mul_i.fix_variables(r_i), w, w.clone())
                assert dmsg[1] == sum(ow.to_evaluations()) % OF.p
            def both_receive(r):
                dp.receive_verifier_msg(("SumCheckRoundResult", ("JthRound", r)))
                if two_adic_ok:
                    op.receive_verifier_msg(("SumCheckRoundResult", ("JthRound", r)))

            for j in range(num_vars - 1):  # message j is sent before challenge r_j is drawn
                dm = dp.round_msg(j)
                assert dm[0] == "SumCheckProverMessage"
                if two_adic_ok:
                    assert dm[1].coeffs == op.round_msg(j)[1].coeffs, (sizes, i, j)
                both_receive(rnd.randrange(OF.p))
            both_receive(rnd.randrange(OF.p))  # Verifier::final_random_point
            dm = dp.round_msg(num_vars - 1)
            assert dm[0] == "FinalRoundMessage"
            if two_adic_ok:
                om = op.round_msg(num_vars - 1)
                assert dm[1].coeffs == om[1].coeffs
                assert dm[2].coeffs == [(d, c) for d, c in om[2].coeffs if c != 0]
            half = num_vars // 2
            b, c = dp.r[:half], dp.r[half:]
            w = O.DenseMLE(OF, half, op.layers[i + 1])
            for t in (0, 1, 5):
                assert dm[2].evaluate(t) == w.evaluate([(bb + t * (cc - bb)) % OF.p for bb, cc in zip(b, c)])


def test_restrict_poly_zero_term_deviation():
    """b[bit] = 0 puts an explicit (0, 0) term into the reference's line factor (from_coefficients_vec strips only
    TRAILING zeros) and the term-by-term products carry it into q; the engine's q is the unique interpolant with zero
    terms dropped.  Same values everywhere, different term lists -- the one representation difference to the
    reference, stated in DESIGN.md section 5."""
    OF, F = O.FP389, T.Field(389)
    evals, b, c = [0, 5, 0, 0], [0, 0], [3, 2]
    want = O.restrict_poly(OF, b, c, O.DenseMLE(OF, 2, evals))
    assert want.coeffs[0] == (0, 0)  # the explicit zero term of the reference's bookkeeping
    m = T.DenseMultilinearExtension.from_evaluations_vec(F, 2, evals)
    line = lambda t: [(bb + t * (cc - bb)) % OF.p for bb, cc in zip(b, c)]
    got = T.evals_to_univariate(F, T.KIND_GKR_W, [m.evaluate(line(t)) for t in range(3)])
    assert got.coeffs == [(d, cf) for d, cf in want.coeffs if cf != 0] and got.coeffs != list(want.coeffs)
    for t in range(6):
        assert got.evaluate(t) == want.evaluate(t)


def test_wiring_eval_matches_dense_tables():
    OF, F = O.FP1572869, T.Field(1572869)
    rnd = random.Random(3)
    layers = random_layers(rnd, [4, 8], 8)
    oc, dc = to_oracle_circuit(layers, 8), Circuit(F, layers, 8)
    for i in range(2):
        kc, kn = oc.num_vars_at(i), oc.num_vars_at(i + 1)
        r_i = [rnd.randrange(OF.p) for _ in range(kc)]
        b = [rnd.randrange(OF.p) for _ in range(kn)]
        c = [rnd.randrange(OF.p) for _ in range(kn)]
        assert dc.wiring_eval(i, r_i, b, c) == (oc.add_i_ext(OF, r_i, i).evaluate(b + c), oc.mul_i_ext(OF, r_i, i).evaluate(b + c))


def big_circuit(F, width_bits, depth, seed):
    rng = np.random.default_rng(seed)
    S = 1 << width_bits
    sizes = [S] * depth
    types = rng.integers(0, 2, size=S * depth, dtype=np.uint8)
    in0 = rng.integers(0, S, size=S * depth, dtype=np.uint32)
    in1 = rng.integers(0, S, size=S * depth, dtype=np.uint32)
    return Circuit.from_arrays(F, sizes, types, in0, in1, S)


@pytest.mark.parametrize("p,width_bits,depth", [(1572869, 12, 3), (O.BLS12_381_FR.p, 10, 2)])
def test_wide_circuit_verifies(p, width_bits, depth):
    """Widths the dense formulation cannot reach (2^36 wiring entries per layer at width 2^12): the reference's
    verifier logic accepts the gate-list prover and check_input holds."""
    OF, F = O.Field(p), T.Field(p)
    c = big_circuit(F, width_bits, depth, 5)
    rng = np.random.default_rng(6)
    inp = F.to_mont([int(x) % p for x in rng.integers(0, 2**62, size=1 << width_bits)])
    assert run_gkr(F, c, inp, None, PyRng(OF, 7))


class ReplayRng:
    """Hands out pre-drawn challenges in order (public coins: the verifier's draws do not depend on the messages)."""

    def __init__(self, values):
        self.values, self.pos = list(values), 0

    def draw(self):
        v = self.values[self.pos]
        self.pos += 1
        return v


def run_gkr_batched(F, circuit, inp, rng):
    """Same protocol with prove_layer: the 2k challenges of a layer are drawn first, the prover produces the whole
    layer proof in one call, and the reference's verifier logic replays the messages against those challenges."""
    prover = GkrProver(circuit, inp)
    verifier = GkrVerifier(circuit)
    kind, r_i = verifier.receive_prover_msg(prover.start_protocol(), rng)
    transcripts = []
    for i in range(circuit.layers_len()):
        k = circuit.num_vars_at(i + 1)
        ch = [rng.draw() for _ in range(2 * k)]
        start, raw = prover.prove_layer(i, r_i, ch)
        msgs = prover.layer_messages(raw)
        transcripts.append((start, msgs))
        replay = ReplayRng(ch)
        verifier.receive_prover_msg(start, replay)
        for m in msgs[:-1]:
            verifier.receive_prover_msg(m, replay)
        verifier.final_random_point(replay)
        assert replay.pos == 2 * k
        kind, r_i = verifier.receive_prover_msg(msgs[-1], rng)  # draws the line point with the live rng
        assert kind == "R"
    return verifier.check_input(inp), transcripts


@pytest.mark.parametrize("p,width_bits,depth", [(389, 3, 3), (1572869, 9, 4), (O.BLS12_381_FR.p, 6, 2)])
def test_batched_layer_proof_equals_round_by_round(p, width_bits, depth):
    OF, F = O.Field(p), T.Field(p)
    c = big_circuit(F, width_bits, depth, 11)
    rng = np.random.default_rng(12)
    inp = F.to_mont([int(x) % p for x in rng.integers(0, 2**62, size=1 << width_bits)])
    ok, batched = run_gkr_batched(F, c, inp, PyRng(OF, 13))
    assert ok
    # round by round with the same challenge stream: every message must be the same polynomial
    prover, verifier, live = GkrProver(c, inp), GkrVerifier(c), PyRng(OF, 13)
    kind, r_i = verifier.receive_prover_msg(prover.start_protocol(), live)
    for i in range(c.layers_len()):
        k = c.num_vars_at(i + 1)
        ch = [live.draw() for _ in range(2 * k)]
        replay = ReplayRng(ch)
        start = prover.start_round(i, r_i)
        assert start == batched[i][0]
        verifier.receive_prover_msg(start, replay)
        for j in range(2 * k - 1):
            m = prover.round_msg(j)
            assert m[1] == batched[i][1][j][1], (i, j)
            prover.receive_verifier_msg(verifier.receive_prover_msg(m, replay))
        prover.receive_verifier_msg(verifier.final_random_point(replay))
        m = prover.round_msg(2 * k - 1)
        assert m[1] == batched[i][1][-1][1] and m[2] == batched[i][1][-1][2]
        kind, r_i = verifier.receive_prover_msg(m, live)
