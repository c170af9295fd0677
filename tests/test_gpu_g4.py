"""The 4-limb kernels of csrc/g4.cuh (one point fewer thanks to the claim, leading coefficient instead of the highest
point, lazy sums, leaner carry chains; fourth generation: unreduced last products in 544-bit shared-memory accumulators,
p = 1 (mod 2^32) variant) and csrc/g29.cuh against the C oracle and against round 1's kernel."""
import random

import numpy as np
import pytest

from oracle import pyoracle as O
from oracle.coracle import CField

import thaler_study_b200 as T

pytestmark = pytest.mark.gpu

BLS = O.BLS12_381_FR.p
BN254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # 254 bits
P255 = (1 << 255) - 19                                                                   # 255 bits, top word all ones
P4 = [BLS, BN254, P255]
pid = lambda p: f"p{p.bit_length()}_{p % 1000}"


@pytest.mark.parametrize("p", P4, ids=pid)
def test_g4_kernel_matches_oracle_and_first_generation(p):
    F, cf = T.Field(p), CField(p)
    assert F.policy == 4
    rnd = random.Random(p & 0xFFFF)
    for v in (2, 3, 4, 7, 11, 14):
        seeds = [rnd.randrange(1 << 20) for _ in range(4)]
        tabs_c = [cf.synth(s, 0, 1 << v) for s in seeds]
        tabs_g = [T.DenseMultilinearExtension.synthetic(F, v, s) for s in seeds]
        r = rnd.randrange(p)
        for K in (1, 2, 3, 4):
            g = T.ProductMLE.new(tabs_g[:K])
            f_c = [cf.fix_variable(t, cf.to_mont([r])) for t in tabs_c[:K]]
            want = cf.from_mont(cf.product_round_evals(f_c, K + 1))
            claim = (want[0] + want[1]) % p
            # wide accumulators (generic / p = 1 mod 2^32 variant), radix-2^29 lazy carries (g29.cuh), 32-bit-limb chains, round 1's kernel
            for flag, p0one in ((3, 1), (3, 0), (2, 1), (1, 1), (0, 1)):
                T.set_option("g4_kernel", flag)
                T.set_option("g4_p0one", p0one)
                g_new, ev_new = g.fix_and_round_evals(r, claim=claim)
                assert ev_new == want, (v, K, flag, p0one)
                for k in range(K):
                    assert np.array_equal(g_new.table(k).to_evaluations_mont(), f_c[k]), (v, K, flag)
                # Prover::new's pass (no claim): k_round_evals_g4w / k_round_evals_g29 / k_round_evals
                assert g_new.round_evals() == want, (v, K, flag, p0one)
            _, ev_plain = g.fix_and_round_evals(r)
            assert ev_plain == want, (v, K)
    T.reset_options()


@pytest.mark.parametrize("p", P4, ids=pid)
def test_g4_kernel_extreme_values(p):
    """Entries 0 and p-1 in every pattern, challenge p-1: the lazy difference t1 - t0 + p reaches both ends of (0, 2p),
    the unreduced products their upper bound, and every sum wraps many times."""
    F, OF = T.Field(p), O.Field(p)
    v = 9
    pats = [[0, p - 1, p - 1, 0], [p - 1, 0, 0, p - 1], [p - 1] * 4, [p - 1, p - 2, 1, 0]]
    for K in (1, 2, 3, 4):
        vals = [(pats[(k + K) % 4] * (1 << (v - 2))) for k in range(K)]
        og = O.ProductMLE(OF, [O.DenseMLE(OF, v, t) for t in vals])
        g = T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, v, t) for t in vals])
        for r in (p - 1, 0, 1, (p + 1) // 2):
            of = og.fix_variables([r])
            want = of.round_evals()
            for flag, p0one in ((3, 1), (3, 0), (2, 1), (1, 1)):
                T.set_option("g4_kernel", flag)
                T.set_option("g4_p0one", p0one)
                g2, ev = g.fix_and_round_evals(r, claim=(want[0] + want[1]) % p)
                assert ev == want, (K, r, flag, p0one)
                assert g2.round_evals() == want, (K, r, flag, p0one)
                for k in range(K):
                    assert g2.table(k).to_evaluations() == of.tables[k].evals
    T.reset_options()


@pytest.mark.parametrize("p", [BLS, BN254], ids=pid)
def test_transcripts_through_the_g4_kernel(p):
    """With the resident kernels switched off every round after the first is a per-round launch, which passes the claim
    and therefore runs the g4 kernel: transcript bytes against the Python oracle, K = 1..4, and against g4_kernel = 0."""
    OF, F = O.Field(p), T.Field(p)
    rnd = random.Random(7)
    T.set_option("tail_vars", 0)
    for K, v in ((1, 6), (2, 7), (3, 8), (4, 5)):
        vals = [[rnd.randrange(p) for _ in range(1 << v)] for _ in range(K)]
        og = O.ProductMLE(OF, [O.DenseMLE(OF, v, t) for t in vals])
        want = O.generate_transcript(OF, O.Prover(og))
        outs = []
        for flag in (3, 2, 1, 0):
            T.set_option("g4_kernel", flag)
            g = T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, v, t) for t in vals])
            T.launch_count(reset=True)
            outs.append(T.generate_transcript(T.Prover(g)))
            assert T.launch_count() == v  # Prover::new's pass + one launch per later round: no resident kernel ran
        assert all(o == want for o in outs), (K, v)
    # matrix_multiplication::G (interpolate_quadratic_poly conventions) through the same kernel
    v = 8
    a, b = [rnd.randrange(p) for _ in range(1 << v)], [rnd.randrange(p) for _ in range(1 << v)]
    og = O.MatMulG(OF, O.DenseMLE(OF, v, a), O.DenseMLE(OF, v, b))
    T.set_option("g4_kernel", 2)
    dg = T.MatMulG.from_tables(T.DenseMultilinearExtension.from_evaluations_vec(F, v, a), T.DenseMultilinearExtension.from_evaluations_vec(F, v, b))
    assert T.generate_transcript(T.Prover(dg)) == O.generate_transcript(OF, O.Prover(og))
    T.reset_options()


def test_g4_proof_2_24_matches_c_oracle():
    """2^24 entries: rounds 1 and 2 are per-round launches of the g4 kernel under the default options (the resident
    kernel takes over at 2^22), the rest resident; every message against the C oracle."""
    try:
        from test_gpu_fullsize import oracle_messages
    except ImportError:
        from tests.test_gpu_fullsize import oracle_messages

    OF, cf, K, v = O.BLS12_381_FR, CField(BLS), 3, 24
    F = T.Field(BLS)
    seeds = [0xB200 + k for k in range(K)]
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, s) for s in seeds])
    prover = T.Prover(g)
    got_c1 = prover.c_1()
    transcript = T.generate_transcript(prover)
    assert T.verify_transcript(transcript, T.Verifier(v, g))
    tabs_c = [cf.synth(s, 0, 1 << v) for s in seeds]
    want_c1, want = oracle_messages(OF, cf, tabs_c, K, transcript)
    assert got_c1 == want_c1 and transcript == want


def test_wide_kernels_for_every_resident_cta_count():
    """The K = 3 kernels of the fourth generation exist in builds for one, two and three resident CTAs per SM (option
    g4_blocks4; 0 = the measured defaults): same sums, same folded tables, whole transcripts."""
    p = BLS
    F, cf, OF = T.Field(p), CField(p), O.BLS12_381_FR
    rnd = random.Random(11)
    v, K = 13, 3
    seeds = [rnd.randrange(1 << 20) for _ in range(K)]
    tabs_c = [cf.synth(s, 0, 1 << v) for s in seeds]
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, s) for s in seeds])
    r = rnd.randrange(p)
    f_c = [cf.fix_variable(t, cf.to_mont([r])) for t in tabs_c]
    want = cf.from_mont(cf.product_round_evals(f_c, K + 1))
    want0 = cf.from_mont(cf.product_round_evals(tabs_c, K + 1))
    T.set_option("tail_vars", 0)
    outs = []
    for blocks in (0, 1, 2, 3):
        for p0one in (1, 0):
            T.set_option("g4_blocks4", blocks)
            T.set_option("g4_p0one", p0one)
            assert g.round_evals() == want0, (blocks, p0one)
            g2, ev = g.fix_and_round_evals(r, claim=(want[0] + want[1]) % p)
            assert ev == want, (blocks, p0one)
            for k in range(K):
                assert np.array_equal(g2.table(k).to_evaluations_mont(), f_c[k]), (blocks, p0one)
            outs.append(T.generate_transcript(T.Prover(g)))
    assert all(o == outs[0] for o in outs)
    T.reset_options()
