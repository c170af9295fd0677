"""CPU checks (Python integers) of the identities the fourth-generation 4-limb kernels rest on (csrc/g4.cuh, g4_mle.cuh,
DESIGN.md section 4b): the GPU tests compare the kernels' results with the oracle; these pin the bounds the unreduced
accumulators and the table fold were sized on, for the three 4-limb moduli the GPU tests use."""
import random

import pytest

BLS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BN254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617
P255 = (1 << 255) - 19
R = 1 << 256
W32 = (1 << 32) - 1


def redc_words(t, p, steps):
    """`steps` word-serial Montgomery steps on the integer t: returns (t + m p) / 2^(32 steps) with m < 2^(32 steps)"""
    n0 = (-pow(p, -1, 1 << 32)) % (1 << 32)
    for _ in range(steps):
        m = ((t & W32) * n0) & W32
        t = (t + m * p) >> 32
    return t


@pytest.mark.parametrize("p", [BLS, BN254, P255])
def test_table_fold_is_r_times_d_below_2p(p):
    """ArithT::mul_fixed_raw: T_i = r 2^(32 i + 64) mod p; W = sum_i d_i T_i < 2^35 p; two REDC word steps give a value
    congruent to r d, below p (1 + 2^-29) -- one conditional subtraction makes it canonical.  d = t1 - t0 + p in (0, 2p)."""
    rnd = random.Random(p & 0xFFFF)
    for _ in range(200):
        r = rnd.randrange(p)  # the challenge as a plain integer (engine.cu takes it out of Montgomery form)
        tab = [(r << (32 * i + 64)) % p for i in range(8)]
        t0, t1 = rnd.randrange(p), rnd.randrange(p)  # Montgomery-form table entries are just residues here
        d = t1 - t0 + p
        assert 0 < d < 2 * p < R
        w = sum(((d >> (32 * i)) & W32) * tab[i] for i in range(8))
        assert w < (p << 35)
        out = redc_words(w, p, 2)
        assert out % p == (r * (t1 - t0)) % p
        assert out < p + (p >> 29) + 1 and out < 2 * p
    # extreme operands
    for r, t0, t1 in ((p - 1, 0, p - 1), (p - 1, p - 1, 0), (0, 5, 7), (1, p - 1, p - 1)):
        tab = [(r << (32 * i + 64)) % p for i in range(8)]
        d = t1 - t0 + p
        w = sum(((d >> (32 * i)) & W32) * tab[i] for i in range(8))
        assert redc_words(w, p, 2) % p == (r * (t1 - t0)) % p and redc_words(w, p, 2) < 2 * p


@pytest.mark.parametrize("p", [BLS, BN254, P255])
def test_unreduced_sums_reduce_to_the_sum_of_montgomery_products(p):
    """wacc_reduce / wide_reduce: for T = sum_i P_i c_i = C0 + C1 2^256 + C2 2^512 (factors below p, up to 2^24 terms)
    T / R = C0 / R + C1 + C2 R (mod p) = the sum of the Montgomery products P_i c_i / R; T fits 544 bits."""
    rnd = random.Random(3)
    rinv = pow(R, -1, p)
    for n in (1, 2, 33, 1000):
        terms = [(rnd.randrange(p), rnd.randrange(p)) for _ in range(n)] + [(p - 1, p - 1)] * 3
        t = sum(a * b for a, b in terms)
        assert t < 1 << 544 and (1 << 24) * (p - 1) ** 2 < 1 << 544
        c0, c1, c2 = t & (R - 1), (t >> 256) & (R - 1), t >> 512
        assert c2 < 1 << 32
        want = sum(a * b * rinv for a, b in terms) % p
        assert (c0 * rinv + c1 + c2 * R) % p == want
        # the kernel's way: montmul(C0, 1), montmul(montmul(C1, R^2), 1), montmul(C2, R^2)
        mm = lambda x, y: x * y * rinv % p
        r2 = R * R % p
        assert (mm(c0, 1) + mm(mm(c1, r2), 1) + mm(c2, r2)) % p == want


@pytest.mark.parametrize("p", [BLS, BN254, P255])
def test_first_level_product_at_two_by_interpolation(p):
    """k_round_evals_g4w, K = 3: q(X) = (a0 + X da)(b0 + X db) is quadratic, so q(2) = 2 q(1) - q(0) + 2 q(inf)"""
    rnd = random.Random(4)
    for _ in range(100):
        a0, a1, b0, b1 = (rnd.randrange(p) for _ in range(4))
        da, db = (a1 - a0) % p, (b1 - b0) % p
        q0, q1, qinf = a0 * b0 % p, a1 * b1 % p, da * db % p
        q2 = (a1 + da) * (b1 + db) % p
        assert (2 * q1 - q0 + 2 * qinf) % p == q2


def test_p_equal_one_mod_2_32_needs_no_n0_multiplication():
    """ArithT<true>: p = 1 (mod 2^32) => n0 = -p^-1 = 2^32 - 1 and m p[0] = m, so word 0 + m = 0 (mod 2^32) with carry
    (word 0 != 0); BLS12-381 Fr qualifies, BN254's scalar field and 2^255 - 19 do not."""
    assert BLS % (1 << 32) == 1 and BN254 % (1 << 32) != 1 and P255 % (1 << 32) != 1
    n0 = (-pow(BLS, -1, 1 << 32)) % (1 << 32)
    assert n0 == W32
    for e0 in (0, 1, 12345, W32):
        m = (e0 * n0) & W32
        assert m == (-e0) & W32 and (e0 + m) in (0, 1 << 32) and ((e0 + m) >> 32) == (1 if e0 else 0)
