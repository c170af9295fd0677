"""CPU checks of the identities and integer bounds behind the 21-bit-triple pair pass (thaler_study_b200/csrc/pairs.cuh:
k_grid_sp_pf_w21, k_pair_pass_sp_w21, fold2) -- the device code restated with Python integers, against the reference's fold

    t[b] = t[2b] + r (t[2b+1] - t[2b])          ([ARK] DenseMultilinearExtension::fix_variables)

applied twice.  Values live in ark's Montgomery form with R = 2^64 (one limb); the small-prime policy multiplies with ONE
32-bit Montgomery step (redc32) and keeps the missing 2^-32 in per-pass constants (field.cuh: PolSP)."""
import random

import pytest

MODULI = [5, 389, 1572869, 268435399]  # the reference's three fields and the largest prime below 2^28 the policy admits
R = 1 << 64


def redc32(T, p, ninv):
    """field.cuh PolSP::redc32: T * 2^-32 mod p, result < T / 2^32 + p."""
    m = ((T & 0xFFFFFFFF) * ninv) & 0xFFFFFFFF
    s = T + m * p
    assert s & 0xFFFFFFFF == 0 and s < (1 << 64)
    return s >> 32


def reduce_once(a, p):
    return a - p if a >= p else a


def fold_const(r_m, p, ninv):
    return reduce_once(redc32(r_m, p, ninv), p)


def fold_c(t0, t1, rc, p, ninv):
    return reduce_once(reduce_once(t0 + redc32((t1 - t0 + p) * rc, p, ninv), p), p)


def mont_mul(a, b, p):  # PolSP::mul: a b 2^-64 mod p, canonical
    return a * b * pow(R, -1, p) % p


def fold2_const(ra_m, rb_m, p, ninv):
    one_m = R % p
    na, nb = (one_m - ra_m) % p, (one_m - rb_m) % p
    return [fold_const(mont_mul(x, y, p), p, ninv) for x, y in ((na, nb), (ra_m, nb), (na, rb_m), (ra_m, rb_m))]


def fold2(t, w, p, ninv):
    s = t[0] * w[0] + t[1] * w[1] + t[2] * w[2] + t[3] * w[3]
    assert s < 4 * p * p <= (1 << 58)  # four unreduced products fit the 64-bit sum with room to spare
    x = redc32(s, p, ninv)
    assert x < p + p // 4 + 1  # one conditional subtraction is enough
    return reduce_once(x, p)


@pytest.mark.parametrize("p", MODULI)
def test_bilinear_two_variable_fold_equals_the_nested_folds(p):
    ninv = (-pow(p, -1, 1 << 32)) % (1 << 32)
    rnd = random.Random(p)
    edge = [0, 1, p - 1, p - 2, R % p]
    for it in range(400):
        pick = (lambda: rnd.choice(edge)) if it < 60 else (lambda: rnd.randrange(p))
        ra, rb = pick(), pick()  # canonical challenges
        ra_m, rb_m = ra * R % p, rb * R % p
        x = [pick() for _ in range(4)]  # canonical table entries x[y1 + 2 y2]
        t = [v * R % p for v in x]
        # the reference: fix variable 0 by ra, then (what was) variable 1 by rb
        lo, hi = (x[0] + ra * (x[1] - x[0])) % p, (x[2] + ra * (x[3] - x[2])) % p
        want = (lo + rb * (hi - lo)) % p * R % p
        rac, rbc = fold_const(ra_m, p, ninv), fold_const(rb_m, p, ninv)
        nested = fold_c(fold_c(t[0], t[1], rac, p, ninv), fold_c(t[2], t[3], rac, p, ninv), rbc, p, ninv)
        w = fold2_const(ra_m, rb_m, p, ninv)
        assert all(0 <= wi < p for wi in w)
        assert nested == want
        assert fold2(t, w, p, ninv) == want


@pytest.mark.parametrize("p", [5, 389, 1572869])
def test_21_bit_triples_round_trip(p):
    """word i = A[i] | B[i] << 21 | C[i] << 42 (k_grid_sp_pf_w21) and its three fields back (k_pair_pass_sp_w21): exact for
    canonical Montgomery values of a field of at most 21 bits -- which is what the engine requires of its inputs."""
    assert p.bit_length() <= 21
    rnd = random.Random(7 * p)
    for it in range(2000):
        a, b, c = (rnd.choice([0, 1, p - 1]) if it < 30 else rnd.randrange(p) for _ in range(3))
        w = a | (b << 21) | (c << 42)
        assert w < (1 << 63)
        assert [(w >> (21 * k)) & 0x1FFFFF for k in range(3)] == [a, b, c]


def test_bytes_per_index_of_a_proof_with_triples():
    """DESIGN.md 4a / bench.py: K = 3, E = 8 -- 24 + 8 (grid pass) + 8 + 3 (first pair pass) + 3 (1 + 1/4 + ...) + 0.75 (...) bytes."""
    import bench

    v, K, E = 28, 3, 8
    pb = bench.pass_bytes(v, K, E)
    assert pb[0] == K * ((1 << v) * E + (1 << (v - 2)) * 4)
    with_triples = K * (1 << v) * E + (1 << v) * 8 + ((1 << v) * 8 + K * (1 << (v - 2)) * 4) + sum(pb[1:])
    without = K * (1 << v) * E + sum(pb)
    assert without - with_triples == 8 << v  # 24 bytes per index read a second time against 8 written + 8 read
    assert abs(with_triples / (1 << v) - 48.0) < 0.01 and abs(without / (1 << v) - 56.0) < 0.01
