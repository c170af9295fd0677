"""CPU test of csrc/lazy29.hpp (radix-2^29 lazy-carry arithmetic of the 4-limb kernels): the header is compiled for
the host as it is and checked against Python integers -- conversions, the k p "big limb" constant, the Montgomery
product x y 2^-261 with lazy operands at the documented bounds, and the head-room of the 64-bit column sums."""
import ctypes as C
import os
import random
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BLS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BN254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617
P255 = (1 << 255) - 19
M29 = (1 << 29) - 1


class Desc(C.Structure):
    _fields_ = [("p", C.c_uint32 * 9), ("kp", C.c_uint32 * 9), ("n0", C.c_uint32), ("k", C.c_uint32)]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("l29") / "lazy29_shim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "native", "lazy29_shim.cpp")])
    return C.CDLL(so)


def limbs_of(x, n=9):
    out = [(x >> (29 * j)) & M29 for j in range(n - 1)]
    out.append(x >> (29 * (n - 1)))
    return out


def val(limbs):
    return sum(int(l) << (29 * j) for j, l in enumerate(limbs))


def arr(xs):
    return (C.c_uint32 * len(xs))(*xs)


def desc(lib, p):
    d = Desc()
    p64 = (C.c_uint64 * 4)(*[(p >> (64 * i)) & (2**64 - 1) for i in range(4)])
    assert lib.l29_make_desc(p64, p.bit_length(), C.byref(d)) == 1
    return d


@pytest.mark.parametrize("p", [BLS, BN254, P255])
def test_desc_and_conversions(lib, p):
    d = desc(lib, p)
    assert val(d.p) == p and all(l <= M29 for l in list(d.p)[:8])
    assert (p * d.n0 + 1) % (1 << 29) == 0
    assert val(d.kp) == d.k * p and (1 << 257) <= d.k * p < (1 << 258)
    assert all(l >= M29 for l in list(d.kp)[:8]) and d.kp[8] >= 1 << 24
    rnd = random.Random(1)
    for x in [0, 1, p - 1, (1 << 256) - 1] + [rnd.randrange(1 << 256) for _ in range(200)]:
        w = arr([(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
        lim = (C.c_uint32 * 9)()
        lib.l29_from_words(w, lim)
        assert list(lim) == limbs_of(x)
        w9 = (C.c_uint32 * 9)()
        lib.l29_to_words(lim, w9)
        assert sum(int(v) << (32 * i) for i, v in enumerate(w9)) == x
    # lazy limbs (up to 2^32 - 1 each) through to_words and normalise: carries resolved, value kept (top word = overflow)
    for _ in range(200):
        lz = [rnd.randrange(1 << 32) for _ in range(8)] + [rnd.randrange(1 << 26)]
        w9 = (C.c_uint32 * 9)()
        lib.l29_to_words(arr(lz), w9)
        assert sum(int(v) << (32 * i) for i, v in enumerate(w9)) == val(lz)
        out = (C.c_uint32 * 9)()
        lib.l29_normalise(arr(lz), out)
        assert val(out) == val(lz) and all(l <= M29 for l in list(out)[:8])


@pytest.mark.parametrize("p", [BLS, BN254, P255])
def test_mont_and_lazy_difference(lib, p):
    d = desc(lib, p)
    rnd = random.Random(2)
    rinv = pow(1 << 261, -1, p)
    worst = 0
    cases = [(0, 0), (p - 1, p - 1), (1, p - 1)] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(300)]
    for a, b in cases:
        # canonical x canonical
        out, mc = (C.c_uint32 * 9)(), C.c_uint64()
        lib.l29_mont(C.byref(d), arr(limbs_of(a)), arr(limbs_of(b)), out, C.byref(mc))
        v = val(out)
        assert v % p == a * b * rinv % p and v < a * b // (1 << 261) + p + 1 and all(l <= M29 for l in list(out)[:8])
        worst = max(worst, mc.value)
        # lazy difference (a - b + k p, limbs up to 2^31) x canonical: the fold's product
        diff = (C.c_uint32 * 9)()
        lib.l29_sub_kp(C.byref(d), arr(limbs_of(a)), arr(limbs_of(b)), diff)
        assert val(diff) == a - b + d.k * p and max(diff) < 1 << 31
        lib.l29_mont(C.byref(d), diff, arr(limbs_of(b)), out, C.byref(mc))
        assert val(out) % p == (a - b) * b * rinv % p and val(out) < (d.k + 1) * p * p // (1 << 261) + p + 1
        worst = max(worst, mc.value)
        # v2 = hi + (hi - lo + k p): limbs up to 2^31.2, against a normalised product of earlier factors (below 2.6 p)
        v2 = [int(x) + int(y) for x, y in zip(limbs_of(a), diff)]
        prev = limbs_of((a * 7 + b) % (2 * p) + p // 2)
        lib.l29_mont(C.byref(d), arr(prev), arr(v2), out, C.byref(mc))
        assert val(out) % p == val(prev) * (2 * a - b) * rinv % p
        worst = max(worst, mc.value)
    # extreme limbs: every limb at its documented maximum
    big = [(1 << 31) - 1] * 8 + [1 << 26]  # v2 = hi + hi + kp - lo limb by limb: below 2^29 + 2^29 + 2^30
    nrm = [M29] * 8 + [(1 << 26) - 1]
    out, mc = (C.c_uint32 * 9)(), C.c_uint64()
    lib.l29_mont(C.byref(d), arr(nrm), arr(big), out, C.byref(mc))
    assert val(out) % p == val(nrm) * val(big) * rinv % p
    worst = max(worst, mc.value)
    assert worst < 0.76 * 2**64, worst / 2**64  # 9 (2^29 2^31 + 2^58) + carries = 0.70 * 2^64: the columns never wrap


@pytest.mark.parametrize("p", [BLS, BN254, P255])
def test_product_scanning_forms(lib, p):
    """mont_ps / mont_ps_par (scripts/mont29_bench.cu measures them on the GPU): same value as the interleaved form,
    the digit-split result with limbs below 2^30 + 2^6, and the p = 1 (mod 2^29) variant wherever the modulus allows."""
    d = desc(lib, p)
    rnd = random.Random(5)
    rinv = pow(1 << 261, -1, p)
    modes = [0, 2] + ([1, 3] if p % (1 << 29) == 1 else [])
    cases = [(0, 0), (p - 1, p - 1), (1, p - 1)] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(200)]
    for a, b in cases:
        for mode in modes:
            out, mc = (C.c_uint32 * 9)(), C.c_uint64()
            lib.l29_mont_ps(C.byref(d), arr(limbs_of(a)), arr(limbs_of(b)), mode, out, C.byref(mc))
            v = val(out)
            assert v % p == a * b * rinv % p and v < a * b // (1 << 261) + p + 1, (mode, a, b)
            if mode & 2:
                assert all(l < (1 << 30) + (1 << 6) for l in list(out)[:8])
                # a digit-split result is a valid operand of the next product
                out2 = (C.c_uint32 * 9)()
                lib.l29_mont_ps(C.byref(d), out, arr(limbs_of(b)), mode, out2, C.byref(mc))
                assert val(out2) % p == v * b * rinv % p and mc.value < 0.9 * 2**64
            else:
                assert all(l <= M29 for l in list(out)[:8])
