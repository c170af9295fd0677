"""CPU-only checks of the C ABI: the library loads, exports every symbol the header declares, and its HOST
protocol layer (field conversion, interpolation to SparsePolynomial, serialization, hash-to-field) agrees with
the Python oracle.  No device compute is called here."""
import os
import random
import re

import pytest

from oracle import pyoracle as O

import thaler_study_b200 as T
from thaler_study_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = [O.FP5, O.FP389, O.FP1572869, O.Field((1 << 61) - 1), O.Field(0xFFFFFFFF00000001), O.BLS12_381_FR]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sumcheck_b200.h")).read()
    declared = set(re.findall(r"\b(scb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert b"sm_100a" in _lib.lib.scb_version()


@pytest.mark.parametrize("k,nv,chunk_log2,workers,raw_lane", [(1, 6, 6, 1, 0), (3, 12, 6, 4, 1), (4, 14, 8, 7, 1), (2, 16, 10, 3, 0),
                                                              (3, 20, 16, 8, 1), (1, 10, 10, 2, 1), (2, 13, 7, 3, 1), (3, 15, 9, 5, 0)])
def test_packed_upload_scheduler_on_a_memcpy_backend(k, nv, chunk_log2, workers, raw_lane):
    """host/hostpack.hpp: pack workers take chunks from the front, the raw lane from the back; every entry must arrive
    narrowed at its place exactly once (no device involved: the back end copies into host memory)."""
    for seed in (1, 2, 3, 5, 6, 7):  # bit 0: streaming stores, bit 1: 21-bit wire format, bit 2: software prefetch
        _lib.check(_lib.lib.scb_host_pack_selftest(k, nv, chunk_log2, workers, raw_lane, seed))


def test_no_cpu_fallback_without_device():
    if T.device_count() > 0:
        pytest.skip("a device is present")
    F = T.Field(1572869)
    with pytest.raises(T.ScbError) as ei:
        T.DenseMultilinearExtension.from_evaluations_vec(F, 2, [1, 2, 3, 4])
    assert ei.value.code == _lib.SCB_ECUDA


def test_field_create_rejects_bad_moduli():
    for bad in (4, 0, 1):
        with pytest.raises(T.ScbError):
            T.Field(bad)


@pytest.mark.parametrize("OF", FIELDS, ids=lambda F: f"p{F.bits}")
def test_montgomery_representation_is_arks(OF):
    F = T.Field(OF.p)
    rnd = random.Random(3)
    vals = [0, 1, OF.p - 1] + [rnd.randrange(OF.p) for _ in range(50)]
    m = F.to_mont(vals)
    assert F.unpack_raw(m) == [OF.to_mont(v) for v in vals]
    assert F.from_mont(m) == vals
    assert F.policy == (4 if OF.n_limbs == 4 else (0 if OF.bits <= 28 else 1))


@pytest.mark.parametrize("OF", FIELDS, ids=lambda F: f"p{F.bits}")
def test_hash_to_field_matches_oracle(OF):
    F = T.Field(OF.p)
    rnd = random.Random(5)
    for ln in (0, 1, 17, 31, 32, 33, 55, 56, 63, 64, 65, 200, 1000):
        msg = bytes(rnd.randrange(256) for _ in range(ln))
        assert F.hash_to_field(msg) == O.hash_to_field(OF, msg), ln


@pytest.mark.parametrize("OF", FIELDS, ids=lambda F: f"p{F.bits}")
def test_evals_to_univariate_and_serialization(OF):
    F = T.Field(OF.p)
    rnd = random.Random(7)
    cases3 = [[0, 0, 0], [0, 3 % OF.p, 0], [0, 0, 1], [1, 0, 0], [2 % OF.p, 2 % OF.p, 2 % OF.p], [0, 1, 2 % OF.p]]
    cases3 += [[rnd.randrange(OF.p) for _ in range(3)] for _ in range(40)]
    if OF.p == 5:
        cases3 = [[a, b, c] for a in range(5) for b in range(5) for c in range(5)]
    for ev in cases3:
        # matrix_multiplication::G -- interpolate_quadratic_poly incl. explicit zero terms
        want = O.interpolate_quadratic_poly(OF, list(zip([0, 1, 2 % OF.p], ev)))
        got = T.evals_to_univariate(F, T.KIND_MATMUL_G, ev)
        assert got.coeffs == want.coeffs, ev
        assert got.serialize_uncompressed() == want.serialize()
        # triangle / GKR / ProductMLE<2>: Dense -> Sparse
        want2 = O.SparsePoly.from_dense(OF, O.lagrange_to_coeffs(OF, ev))
        for kind in (T.KIND_PRODUCT, T.KIND_TRIANGLE_G, T.KIND_GKR_W):
            got2 = T.evals_to_univariate(F, kind, ev)
            assert got2.coeffs == want2.coeffs, ev
            assert got2.serialize_uncompressed() == want2.serialize()
        x = rnd.randrange(OF.p)
        assert got.evaluate(x) == want.evaluate(x)
    for npts in (2, 4, 5):
        if npts > OF.p:
            continue
        for _ in range(20):
            ev = [rnd.randrange(OF.p) for _ in range(npts)]
            want = O.SparsePoly.from_dense(OF, O.lagrange_to_coeffs(OF, ev))
            got = T.evals_to_univariate(F, T.KIND_PRODUCT, ev)
            assert got.coeffs == want.coeffs
            assert [got.evaluate(i) for i in range(npts)] == ev


def test_domain4_interpolation_equals_lagrange_on_0_1_2():
    # the engine samples X = 0,1,2 where the reference samples the 4th roots of unity
    # (triangle-counting/src/lib.rs:121-131): same polynomial, same coefficients.
    rnd = random.Random(9)
    for OF in (O.FP5, O.FP389, O.FP1572869):
        for _ in range(30):
            coeffs = [rnd.randrange(OF.p) for _ in range(3)]
            f = lambda x: sum(c * pow(x, i, OF.p) for i, c in enumerate(coeffs)) % OF.p
            dom = O.interpolate_domain4(OF, [f(e) for e in O.domain4_elements(OF)])
            lag = O.lagrange_to_coeffs(OF, [f(0), f(1), f(2)])
            assert dom == lag


def test_sha256_scalar_and_sha_extension_paths_agree():
    """host/fiat_shamir.hpp picks the x86 SHA-extension compression at run time; SCB_SHA_SCALAR=1 forces the portable
    one.  Both must give the same hash_to_field values (and the in-process values are pinned to the oracle above)."""
    import os
    import subprocess
    import sys

    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import ctypes as C, numpy as np\n"
        "import thaler_study_b200 as T; T.options_from_env()\n"
        "from thaler_study_b200._lib import lib\n"
        "for p in (5, 389, 1572869, 0xFFFFFFFF00000001, %d):\n"
        "    F = T.Field(p)\n"
        "    for n in (0, 1, 19, 55, 56, 63, 64, 65, 119, 120, 300, 1000):\n"
        "        msg = bytes((i * 31 + n) & 255 for i in range(n))\n"
        "        print(F.hash_to_field(msg))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), O.BLS12_381_FR.p)
    outs = [subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SCB_SHA_SCALAR=v), capture_output=True, text=True, timeout=300)
            for v in ("0", "1")]
    for o in outs:
        assert o.returncode == 0, o.stderr[-2000:]
    assert outs[0].stdout == outs[1].stdout and len(outs[0].stdout.split()) == 60
    # and against the oracle (hashlib)
    vals = outs[0].stdout.split()
    i = 0
    for p in (5, 389, 1572869, 0xFFFFFFFF00000001, O.BLS12_381_FR.p):
        OF = O.Field(p)
        for n in (0, 1, 19, 55, 56, 63, 64, 65, 119, 120, 300, 1000):
            msg = bytes((j * 31 + n) & 255 for j in range(n))
            assert int(vals[i]) == O.hash_to_field(OF, msg), (p, n)
            i += 1


# ----------------------------------------------------------------------------- options (no environment reads)
def test_options_are_set_through_the_abi_only():
    names = T.option_names()
    assert "pairs" in names and "strict_verifier" in names and len(names) == len(set(names))
    assert T.get_option("tail_vars") == 14
    T.set_option("tail_vars", 5)
    assert T.get_option("tail_vars") == 5
    T.reset_options()
    assert T.get_option("tail_vars") == 14
    with pytest.raises(T.ScbError) as ei:
        T.set_option("no_such_switch", 1)
    assert ei.value.code == _lib.SCB_EINVAL
    # the harness helper maps SCB_<NAME> variables to scb_set_option calls; the library itself has no getenv
    assert T.options_from_env({"SCB_TAIL_VARS": "7", "SCB_PAIRS": "0", "SCB_UNRELATED": "1", "HOME": "/"}) == {"tail_vars": 7, "pairs": 0}
    assert T.get_option("tail_vars") == 7 and T.get_option("pairs") == 0
    T.reset_options()
    src = os.path.join(ROOT, "thaler_study_b200", "csrc")
    for dirpath, _, files in os.walk(src):
        for fn in files:
            if fn.endswith((".cu", ".cuh", ".cpp", ".hpp", ".inc")):
                assert "getenv" not in open(os.path.join(dirpath, fn)).read(), fn


# ----------------------------------------------------------------------------- verifier hardening (host-only parts)
def _final_round_without_link(strict):
    """n = 2, no oracle: g_1 consistent with a false c_1, g_2 NOT linked to g_1(r_1).  The reference's last-round
    branch (sum-check-protocol/src/lib.rs:298-310) goes straight to the oracle; the strict verifier rejects first."""
    F = T.Field(1572869)
    T.set_option("strict_verifier", 1 if strict else 0)
    v = T.Verifier(2, None, F)
    v.set_c_1(7)                                      # g_1(0) + g_1(1) = 3 + 4

    class R:
        def __init__(self, vals):
            self.vals = list(vals)

        def draw(self):
            return self.vals.pop(0)

    rng = R([5, 6])
    assert v.round(T.SparsePolynomial(F, [(0, 3), (1, 1)]), rng) == ("JthRound", 5)   # g_1 = 3 + X, g_1(5) = 8
    return v.round(T.SparsePolynomial(F, [(0, 1), (1, 1)]), rng)                      # g_2(0) + g_2(1) = 3 != 8


def test_strict_verifier_checks_the_link_in_the_final_round():
    with pytest.raises(T.ProverClaimMismatch):
        _final_round_without_link(strict=True)
    with pytest.raises(T.NoPolySet):   # the reference's literal logic gets as far as the oracle call
        _final_round_without_link(strict=False)


def test_verify_transcript_validates_offsets():
    import ctypes as C

    import numpy as np

    F = T.Field(389)
    v = T.Verifier(3, None, F)
    raw = bytes(40)
    buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
    acc = C.c_int(1)
    for offs in ([0, 50], [1, 10], [0, 20, 10]):
        o = np.array(offs, dtype=np.uint64)
        rc = _lib.lib.scb_fs_verify_transcript(v._h, buf, len(raw), o.ctypes.data_as(_lib.u64p), len(offs) - 1, C.byref(acc))
        assert rc == _lib.SCB_EINVAL and acc.value == 0, offs
