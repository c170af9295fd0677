"""GPU tests of scb_poly_product_from_host (upload_engine.inc): ProductMLE<K> built straight from host tables.  For the
reference's small-prime fields the tables cross PCIe as packed uint32 (host threads narrow chunks into pinned staging;
a second lane copies chunks as they are and narrows them on the device).  Whatever the lane, the handle must hold the
same field elements and prove to the same transcript bytes as the plain 8-byte upload and the oracle."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import pyoracle as O
from oracle.coracle import CField

import thaler_study_b200 as T
from thaler_study_b200 import _lib
from thaler_study_b200._lib import check, lib

pytestmark = pytest.mark.gpu


def _stats():
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    check(lib.scb_host_pack_stats(C.byref(a), C.byref(b), C.byref(c)))
    _stats.h2d_bytes = c.value
    return a.value, b.value


def _small(threads=3, raw=1, chunk_log2=6, wire=21, nt=1):
    T.set_option("host_pack_min_vars", 6)
    T.set_option("host_pack_wire", wire)  # 21: three entries per word when p < 2^21; 32: uint32
    T.set_option("host_pack_nt", nt)
    T.set_option("host_pack_chunk_log2", chunk_log2)
    T.set_option("host_pack_threads", threads)
    T.set_option("host_pack_raw", 2 if raw else 0)  # 2: device lane even from pageable (numpy) memory


@pytest.mark.parametrize("p", [5, 389, 1572869])
@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_packed_upload_matches_oracle_transcript(p, K, monkeypatch):
    if K >= p:
        pytest.skip("degree >= characteristic")
    OF, F = O.Field(p), T.Field(p)
    rnd = random.Random(1000 * K + p)
    # chunk sizes 2^6 and 2^7: a chunk's last 21-bit word holds one entry (64 = 3*21 + 1) or two (128 = 3*42 + 2)
    for v, threads, raw, wire, nt, cl in ((6, 1, 1, 21, 1, 6), (7, 2, 0, 21, 0, 6), (7, 2, 0, 32, 1, 6), (9, 3, 1, 32, 0, 6), (12, 4, 1, 21, 1, 6),
                                          (12, 4, 0, 21, 1, 6), (9, 3, 1, 21, 0, 7), (11, 2, 0, 21, 1, 7)):
        _small(threads, raw, chunk_log2=cl, wire=wire, nt=nt)
        vals = [[rnd.randrange(p) for _ in range(1 << v)] for _ in range(K)]
        g = T.ProductMLE.from_host_tables(F, v, vals)
        packed_chunks, raw_chunks = _stats()
        assert packed_chunks + raw_chunks == K * (1 << (v - cl)) and (raw or raw_chunks == 0)
        per_chunk = -(-(1 << cl) // 3) * 8 if wire == 21 else (1 << cl) * 4  # ceil(chunk / 3) words, or chunk uint32
        assert _stats.h2d_bytes == packed_chunks * per_chunk + raw_chunks * (1 << cl) * 8
        assert g.num_vars() == v
        for k in range(K):
            assert g.table(k).to_evaluations() == vals[k]
        og = O.ProductMLE(OF, [O.DenseMLE(OF, v, t) for t in vals])
        assert g.sum() == O.Prover(og).c_1()
        want = O.generate_transcript(OF, O.Prover(og))
        assert T.generate_transcript(T.Prover(g)) == want
        plain = T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, v, t) for t in vals])
        assert T.generate_transcript(T.Prover(plain)) == want
        assert T.verify_transcript(want, T.Verifier(v, g))
        # interactive rounds on the packed handle
        pr, r = T.Prover(g), 1
        opr = O.Prover(og)
        for j in range(v):
            assert list(pr.round(r, j).coeffs) == [(int(d), int(c)) for d, c in opr.round(r, j).coeffs]
            r = rnd.randrange(p)


def test_packed_upload_rejects_non_canonical_entries(monkeypatch):
    _small()
    F = T.Field(1572869)
    v = 10
    for lane_raw, wire in ((0, 21), (1, 21), (0, 32), (1, 32)):
        T.set_option("host_pack_raw", 2 if lane_raw else 0)
        T.set_option("host_pack_wire", wire)
        for where in (0, 517, (1 << v) - 1):
            # a real comparison with p: 2^bits(p), p itself, the largest 21-bit value and a 64-bit value all fail
            for bad in (1 << 21, 1572869, (1 << 21) - 1, (1 << 63) + 5):
                t = np.zeros([1 << v, 1], dtype=np.uint64)
                t[where, 0] = bad
                with pytest.raises(T.ScbError) as ei:
                    T.ProductMLE.from_host_tables(F, v, [t, np.ones([1 << v, 1], dtype=np.uint64)])
                assert ei.value.code == _lib.SCB_EINVAL
            t = np.zeros([1 << v, 1], dtype=np.uint64)
            t[where, 0] = 1572868  # p - 1 is fine
            T.ProductMLE.from_host_tables(F, v, [t, np.ones([1 << v, 1], dtype=np.uint64)])
    # the plain 8-byte uploads check the same precondition on the device
    T.set_option("host_pack", 0)
    for Fp, n, bad in ((F, 1, 1572869), (T.Field((1 << 61) - 1), 1, (1 << 61) - 1), (T.Field(O.BLS12_381_FR.p), 4, O.BLS12_381_FR.p)):
        t = np.zeros([1 << v, n], dtype=np.uint64)
        for l in range(n):
            t[33, l] = (bad >> (64 * l)) & 0xFFFFFFFFFFFFFFFF
        with pytest.raises(T.ScbError) as ei:
            T.DenseMultilinearExtension.from_evaluations_vec(Fp, v, t)
        assert ei.value.code == _lib.SCB_EINVAL
        with pytest.raises(T.ScbError) as ei:
            T.ProductMLE.from_host_tables(Fp, v, [t, t])
        assert ei.value.code == _lib.SCB_EINVAL
        t[33, 0] -= 1
        T.DenseMultilinearExtension.from_evaluations_vec(Fp, v, t)


@pytest.mark.parametrize("p", [(1 << 61) - 1, O.BLS12_381_FR.p], ids=["p61", "bls12_381_fr"])
def test_other_fields_take_the_plain_copy(p, monkeypatch):
    _small()
    OF, F = O.Field(p), T.Field(p)
    rnd = random.Random(9)
    v, K = 8, 2
    vals = [[rnd.randrange(p) for _ in range(1 << v)] for _ in range(K)]
    g = T.ProductMLE.from_host_tables(F, v, vals)
    for k in range(K):
        assert g.table(k).to_evaluations() == vals[k]
    want = O.generate_transcript(OF, O.Prover(O.ProductMLE(OF, [O.DenseMLE(OF, v, t) for t in vals])))
    assert T.generate_transcript(T.Prover(g)) == want


@pytest.mark.parametrize("switch", ["default", "no_raw_lane", "wire32", "off"])
def test_large_tables_default_switches(switch, monkeypatch):
    """2^24-entry tables with the default chunking (2^20 entries): both lanes run; the proof equals the plain upload's,
    which the C oracle anchors through the round sums of the first rounds."""
    T.set_option("host_pack_raw", 2)  # numpy tables are pageable: ask for the device lane explicitly
    if switch == "no_raw_lane":
        T.set_option("host_pack_raw", 0)
    if switch == "wire32":
        T.set_option("host_pack_wire", 32)
    if switch == "off":
        T.set_option("host_pack", 0)
    p, v, K = 1572869, 24, 3
    F, cf = T.Field(p), CField(p)
    tabs = [cf.synth(500 + k, 0, 1 << v) for k in range(K)]  # Montgomery words, ark's layout
    g = T.ProductMLE.from_host_tables(F, v, tabs)
    if switch != "off":
        packed_chunks, raw_chunks = _stats()
        assert packed_chunks + raw_chunks == K * (1 << (v - 20))
        assert switch != "no_raw_lane" or raw_chunks == 0
    assert g.round_evals() == cf.from_mont(cf.product_round_evals(tabs, K + 1))
    for k in range(K):
        assert np.array_equal(g.table(k).to_evaluations_mont().reshape(-1), np.asarray(tabs[k]).reshape(-1))
    plain = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 500 + k) for k in range(K)])
    got = T.generate_transcript(T.Prover(g))
    assert got == T.generate_transcript(T.Prover(plain))
    assert T.verify_transcript(got, T.Verifier(v, plain))


@pytest.mark.parametrize("p", [389, 1572869])
def test_mle_from_host_narrows_large_tables_on_the_way(p, monkeypatch):
    """scb_mle_from_host / vsbw_multilinear_from_evaluations with 2^22-entry host tables (pageable numpy memory: host
    lane only) take the packed upload and widen on the device: same table, same evaluation as the device-made one."""
    F, cf = T.Field(p), CField(p)
    v = 22
    tab = cf.synth(77, 0, 1 << v)
    rnd = random.Random(p)
    r = [rnd.randrange(p) for _ in range(v)]
    dev = T.DenseMultilinearExtension.synthetic(F, v, 77)
    want = dev.evaluate_be(r)
    outs = []
    for switch in ("1", "0"):
        T.set_option("host_pack", int(switch))
        m = T.DenseMultilinearExtension.from_evaluations_vec(F, v, tab)
        if switch == "1":
            packed_chunks, raw_chunks = _stats()
            assert (packed_chunks, raw_chunks) == (4, 0)  # 2^22 entries in 2^20-entry chunks, nothing through the device lane
        assert np.array_equal(m.to_evaluations_mont().reshape(-1), np.asarray(tab).reshape(-1))
        assert m.evaluate_be(r) == want
        outs.append(T.vsbw_multilinear_from_evaluations(F, tab, r))
        assert T.cti_multilinear_from_evaluations(F, tab, r) == want
    assert outs == [want, want]
    bad = np.array(tab, copy=True)
    bad[12345] = 1 << 40
    T.set_option("host_pack", 1)
    with pytest.raises(T.ScbError) as ei:
        T.DenseMultilinearExtension.from_evaluations_vec(F, v, bad)
    assert ei.value.code == _lib.SCB_EINVAL
