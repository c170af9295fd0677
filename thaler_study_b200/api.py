"""Thin Python mirror of the reference's types over the C ABI (ctypes), for tests and benches.

Names follow /root/reference: ``DenseMultilinearExtension`` ([ARK]), ``SumCheckPolynomial`` implementors
``ProductMLE`` / ``MatMulG`` (matrix_multiplication::G) / ``TriangleG`` (triangle_counting::G) / ``GkrW``
(gkr_protocol::round_polynomial::W), ``Prover`` / ``Verifier`` (sum_check_protocol), ``generate_transcript`` /
``verify_transcript`` (fiat_shamir) and the two free functions of multilinear_extensions.

Field elements cross this layer as canonical Python ints (like ``Fp::from_bigint`` / ``into_bigint``); bulk
tables can also be given as numpy ``uint64[count, n_limbs]`` arrays of Montgomery limbs (ark's memory format).
All arithmetic happens in libsumcheck_b200.so: the device kernels and the C++ host protocol layer.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from ._lib import SCB_OK, ScbError, check, lib, u8p, u64p

MAX_TERMS = 64


def _p64(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


class Field:
    """A prime field; mirrors ``#[derive(MontConfig)] #[modulus = "..."]`` + ``Fp64<MontBackend<_,1>>``."""

    def __init__(self, modulus: int):
        self.p = int(modulus)
        self.n = 1 if self.p.bit_length() <= 64 else 4
        if self.p.bit_length() > 256:
            raise ValueError("modulus wider than 4 limbs")
        limbs = (C.c_uint64 * 4)(*[(self.p >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)])
        self._h = C.c_void_p()
        check(lib.scb_field_create(self.n, limbs, C.byref(self._h)))
        bits = C.c_uint32()
        check(lib.scb_field_modulus_bits(self._h, C.byref(bits)))
        self.bits = bits.value
        pol = C.c_uint32()
        check(lib.scb_field_policy(self._h, C.byref(pol)))
        self.policy = pol.value

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.scb_field_free(self._h)
            self._h = None

    @property
    def ser_bytes(self) -> int:
        return (self.bits + 7) // 8

    # canonical ints <-> Montgomery limb arrays
    def pack_raw(self, vals: Sequence[int]) -> np.ndarray:
        a = np.zeros((len(vals), self.n), dtype=np.uint64)
        if self.n == 1:
            a[:, 0] = np.array([int(v) for v in vals], dtype=np.uint64) if len(vals) else a[:, 0]
        else:
            for i, v in enumerate(vals):
                v = int(v)
                for l in range(self.n):
                    a[i, l] = (v >> (64 * l)) & 0xFFFFFFFFFFFFFFFF
        return a

    def unpack_raw(self, a: np.ndarray) -> List[int]:
        a = a.reshape(-1, self.n)
        if self.n == 1:
            return [int(x) for x in a[:, 0].tolist()]
        return [sum(int(x) << (64 * l) for l, x in enumerate(row)) for row in a.tolist()]

    def to_mont(self, vals: Sequence[int]) -> np.ndarray:
        raw = self.pack_raw([int(v) % self.p for v in vals])
        out = np.empty_like(raw)
        check(lib.scb_field_to_mont(self._h, _p64(raw), _p64(out), len(vals)))
        return out

    def from_mont(self, a: np.ndarray) -> List[int]:
        a = np.ascontiguousarray(a.reshape(-1, self.n))
        out = np.empty_like(a)
        check(lib.scb_field_from_mont(self._h, _p64(a), _p64(out), a.shape[0]))
        return self.unpack_raw(out)

    def elem(self, v: int) -> np.ndarray:
        return self.to_mont([v])

    def hash_to_field(self, msg: bytes) -> int:
        out = np.zeros((1, self.n), dtype=np.uint64)
        buf = (C.c_uint8 * max(len(msg), 1)).from_buffer_copy(msg.ljust(1, b"\0"))
        check(lib.scb_hash_to_field(self._h, buf, len(msg), _p64(out)))
        return self.from_mont(out)[0]


Table = Union[Sequence[int], np.ndarray]


def _as_mont(F: Field, evals: Table) -> np.ndarray:
    if isinstance(evals, np.ndarray) and evals.dtype == np.uint64:
        return np.ascontiguousarray(evals.reshape(-1, F.n))
    return F.to_mont(list(evals))


class DenseMultilinearExtension:
    """[ARK] ark_poly::DenseMultilinearExtension<F> with the table resident in HBM."""

    def __init__(self, F: Field, handle: C.c_void_p):
        self.F = F
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.scb_mle_free(self._h)
            self._h = None

    @staticmethod
    def from_evaluations_vec(F: Field, num_vars: int, evals: Table) -> "DenseMultilinearExtension":
        m = _as_mont(F, evals)
        if m.shape[0] != 1 << num_vars:
            raise ValueError("The size of evaluations should be 2^num_vars.")
        h = C.c_void_p()
        check(lib.scb_mle_from_host(F._h, num_vars, _p64(m), C.byref(h)))
        return DenseMultilinearExtension(F, h)

    from_evaluations_slice = from_evaluations_vec

    @staticmethod
    def synthetic(F: Field, num_vars: int, seed: int, start: int = 0) -> "DenseMultilinearExtension":
        h = C.c_void_p()
        check(lib.scb_mle_synthetic(F._h, num_vars, seed, start, C.byref(h)))
        return DenseMultilinearExtension(F, h)

    @staticmethod
    def from_device(F: Field, num_vars: int, data_ptr: int, copy: bool = False) -> "DenseMultilinearExtension":
        h = C.c_void_p()
        check(lib.scb_mle_from_device(F._h, num_vars, C.c_void_p(data_ptr), 1 if copy else 0, C.byref(h)))
        return DenseMultilinearExtension(F, h)

    def clone(self) -> "DenseMultilinearExtension":
        h = C.c_void_p()
        check(lib.scb_mle_clone(self._h, C.byref(h)))
        return DenseMultilinearExtension(self.F, h)

    @property
    def num_vars(self) -> int:
        o = C.c_uint32()
        check(lib.scb_mle_num_vars(self._h, C.byref(o)))
        return o.value

    @property
    def device_ptr(self) -> int:
        o = C.c_void_p()
        check(lib.scb_mle_device_ptr(self._h, C.byref(o)))
        return o.value or 0

    def fix_variables(self, partial_point: Sequence[int]) -> "DenseMultilinearExtension":
        pt = self.F.to_mont(list(partial_point)) if len(partial_point) else np.zeros((1, self.F.n), dtype=np.uint64)
        h = C.c_void_p()
        check(lib.scb_mle_fix_variables(self._h, _p64(pt), len(partial_point), C.byref(h)))
        return DenseMultilinearExtension(self.F, h)

    def evaluate(self, point: Sequence[int]) -> int:
        pt = self.F.to_mont(list(point)) if len(point) else np.zeros((1, self.F.n), dtype=np.uint64)
        out = np.zeros((1, self.F.n), dtype=np.uint64)
        check(lib.scb_mle_evaluate(self._h, _p64(pt), len(point), _p64(out)))
        return self.F.from_mont(out)[0]

    def evaluate_many(self, points: Sequence[Sequence[int]]) -> List[int]:
        """Several LSB-first evaluations of this table in one call (scb_mle_evaluate_many): the GKR prover's
        restrict_poly evaluates W~ at k + 1 points of a line (gkr-protocol/src/lib.rs:291-321)."""
        flat = [x for pt in points for x in pt]
        pts = self.F.to_mont(flat)
        out = np.zeros((len(points), self.F.n), dtype=np.uint64)
        check(lib.scb_mle_evaluate_many(self._h, _p64(pts), len(points[0]), len(points), _p64(out)))
        return self.F.from_mont(out)

    def evaluate_be(self, r: Sequence[int]) -> int:
        pt = self.F.to_mont(list(r)) if len(r) else np.zeros((1, self.F.n), dtype=np.uint64)
        out = np.zeros((1, self.F.n), dtype=np.uint64)
        check(lib.scb_mle_evaluate_be(self._h, _p64(pt), len(r), _p64(out)))
        return self.F.from_mont(out)[0]

    def relabel(self, a: int, b: int, k: int) -> "DenseMultilinearExtension":
        h = C.c_void_p()
        check(lib.scb_mle_relabel(self._h, a, b, k, C.byref(h)))
        return DenseMultilinearExtension(self.F, h)

    def to_evaluations_mont(self) -> np.ndarray:
        n = 1 << self.num_vars
        out = np.empty((n, self.F.n), dtype=np.uint64)
        check(lib.scb_mle_to_evaluations(self._h, _p64(out), n))
        return out

    def to_evaluations(self) -> List[int]:
        return self.F.from_mont(self.to_evaluations_mont())


# ---------------------------------------------------------------- multilinear-extensions free functions
def vsbw_multilinear_from_evaluations(F: Field, evals: Table, r: Sequence[int]) -> int:
    """multilinear-extensions/src/lib.rs:6-24."""
    m = _as_mont(F, evals)
    pt = F.to_mont(list(r)) if len(r) else np.zeros((1, F.n), dtype=np.uint64)
    out = np.zeros((1, F.n), dtype=np.uint64)
    check(lib.scb_vsbw_multilinear_from_evaluations(F._h, _p64(m), m.shape[0], _p64(pt), len(r), _p64(out)))
    return F.from_mont(out)[0]


def cti_multilinear_from_evaluations(F: Field, evals: Table, r: Sequence[int]) -> int:
    """multilinear-extensions/src/lib.rs:29-48."""
    m = _as_mont(F, evals)
    pt = F.to_mont(list(r)) if len(r) else np.zeros((1, F.n), dtype=np.uint64)
    out = np.zeros((1, F.n), dtype=np.uint64)
    check(lib.scb_cti_multilinear_from_evaluations(F._h, _p64(m), m.shape[0], _p64(pt), len(r), _p64(out)))
    return F.from_mont(out)[0]


# ---------------------------------------------------------------- univariate::SparsePolynomial
class SparsePolynomial:
    """[ARK] univariate::SparsePolynomial<F> as a list of (degree, canonical coefficient)."""

    def __init__(self, F: Field, coeffs: Sequence[Tuple[int, int]]):
        self.F = F
        self.coeffs = [(int(d), int(c)) for d, c in coeffs]

    def _arrays(self):
        n = len(self.coeffs)
        deg = np.array([d for d, _ in self.coeffs] + [0], dtype=np.uint64)
        co = self.F.to_mont([c for _, c in self.coeffs] + [0])
        return n, deg, co

    def evaluate(self, x: int) -> int:
        n, deg, co = self._arrays()
        out = np.zeros((1, self.F.n), dtype=np.uint64)
        check(lib.scb_unipoly_evaluate(self.F._h, _p64(deg), _p64(co), n, _p64(self.F.elem(x)), _p64(out)))
        return self.F.from_mont(out)[0]

    def serialize_uncompressed(self) -> bytes:
        n, deg, co = self._arrays()
        cap = 8 + n * (8 + self.F.ser_bytes) + 8
        buf = (C.c_uint8 * cap)()
        ln = C.c_size_t()
        check(lib.scb_unipoly_serialize(self.F._h, _p64(deg), _p64(co), n, buf, cap, C.byref(ln)))
        return bytes(buf[: ln.value])

    def __eq__(self, o):
        return isinstance(o, SparsePolynomial) and self.coeffs == o.coeffs

    def __repr__(self):
        return f"SparsePolynomial({self.coeffs})"


def _terms_out(F: Field):
    deg = np.zeros(MAX_TERMS, dtype=np.uint64)
    co = np.zeros((MAX_TERMS, F.n), dtype=np.uint64)
    return deg, co, C.c_uint32()


def _poly_from_out(F: Field, deg, co, n) -> SparsePolynomial:
    k = n.value
    return SparsePolynomial(F, list(zip([int(d) for d in deg[:k]], F.from_mont(co[:k]) if k else [])))


def evals_to_univariate(F: Field, kind: int, evals: Sequence[int]) -> SparsePolynomial:
    deg, co, n = _terms_out(F)
    ev = F.to_mont(list(evals))
    check(lib.scb_evals_to_univariate(F._h, kind, _p64(ev), len(evals), _p64(deg), _p64(co), MAX_TERMS, C.byref(n)))
    return _poly_from_out(F, deg, co, n)


def evals_to_univariate_mont(F: Field, kind: int, ev_mont: np.ndarray) -> SparsePolynomial:
    """Same as evals_to_univariate with the sums given as Montgomery limbs uint64[n_points, n_limbs]."""
    deg, co, n = _terms_out(F)
    ev = np.ascontiguousarray(ev_mont.reshape(-1, F.n))
    check(lib.scb_evals_to_univariate(F._h, kind, _p64(ev), ev.shape[0], _p64(deg), _p64(co), MAX_TERMS, C.byref(n)))
    return _poly_from_out(F, deg, co, n)


# ---------------------------------------------------------------- SumCheckPolynomial implementors
KIND_PRODUCT, KIND_MATMUL_G, KIND_TRIANGLE_G, KIND_GKR_W = 0, 1, 2, 3


class SumCheckPolynomial:
    """trait SumCheckPolynomial<F> (sum-check-protocol/src/lib.rs:121-156) over a device handle."""

    def __init__(self, F: Field, handle: C.c_void_p):
        self.F = F
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.scb_poly_free(self._h)
            self._h = None

    def _wrap(self, h) -> "SumCheckPolynomial":
        return type(self)(self.F, h)

    def clone(self):
        h = C.c_void_p()
        check(lib.scb_poly_clone(self._h, C.byref(h)))
        return self._wrap(h)

    @property
    def kind(self) -> int:
        o = C.c_uint32()
        check(lib.scb_poly_kind_of(self._h, C.byref(o)))
        return o.value

    @property
    def n_points(self) -> int:
        o = C.c_uint32()
        check(lib.scb_poly_n_points(self._h, C.byref(o)))
        return o.value

    def allow_packed(self, enable: bool = True) -> "SumCheckPolynomial":
        """Let polynomials derived from this handle keep packed uint32 folded tables (scb_poly_allow_packed)."""
        check(lib.scb_poly_allow_packed(self._h, 1 if enable else 0))
        return self

    def table(self, idx: int) -> DenseMultilinearExtension:
        h = C.c_void_p()
        check(lib.scb_poly_table(self._h, idx, C.byref(h)))
        return DenseMultilinearExtension(self.F, h)

    # ---- the five trait methods
    def evaluate(self, point: Sequence[int]) -> Optional[int]:
        pt = self.F.to_mont(list(point)) if len(point) else np.zeros((1, self.F.n), dtype=np.uint64)
        out = np.zeros((1, self.F.n), dtype=np.uint64)
        rc = lib.scb_poly_evaluate(self._h, _p64(pt), len(point), _p64(out))
        if rc == -1:  # SCB_EINVAL -> None (dimension mismatch, :124-126)
            return None
        check(rc)
        return self.F.from_mont(out)[0]

    def fix_variables(self, partial_point: Sequence[int]):
        pt = self.F.to_mont(list(partial_point)) if len(partial_point) else np.zeros((1, self.F.n), dtype=np.uint64)
        h = C.c_void_p()
        check(lib.scb_poly_fix_variables(self._h, _p64(pt), len(partial_point), C.byref(h)))
        return self._wrap(h)

    def to_univariate(self) -> SparsePolynomial:
        deg, co, n = _terms_out(self.F)
        check(lib.scb_poly_to_univariate(self._h, _p64(deg), _p64(co), MAX_TERMS, C.byref(n)))
        return _poly_from_out(self.F, deg, co, n)

    def num_vars(self) -> int:
        o = C.c_uint32()
        check(lib.scb_poly_num_vars(self._h, C.byref(o)))
        return o.value

    def to_evaluations(self) -> List[int]:
        n = 1 << self.num_vars()
        out = np.empty((n, self.F.n), dtype=np.uint64)
        check(lib.scb_poly_to_evaluations(self._h, _p64(out), n))
        return self.F.from_mont(out)

    # ---- device halves
    def sum(self) -> int:
        out = np.zeros((1, self.F.n), dtype=np.uint64)
        check(lib.scb_poly_sum(self._h, _p64(out)))
        return self.F.from_mont(out)[0]

    def round_evals(self, n_points: Optional[int] = None) -> List[int]:
        npts = self.n_points if n_points is None else n_points
        out = np.zeros((max(npts, 1), self.F.n), dtype=np.uint64)
        check(lib.scb_poly_round_evals(self._h, npts, _p64(out)))
        return self.F.from_mont(out[:npts])

    def fix_and_round_evals(self, r: int, n_points: Optional[int] = None, claim: Optional[int] = None):
        """Fused `g = g.fix_variables(&[r]); g.to_univariate()` (sums at X = 0..d).  claim = g(0) + g(1) of the message
        about to be computed lets 4-limb fields take the leaner kernel (scb_poly_fix_and_round_evals_claim)."""
        npts = self.n_points if n_points is None else n_points
        out = np.zeros((max(npts, 1), self.F.n), dtype=np.uint64)
        h = C.c_void_p()
        if claim is None:
            check(lib.scb_poly_fix_and_round_evals(self._h, _p64(self.F.elem(r)), npts, C.byref(h), _p64(out)))
        else:
            check(lib.scb_poly_fix_and_round_evals_claim(self._h, _p64(self.F.elem(r)), _p64(self.F.elem(claim)), npts, C.byref(h), _p64(out)))
        return self._wrap(h), self.F.from_mont(out[:npts])

    # ---- two rounds per pass (small-prime fields, product polynomials; csrc/pairs.cuh)
    def grid_evals(self) -> List[List[int]]:
        """H[a][b] = sum over x'' of prod_k f_k(a, b, x'') for a, b in 0..K (canonical integers)."""
        npts = self.n_points
        out = np.zeros((npts * npts, self.F.n), dtype=np.uint64)
        check(lib.scb_poly_grid_evals(self._h, _p64(out)))
        flat = self.F.from_mont(out)
        return [flat[a * npts:(a + 1) * npts] for a in range(npts)]

    def pair_pass(self, ra: int, rb: int):
        """Fold the two lowest variables by (ra, rb); returns (folded polynomial, its grid)."""
        npts = self.n_points
        out = np.zeros((npts * npts, self.F.n), dtype=np.uint64)
        h = C.c_void_p()
        check(lib.scb_poly_pair_pass(self._h, _p64(self.F.elem(ra)), _p64(self.F.elem(rb)), C.byref(h), _p64(out)))
        flat = self.F.from_mont(out)
        return self._wrap(h), [flat[a * npts:(a + 1) * npts] for a in range(npts)]

    def round_evals_device(self, d_out_ptr: int, n_points: Optional[int] = None) -> None:
        npts = self.n_points if n_points is None else n_points
        check(lib.scb_poly_round_evals_device(self._h, npts, C.c_void_p(d_out_ptr)))

    def fix_and_round_evals_device(self, r: int, d_out_ptr: int, n_points: Optional[int] = None):
        npts = self.n_points if n_points is None else n_points
        h = C.c_void_p()
        check(lib.scb_poly_fix_and_round_evals_device(self._h, _p64(self.F.elem(r)), npts, C.byref(h), C.c_void_p(d_out_ptr)))
        return self._wrap(h)


class ProductMLE(SumCheckPolynomial):
    """Product of K dense MLEs over the same variables (the impl BASELINE configs 1 and 5 name)."""

    @staticmethod
    def new(tables: Sequence[DenseMultilinearExtension]) -> "ProductMLE":
        arr = (C.c_void_p * len(tables))(*[t._h for t in tables])
        h = C.c_void_p()
        check(lib.scb_poly_product(arr, len(tables), C.byref(h)))
        return ProductMLE(tables[0].F, h)

    @staticmethod
    def from_host_tables(F: Field, num_vars: int, tables: Sequence[Table]) -> "ProductMLE":
        """K host tables (ark's in-memory format when given as uint64 arrays) -> one handle, in one call: the upload
        packs small-prime tables to 32 bits on the way (scb_poly_product_from_host)."""
        ms = [_as_mont(F, t) for t in tables]
        for m in ms:
            if m.shape[0] != 1 << num_vars:
                raise ValueError("The size of evaluations should be 2^num_vars.")
        arr = (u64p * len(ms))(*[_p64(m) for m in ms])
        h = C.c_void_p()
        check(lib.scb_poly_product_from_host(F._h, len(ms), num_vars, arr, C.byref(h)))
        return ProductMLE(F, h)


class MatMulG(SumCheckPolynomial):
    """matrix_multiplication::G (matrix-multiplication/src/lib.rs:12-15,62-147)."""

    @staticmethod
    def new(F: Field, n: int, a: Table, b: Table, point: Sequence[int]) -> "MatMulG":
        ma, mb = _as_mont(F, a), _as_mont(F, b)
        if ma.shape[0] != 1 << (2 * n) or mb.shape[0] != 1 << (2 * n) or len(point) != 2 * n:
            raise ValueError("G::new: a, b must have 2^(2n) entries and point 2n coordinates")
        h = C.c_void_p()
        check(lib.scb_poly_matmul_g_new(F._h, n, _p64(ma), _p64(mb), _p64(F.to_mont(list(point))), C.byref(h)))
        return MatMulG(F, h)

    @staticmethod
    def from_tables(f_a: DenseMultilinearExtension, f_b: DenseMultilinearExtension) -> "MatMulG":
        h = C.c_void_p()
        check(lib.scb_poly_matmul_g(f_a._h, f_b._h, C.byref(h)))
        return MatMulG(f_a.F, h)


class TriangleG(SumCheckPolynomial):
    """triangle_counting::G (triangle-counting/src/lib.rs:22-27,29-166)."""

    @staticmethod
    def new_adj_matrix(F: Field, num_vars: int, matrix: Sequence[bool]) -> "TriangleG":
        adj = np.ascontiguousarray(np.array([1 if b else 0 for b in matrix], dtype=np.uint8))
        if adj.shape[0] != 1 << num_vars:
            raise ValueError("The size of evaluations should be 2^num_vars.")
        h = C.c_void_p()
        check(lib.scb_poly_triangle_g_new(F._h, num_vars, adj.ctypes.data_as(u8p), C.byref(h)))
        return TriangleG(F, h)


class GkrW(SumCheckPolynomial):
    """gkr_protocol::round_polynomial::W (gkr-protocol/src/round_polynomial.rs:23-119)."""

    @staticmethod
    def new(add_i: DenseMultilinearExtension, mul_i: DenseMultilinearExtension, w_b: DenseMultilinearExtension,
            w_c: DenseMultilinearExtension) -> "GkrW":
        h = C.c_void_p()
        check(lib.scb_poly_gkr_w(add_i._h, mul_i._h, w_b._h, w_c._h, C.byref(h)))
        return GkrW(add_i.F, h)


# ---------------------------------------------------------------- Prover / Verifier
class Prover:
    """sum_check_protocol::Prover<F, P> (sum-check-protocol/src/lib.rs:73-117)."""

    def __init__(self, g: SumCheckPolynomial):
        self.F = g.F
        self._h = C.c_void_p()
        check(lib.scb_prover_new(g._h, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.scb_prover_free(self._h)
            self._h = None

    def c_1(self) -> int:
        out = np.zeros((1, self.F.n), dtype=np.uint64)
        check(lib.scb_prover_c_1(self._h, _p64(out)))
        return self.F.from_mont(out)[0]

    def num_vars(self) -> int:
        o = C.c_uint32()
        check(lib.scb_prover_num_vars(self._h, C.byref(o)))
        return o.value

    def round(self, r_prev: int, j: int) -> SparsePolynomial:
        deg, co, n = _terms_out(self.F)
        check(lib.scb_prover_round(self._h, _p64(self.F.elem(r_prev)), j, _p64(deg), _p64(co), MAX_TERMS, C.byref(n)))
        return _poly_from_out(self.F, deg, co, n)


class Verifier:
    """sum_check_protocol::Verifier<F, P> (sum-check-protocol/src/lib.rs:227-331).

    ``round(g_j, rng)`` returns ("JthRound", r_j) or ("FinalRound", bool); ``rng.draw()`` is RngF::draw.
    """

    def __init__(self, n: int, g: Optional[SumCheckPolynomial], F: Optional[Field] = None):
        self.F = F if F is not None else g.F
        self._h = C.c_void_p()
        check(lib.scb_verifier_new(self.F._h, n, g._h if g is not None else None, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.scb_verifier_free(self._h)
            self._h = None

    def set_c_1(self, c_1: int) -> None:
        check(lib.scb_verifier_set_c_1(self._h, _p64(self.F.elem(c_1))))

    def round(self, g_j: SparsePolynomial, rng):
        r_j = rng.draw()
        n, deg, co = g_j._arrays()
        fin, acc = C.c_int(), C.c_int()
        check(lib.scb_verifier_round(self._h, _p64(deg), _p64(co), n, _p64(self.F.elem(r_j)), C.byref(fin), C.byref(acc)))
        if fin.value:
            return ("FinalRound", bool(acc.value))
        return ("JthRound", r_j)


# ---------------------------------------------------------------- fiat-shamir
def generate_transcript(prover: Prover) -> List[bytes]:
    """fiat_shamir::generate_transcript::<F, Prover<F,P>, DefaultFieldHasher<Sha256>> (fiat-shamir/src/lib.rs:75-98)."""
    nv = prover.num_vars()
    cap = 64 + nv * (8 + MAX_TERMS * (8 + prover.F.ser_bytes)) + prover.F.ser_bytes
    buf = (C.c_uint8 * cap)()
    ln = C.c_size_t()
    offs = np.zeros(nv + 2, dtype=np.uint64)
    check(lib.scb_fs_generate_transcript(prover._h, buf, cap, C.byref(ln), _p64(offs)))
    raw = bytes(buf[: ln.value])
    return [raw[int(offs[i]) : int(offs[i + 1])] for i in range(max(nv, 1))]


def verify_transcript(transcript: Sequence[bytes], verifier: Verifier) -> bool:
    """fiat_shamir::verify_transcript (fiat-shamir/src/lib.rs:123-143)."""
    raw = b"".join(transcript)
    offs = np.zeros(len(transcript) + 1, dtype=np.uint64)
    o = 0
    for i, m in enumerate(transcript):
        o += len(m)
        offs[i + 1] = o
    buf = (C.c_uint8 * max(len(raw), 1)).from_buffer_copy(raw.ljust(1, b"\0"))
    acc = C.c_int()
    check(lib.scb_fs_verify_transcript(verifier._h, buf, len(raw), _p64(offs), len(transcript), C.byref(acc)))
    return bool(acc.value)


class Transcript:
    """The Fiat-Shamir hash chain of fiat-shamir/src/lib.rs:75-98 as a state machine (C++ host code)."""

    def __init__(self, F: Field, kind: int):
        self.F = F
        self._h = C.c_void_p()
        check(lib.scb_transcript_new(F._h, kind, C.byref(self._h)))
        self._r = np.zeros((1, F.n), dtype=np.uint64)

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.scb_transcript_free(self._h)
            self._h = None

    def absorb_round_mont(self, parts: np.ndarray) -> np.ndarray:
        """parts: uint64[n_parts, n_points, n_limbs] Montgomery limbs -> challenge r as Montgomery limbs [1, n]."""
        parts = np.ascontiguousarray(parts, dtype=np.uint64)
        n_parts, n_points = parts.shape[0], parts.shape[1]
        check(lib.scb_transcript_absorb_round(self._h, _p64(parts), n_parts, n_points, _p64(self._r)))
        return self._r

    def c_1(self) -> int:
        out = np.zeros((1, self.F.n), dtype=np.uint64)
        check(lib.scb_transcript_c_1(self._h, _p64(out)))
        return self.F.from_mont(out)[0]

    def messages(self) -> List[bytes]:
        ln, n = C.c_size_t(), C.c_uint32()
        check(lib.scb_transcript_bytes(self._h, None, 0, C.byref(ln), None, 0, C.byref(n)))
        buf = (C.c_uint8 * max(ln.value, 1))()
        offs = np.zeros(n.value + 1, dtype=np.uint64)
        check(lib.scb_transcript_bytes(self._h, buf, ln.value, C.byref(ln), _p64(offs), n.value, C.byref(n)))
        raw = bytes(buf[: ln.value])
        return [raw[int(offs[i]) : int(offs[i + 1])] for i in range(n.value)]


def device_count() -> int:
    n = C.c_int()
    check(lib.scb_device_count(C.byref(n)))
    return n.value


def launch_count(reset: bool = False) -> int:
    n = C.c_uint64()
    check(lib.scb_launch_count(C.byref(n), 1 if reset else 0))
    return n.value


def set_stream(cuda_stream_ptr: int) -> None:
    check(lib.scb_set_stream(C.c_void_p(cuda_stream_ptr)))


def synchronize() -> None:
    check(lib.scb_synchronize())
