"""Python big-integer ORACLE for the sum-check hot path of montekki/thaler-study.

TEST INFRASTRUCTURE ONLY.  Nothing under ``thaler_study_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker.

It is a restatement of the reference's algorithm in the reference's own naive
structure, with field elements as canonical Python ints (``x % p``), so that it
can be checked against the Rust by reading.  Every function cites the
reference ``file:line`` it follows (paths relative to ``/root/reference``).

The arithmetic itself lives in the un-vendored crates ``ark-ff`` / ``ark-poly``
/ ``ark-serialize`` ``= "0.6"`` (``Cargo.toml:20-25``, no lock file => patch
version unpinned).  Their *published* behaviour is restated here ([ARK] tags);
the source is not in this container and could not be re-read.

Pinning status
--------------
Pinned against every known-answer vector the reference's own tests hold for
this path (see ``tests/test_oracle_kats.py``): the 5x5 MLE table
(``multilinear-extensions/src/lib.rs:83-89``), the ``fix_variables`` KAT
(``sum-check-protocol/src/lib.rs:400-415``), ``c_1 == (A*B)[i][j]``
(``matrix-multiplication/src/lib.rs:339-340``), ``c_1 == 6*triangles``
(``triangle-counting/src/lib.rs:294-300``), ``restrict_poly -> [32,385,383]``
(``gkr-protocol/src/lib.rs:540-547``), circuit layers
(``gkr-protocol/src/circuit.rs:263-267``), GKR outputs ``[36,6]`` / ``[2,2]``
(``gkr-protocol/src/lib.rs:580-585,658-663``) and the verifier round
invariants.  PARITY UNPINNED (the reference has no vector for them): transcript
*bytes*, ``hash_to_field`` outputs, explicit-zero terms of ``SparsePolynomial``,
any field wider than one limb, any table above 2^10 entries.  For those the
contract is "this oracle, the C oracle (oracle/oracle.c) and the CUDA engine
agree", and DESIGN.md says so.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------
# field helpers  ([ARK] Fp<MontBackend<_,N>,N>; canonical value semantics)
# --------------------------------------------------------------------------


@dataclass(frozen=True)
class Field:
    """A prime field F_p.  Elements are canonical ints in [0, p)."""

    p: int

    @property
    def bits(self) -> int:  # [ARK] MODULUS_BIT_SIZE
        return self.p.bit_length()

    @property
    def n_limbs(self) -> int:  # number of u64 limbs ark would use
        return (self.bits + 63) // 64

    @property
    def ser_bytes(self) -> int:  # [ARK] serialized_size(Compress::No) of an Fp
        return (self.bits + 7) // 8

    def inv(self, a: int) -> int:
        a %= self.p
        if a == 0:
            raise ZeroDivisionError("inverse of zero")
        return pow(a, self.p - 2, self.p)

    def two_adicity(self) -> int:
        n, s = self.p - 1, 0
        while n % 2 == 0:
            n //= 2
            s += 1
        return s

    def mont_r(self) -> int:  # R = 2^(64 N) mod p
        return pow(2, 64 * self.n_limbs, self.p)

    def to_mont(self, a: int) -> int:
        return (a * self.mont_r()) % self.p

    def from_mont(self, a: int) -> int:
        return (a * self.inv(self.mont_r())) % self.p


# The reference's concrete fields and the wider ones the bench uses.
FP5 = Field(5)  # sum-check-protocol/src/lib.rs:349-354
FP389 = Field(389)  # triangle-counting/src/lib.rs:226-231
FP1572869 = Field(1572869)  # triangle-counting/src/lib.rs:272-277
# BLS12-381 scalar field = ark_ed_on_bls12_381::Fq (workspace dep, Cargo.toml:20)
BLS12_381_FR = Field(0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001)


# --------------------------------------------------------------------------
# [ARK] ark_poly::univariate::SparsePolynomial
# --------------------------------------------------------------------------


class SparsePoly:
    """[ARK] ``univariate::SparsePolynomial<F>``: ``Vec<(usize, F)>`` sorted by degree.

    The zero-handling rules below decide transcript bytes (SURVEY section 7,
    hard part 2) and are restated literally.
    """

    __slots__ = ("F", "coeffs")

    def __init__(self, F: Field, coeffs: List[Tuple[int, int]]):
        self.F = F
        self.coeffs = coeffs

    # [ARK] SparsePolynomial::zero()
    @staticmethod
    def zero(F: Field) -> "SparsePoly":
        return SparsePoly(F, [])

    # [ARK] from_coefficients_vec: pop trailing zero terms (in the GIVEN order),
    # then stable-sort by degree, then assert the last term is non-zero.
    @staticmethod
    def from_coefficients_vec(F: Field, coeffs: Sequence[Tuple[int, int]]) -> "SparsePoly":
        c = [(int(d), int(v) % F.p) for d, v in coeffs]
        while c and c[-1][1] == 0:
            c.pop()
        c.sort(key=lambda t: t[0])
        assert (not c) or c[-1][1] != 0
        return SparsePoly(F, c)

    from_coefficients_slice = from_coefficients_vec

    # [ARK] is_zero: empty or all coefficients zero
    def is_zero(self) -> bool:
        return all(v == 0 for _, v in self.coeffs)

    def degree(self) -> int:
        return 0 if self.is_zero() else self.coeffs[-1][0]

    # [ARK] impl Add for &SparsePolynomial: single-pass sorted merge; an
    # equal-degree pair is dropped only if it SUMS to zero; zero operand => clone
    # of the other operand (explicit zero terms included).
    def __add__(self, other: "SparsePoly") -> "SparsePoly":
        if self.is_zero():
            return SparsePoly(self.F, list(other.coeffs))
        if other.is_zero():
            return SparsePoly(self.F, list(self.coeffs))
        p = self.F.p
        res: List[Tuple[int, int]] = []
        i = j = 0
        a, b = self.coeffs, other.coeffs
        while True:
            if i == len(a) and j == len(b):
                break
            if i == len(a):
                res.extend(b[j:])
                break
            if j == len(b):
                res.extend(a[i:])
                break
            da, ca = a[i]
            db, cb = b[j]
            if da < db:
                res.append((da, ca))
                i += 1
            elif da == db:
                s = (ca + cb) % p
                if s != 0:
                    res.append((da, s))
                i += 1
                j += 1
            else:
                res.append((db, cb))
                j += 1
        return SparsePoly(self.F, res)

    # [ARK] SparsePolynomial::mul: BTreeMap accumulate, then from_coefficients_vec
    def mul(self, other: "SparsePoly") -> "SparsePoly":
        if self.is_zero() or other.is_zero():
            return SparsePoly.zero(self.F)
        acc = {}
        p = self.F.p
        for i, ci in self.coeffs:
            for j, cj in other.coeffs:
                acc[i + j] = (acc.get(i + j, 0) + ci * cj) % p
        return SparsePoly.from_coefficients_vec(self.F, sorted(acc.items()))

    # [ARK] Polynomial::evaluate: sum of c * x^i
    def evaluate(self, x: int) -> int:
        p = self.F.p
        if self.is_zero():
            return 0
        return sum(c * pow(x, d, p) for d, c in self.coeffs) % p

    # [ARK] From<DensePolynomial>: keep exactly the non-zero coefficients
    @staticmethod
    def from_dense(F: Field, dense: Sequence[int]) -> "SparsePoly":
        return SparsePoly.from_coefficients_vec(
            F, [(i, c % F.p) for i, c in enumerate(dense) if c % F.p != 0]
        )

    def to_dense(self) -> List[int]:
        if not self.coeffs:
            return []
        out = [0] * (self.coeffs[-1][0] + 1)
        for d, c in self.coeffs:
            out[d] = (out[d] + c) % self.F.p
        while out and out[-1] == 0:
            out.pop()
        return out

    def __eq__(self, other) -> bool:  # [ARK] derive(PartialEq) on coeffs
        return isinstance(other, SparsePoly) and self.coeffs == other.coeffs

    def __repr__(self) -> str:
        return f"SparsePoly({self.coeffs})"

    # [ARK] CanonicalSerialize (derive): Vec<(usize, F)> => u64 LE len, then
    # per item u64 LE degree and the canonical value in ceil(bits/8) LE bytes.
    def serialize(self) -> bytes:
        out = len(self.coeffs).to_bytes(8, "little")
        for d, c in self.coeffs:
            out += d.to_bytes(8, "little") + ser_field(self.F, c)
        return out

    @staticmethod
    def deserialize(F: Field, data: bytes, off: int = 0) -> Tuple["SparsePoly", int]:
        n = int.from_bytes(data[off : off + 8], "little")
        off += 8
        coeffs = []
        for _ in range(n):
            d = int.from_bytes(data[off : off + 8], "little")
            off += 8
            c = int.from_bytes(data[off : off + F.ser_bytes], "little")
            off += F.ser_bytes
            coeffs.append((d, c))
        return SparsePoly(F, coeffs), off


def ser_field(F: Field, a: int) -> bytes:
    """[ARK] Fp::serialize_uncompressed: canonical value, ceil(bits/8) bytes LE."""
    return int(a % F.p).to_bytes(F.ser_bytes, "little")


# --------------------------------------------------------------------------
# [ARK] ark_poly::DenseMultilinearExtension  (variable i <-> index bit i)
# --------------------------------------------------------------------------


class DenseMLE:
    __slots__ = ("F", "num_vars", "evals")

    def __init__(self, F: Field, num_vars: int, evals: Sequence[int]):
        assert len(evals) == 1 << num_vars, "size of evaluations must be 2^num_vars"
        self.F = F
        self.num_vars = num_vars
        self.evals = [int(e) % F.p for e in evals]

    def clone(self) -> "DenseMLE":
        return DenseMLE(self.F, self.num_vars, self.evals)

    # [ARK] fix_variables: for each coordinate r (left to right):
    #   t[b] = t[2b] + r * (t[2b+1] - t[2b])
    def fix_variables(self, partial_point: Sequence[int]) -> "DenseMLE":
        assert len(partial_point) <= self.num_vars, "invalid size of partial point"
        p = self.F.p
        poly = list(self.evals)
        nv = self.num_vars
        for i, r in enumerate(partial_point, start=1):
            for b in range(1 << (nv - i)):
                left, right = poly[2 * b], poly[2 * b + 1]
                poly[b] = (left + r * (right - left)) % p
        dim = len(partial_point)
        return DenseMLE(self.F, nv - dim, poly[: 1 << (nv - dim)])

    # [ARK] Polynomial::evaluate = fix_variables(point)[0]
    def evaluate(self, point: Sequence[int]) -> int:
        assert len(point) == self.num_vars
        return self.fix_variables(point).evals[0]

    # [ARK] relabel(a, b, k): swap index bit-blocks [a,a+k) <-> [b,b+k)
    def relabel(self, a: int, b: int, k: int) -> "DenseMLE":
        if a > b:
            a, b = b, a
        if a == b or k == 0:
            return self.clone()
        assert b + k <= self.num_vars and a + k <= b
        ev = list(self.evals)
        mask = (1 << k) - 1
        for i in range(len(ev)):
            x = ((i >> a) ^ (i >> b)) & mask
            j = i ^ ((x << a) | (x << b))
            if i < j:
                ev[i], ev[j] = ev[j], ev[i]
        return DenseMLE(self.F, self.num_vars, ev)

    def to_evaluations(self) -> List[int]:
        return list(self.evals)


# --------------------------------------------------------------------------
# multilinear-extensions/src/lib.rs
# --------------------------------------------------------------------------


def vsbw_multilinear_from_evaluations(F: Field, evals: Sequence[int], r: Sequence[int]) -> int:
    """multilinear-extensions/src/lib.rs:6-24 (chi table by doubling, r[0] -> MSB)."""
    p = F.p
    table = [1]
    for r_j in r:
        new = []
        for e in table:
            new.append(e * (1 - r_j) % p)
            new.append(e * r_j % p)
        table = new
    acc = 0
    for w, pj in zip(table, evals):
        acc = (acc + w * pj) % p
    return acc


def lagrange_basis_poly_at(F: Field, x: Sequence[int], w: Sequence[int]) -> int:
    """multilinear-extensions/src/lib.rs:50-60."""
    p = F.p
    res = 1
    for xi, wi in zip(x, w):
        res = res * (xi * wi + (1 - xi) * (1 - wi)) % p
    return res


def cti_multilinear_from_evaluations(F: Field, evals: Sequence[int], r: Sequence[int]) -> int:
    """multilinear-extensions/src/lib.rs:29-48 (streaming, big-endian bits :37-42)."""
    p = F.p
    res = 0
    n = len(r)
    for i, e in enumerate(evals):
        w = [1 if (i >> j) & 1 else 0 for j in reversed(range(n))]
        res = (res + e * lagrange_basis_poly_at(F, r, w)) % p
    return res


# --------------------------------------------------------------------------
# [ARK] radix-2 domain of size 4 + Evaluations::interpolate
# --------------------------------------------------------------------------


def domain4_elements(F: Field, generator: int = 2) -> List[int]:
    """[ARK] GeneralEvaluationDomain::new(3) -> Radix2 domain of size 4; elements 1,w,w^2,w^3.

    Needs two-adicity >= 2 (true for 5, 389, 1572869).  ``generator`` is the
    ``#[generator]`` of the MontConfig (2 in every reference field).
    """
    s = F.two_adicity()
    if s < 2:
        raise ValueError("GeneralEvaluationDomain::new(3) is None for this field (unwrap panics)")
    two_adic_root = pow(generator, (F.p - 1) >> s, F.p)
    w = pow(two_adic_root, 1 << (s - 2), F.p)
    return [pow(w, i, F.p) for i in range(4)]


def interpolate_domain4(F: Field, evals: Sequence[int], generator: int = 2) -> List[int]:
    """[ARK] Evaluations::interpolate: IFFT over the size-4 domain, trailing zeros stripped."""
    p = F.p
    elems = domain4_elements(F, generator)
    w_inv = F.inv(elems[1])
    n_inv = F.inv(4)
    coeffs = []
    for k in range(4):
        acc = 0
        for j in range(4):
            acc += evals[j] * pow(w_inv, j * k, p)
        coeffs.append(acc * n_inv % p)
    while coeffs and coeffs[-1] == 0:
        coeffs.pop()
    return coeffs


# --------------------------------------------------------------------------
# sum-check-protocol/src/lib.rs
# --------------------------------------------------------------------------


def boolean_hypercube(n: int) -> Iterable[List[int]]:
    """sum-check-protocol/src/lib.rs:34-70 (LSB-first 0/1 vectors)."""
    for cur in range(1 << n):
        yield [(cur >> i) & 1 for i in range(n)]


class SumCheckPolynomial:
    """trait SumCheckPolynomial<F>, sum-check-protocol/src/lib.rs:121-156."""

    F: Field

    def evaluate(self, point: Sequence[int]) -> Optional[int]:
        raise NotImplementedError

    def fix_variables(self, partial_point: Sequence[int]) -> "SumCheckPolynomial":
        raise NotImplementedError

    def to_univariate(self) -> SparsePoly:
        raise NotImplementedError

    def num_vars(self) -> int:
        raise NotImplementedError

    def to_evaluations(self) -> List[int]:
        raise NotImplementedError


class Prover:
    """sum-check-protocol/src/lib.rs:73-117."""

    def __init__(self, g: SumCheckPolynomial):
        self.g = g
        self.c_1_ = sum(g.to_evaluations()) % g.F.p  # :89
        self.num_vars_ = g.num_vars()
        self.r: List[int] = []

    def c_1(self) -> int:
        return self.c_1_

    def round(self, r_prev: int, j: int) -> SparsePoly:  # :105-112
        if j != 0:
            self.r.append(r_prev)
            self.g = self.g.fix_variables([r_prev])
        return self.g.to_univariate()

    def num_vars(self) -> int:
        return self.num_vars_


class ProverClaimMismatch(Exception):
    """sum-check-protocol/src/lib.rs:26-27."""


class NoPolySet(Exception):
    """sum-check-protocol/src/lib.rs:29-30."""


class Verifier:
    """sum-check-protocol/src/lib.rs:227-331.

    ``round`` returns ("JthRound", r_j) or ("FinalRound", bool); ``rng`` is any
    object with ``draw()`` (RngF, :13-21).
    """

    def __init__(self, n: int, g: Optional[SumCheckPolynomial], F: Optional[Field] = None):
        self.n = n
        self.c_1 = 0
        self.g_part: List[SparsePoly] = []
        self.r: List[int] = []
        self.g = g
        self.F = F if F is not None else g.F

    def set_c_1(self, c_1: int) -> None:
        self.c_1 = c_1

    def round(self, g_j: SparsePoly, rng):
        p = self.F.p
        r_j = rng.draw()
        if not self.r:  # :284-297
            evaluation = (g_j.evaluate(0) + g_j.evaluate(1)) % p
            if self.c_1 != evaluation:
                raise ProverClaimMismatch(f"start {self.c_1}", f"{evaluation}")
            self.g_part.append(g_j)
            self.r.append(r_j)
            return ("JthRound", r_j)
        elif len(self.r) == self.n - 1:  # :298-310
            self.r.append(r_j)
            if self.g is None:
                raise NoPolySet()
            lhs = g_j.evaluate(r_j)
            rhs = self.g.evaluate(self.r)
            assert lhs == rhs, (lhs, rhs)  # :303 assert_eq!
            return ("FinalRound", lhs == rhs)
        else:  # :311-329
            prev_evaluation = self.g_part[-1].evaluate(self.r[-1])
            evaluation = (g_j.evaluate(0) + g_j.evaluate(1)) % p
            if prev_evaluation != evaluation:
                raise ProverClaimMismatch(f"{prev_evaluation}", f"{evaluation}")
            self.g_part.append(g_j)
            self.r.append(r_j)
            return ("JthRound", r_j)


class RandNums:
    """fiat-shamir/src/lib.rs:102-119."""

    def __init__(self, nums: Sequence[int]):
        self.nums = list(nums)
        self.current = 0

    def draw(self) -> int:
        res = self.nums[self.current]
        self.current += 1
        return res


# ---- the slow generic impl for multivariate::SparsePolynomial (:158-224) ----


class SparseMVPoly(SumCheckPolynomial):
    """[ARK] multivariate::SparsePolynomial<F, SparseTerm> + the trait impl at
    sum-check-protocol/src/lib.rs:158-224.  terms: list of (coeff, ((var, power), ...)).

    [ARK] from_coefficients_vec sorts each term's (var,power) list by var and
    combines duplicate vars, sorts terms and merges equal terms (dropping zero sums).
    """

    def __init__(self, F: Field, num_vars: int, terms: Sequence[Tuple[int, Sequence[Tuple[int, int]]]]):
        self.F = F
        self.nv = num_vars
        self.terms = self._normalise(F, terms)

    @staticmethod
    def _norm_term(term: Sequence[Tuple[int, int]]) -> Tuple[Tuple[int, int], ...]:
        acc = {}
        for var, power in term:
            if power != 0:
                acc[var] = acc.get(var, 0) + power
        return tuple(sorted(acc.items()))

    @staticmethod
    def _term_key(t: Tuple[Tuple[int, int], ...]):
        # [ARK] SparseTerm Ord: by total degree, then variable-wise (higher var first)
        deg = sum(pw for _, pw in t)
        return (deg, tuple(sorted(t, reverse=True)))

    @classmethod
    def _normalise(cls, F, terms):
        acc = {}
        for c, t in terms:
            nt = cls._norm_term(t)
            acc[nt] = (acc.get(nt, 0) + c) % F.p
        out = [(c, t) for t, c in acc.items() if c != 0]
        out.sort(key=lambda ct: cls._term_key(ct[1]))
        return out

    def _eval_term(self, term, point) -> int:
        p = self.F.p
        v = 1
        for var, power in term:
            v = v * pow(point[var], power, p) % p
        return v

    def evaluate(self, point):  # :159-161
        p = self.F.p
        return sum(c * self._eval_term(t, point) for c, t in self.terms) % p

    def fix_variables(self, partial_point):  # :163-187
        k = len(partial_point)
        full = list(partial_point) + [1] * (self.nv - k)
        new_terms = []
        for c, t in self.terms:
            ev = self._eval_term(t, full) * c % self.F.p
            nt = tuple((var - k, pw) for var, pw in t if var >= k)
            new_terms.append((ev, nt))
        return SparseMVPoly(self.F, self.nv - k, new_terms)

    def to_univariate(self):  # :189-213
        res = SparsePoly.zero(self.F)
        for pt in boolean_hypercube(self.nv - 1):
            point = [1] + pt
            r = SparsePoly.zero(self.F)
            for c, t in self.terms:
                ev = self._eval_term(t, point) * c % self.F.p
                power = next((pw for var, pw in t if var == 0), 0)
                r = r + SparsePoly.from_coefficients_slice(self.F, [(power, ev)])
            res = res + r
        return res

    def num_vars(self):
        return self.nv

    def to_evaluations(self):  # :219-223
        return [self.evaluate(pt) for pt in boolean_hypercube(self.nv)]


# --------------------------------------------------------------------------
# matrix-multiplication/src/lib.rs
# --------------------------------------------------------------------------


def interpolate_quadratic_poly(F: Field, points: Sequence[Tuple[int, int]]) -> SparsePoly:
    """matrix-multiplication/src/lib.rs:17-60, literally (explicit (0,0) terms included)."""
    p = F.p
    (x0, y0), (x1, y1), (x2, y2) = points
    den1 = (x0 - x1) * (x0 - x2) % p
    den2 = (x1 - x0) * (x1 - x2) % p
    den3 = (x2 - x0) * (x2 - x1) % p

    def scaled(c, y, den):
        return [(d, v * y % p * F.inv(den) % p) for d, v in c]

    c1 = scaled([(0, x1 * x2 % p), (1, (-x1 - x2) % p), (2, 1)], y0, den1)
    c2 = scaled([(0, x0 * x2 % p), (1, (-x0 - x2) % p), (2, 1)], y1, den2)
    c3 = scaled([(0, x0 * x1 % p), (1, (-x0 - x1) % p), (2, 1)], y2, den3)
    p1 = SparsePoly.from_coefficients_vec(F, c1)
    p2 = SparsePoly.from_coefficients_vec(F, c2)
    p3 = SparsePoly.from_coefficients_vec(F, c3)
    return (p1 + p2) + p3


class MatMulG(SumCheckPolynomial):
    """matrix-multiplication/src/lib.rs:12-15, 62-147."""

    def __init__(self, F: Field, f_a: DenseMLE, f_b: DenseMLE):
        self.F = F
        self.f_a = f_a
        self.f_b = f_b

    @staticmethod
    def new(F: Field, n: int, a: Sequence[int], b: Sequence[int], point: Sequence[int]) -> "MatMulG":
        # :77-92
        f_a = DenseMLE(F, 2 * n, list(a)).relabel(0, n, n).fix_variables(point[:n])
        f_b = DenseMLE(F, 2 * n, list(b)).fix_variables(point[n:])
        assert f_a.num_vars == n and f_b.num_vars == n
        return MatMulG(F, f_a, f_b)

    def evaluate(self, point):  # :96-101
        return self.f_a.evaluate(point) * self.f_b.evaluate(point) % self.F.p

    def fix_variables(self, partial_point):  # :103-108
        return MatMulG(self.F, self.f_a.fix_variables(partial_point), self.f_b.fix_variables(partial_point))

    def round_evals(self) -> List[int]:
        # :110-122 (one pass over adjacent pairs, X = 0, 1, 2)
        p = self.F.p
        a, b = self.f_a.evals, self.f_b.evals
        e = [0, 0, 0]
        for i in range(1 << self.num_vars()):
            if i & 1:
                e[1] = (e[1] + a[i] * b[i]) % p
                e[2] = (e[2] + (2 * a[i] - a[i - 1]) * (2 * b[i] - b[i - 1])) % p
            else:
                e[0] = (e[0] + a[i] * b[i]) % p
        return e

    def to_univariate(self):  # :110-131
        e = self.round_evals()
        return interpolate_quadratic_poly(self.F, [(0, e[0]), (1, e[1]), (2 % self.F.p, e[2])])

    def num_vars(self):
        return self.f_a.num_vars

    def to_evaluations(self):  # :137-146
        p = self.F.p
        return [x * y % p for x, y in zip(self.f_a.evals, self.f_b.evals)]


# --------------------------------------------------------------------------
# ProductMLE<K>: the new impl the BASELINE configs name (last
# paragraph).  Direct generalisation of MatMulG: K dense MLEs over the SAME
# variables; message = sums at X = 0..K over adjacent pairs, Lagrange to
# coefficients, Dense -> Sparse (all zero coefficients dropped).
# --------------------------------------------------------------------------


def lagrange_to_coeffs(F: Field, ys: Sequence[int]) -> List[int]:
    """Coefficients of the unique deg<=d polynomial through (0,y0)..(d,yd); trailing zeros stripped."""
    p = F.p
    n = len(ys)
    coeffs = [0] * n
    for i in range(n):
        # numerator polynomial prod_{j != i} (X - j), denominator prod_{j != i} (i - j)
        num = [1]
        den = 1
        for j in range(n):
            if j == i:
                continue
            num = [((num[k - 1] if k > 0 else 0) - j * (num[k] if k < len(num) else 0)) % p for k in range(len(num) + 1)]
            den = den * (i - j) % p
        s = ys[i] * F.inv(den) % p
        for k in range(n):
            coeffs[k] = (coeffs[k] + s * num[k]) % p
    while coeffs and coeffs[-1] == 0:
        coeffs.pop()
    return coeffs


class ProductMLE(SumCheckPolynomial):
    def __init__(self, F: Field, tables: Sequence[DenseMLE]):
        assert len(tables) >= 1 and all(t.num_vars == tables[0].num_vars for t in tables)
        self.F = F
        self.tables = list(tables)

    def evaluate(self, point):
        if len(point) != self.num_vars():
            return None
        v = 1
        for t in self.tables:
            v = v * t.evaluate(point) % self.F.p
        return v

    def fix_variables(self, partial_point):
        return ProductMLE(self.F, [t.fix_variables(partial_point) for t in self.tables])

    def round_evals(self, n_points: Optional[int] = None) -> List[int]:
        p = self.F.p
        K = len(self.tables)
        npts = K + 1 if n_points is None else n_points
        e = [0] * npts
        for b in range(1 << (self.num_vars() - 1)):
            prod = [1] * npts
            for t in self.tables:
                lo, hi = t.evals[2 * b], t.evals[2 * b + 1]
                d = hi - lo
                for x in range(npts):
                    prod[x] = prod[x] * (lo + x * d) % p
            for x in range(npts):
                e[x] = (e[x] + prod[x]) % p
        return e

    def to_univariate(self):
        return SparsePoly.from_dense(self.F, lagrange_to_coeffs(self.F, self.round_evals()))

    def num_vars(self):
        return self.tables[0].num_vars

    def to_evaluations(self):
        p = self.F.p
        out = list(self.tables[0].evals)
        for t in self.tables[1:]:
            out = [x * y % p for x, y in zip(out, t.evals)]
        return out


# --------------------------------------------------------------------------
# triangle-counting/src/lib.rs
# --------------------------------------------------------------------------


def _idx(i: int, j: int, num_vars: int) -> int:
    """triangle-counting/src/lib.rs:170-172, gkr-protocol/src/round_polynomial.rs:123-125."""
    return (i << num_vars) | j


class TriangleG(SumCheckPolynomial):
    """triangle-counting/src/lib.rs:22-27, 29-166."""

    def __init__(self, F: Field, f1: DenseMLE, f2: DenseMLE, f3: DenseMLE, var_len: int, generator: int = 2):
        self.F = F
        self.f_a_1, self.f_a_2, self.f_a_3 = f1, f2, f3
        self.var_len = var_len
        self.generator = generator

    @staticmethod
    def new_adj_matrix(F: Field, num_vars: int, matrix: Sequence[bool], generator: int = 2) -> "TriangleG":
        g = DenseMLE(F, num_vars, [1 if b else 0 for b in matrix])  # :32-51
        return TriangleG(F, g.clone(), g.clone(), g, num_vars // 2, generator)

    def x_vars_num(self):  # :53-55
        return max(self.f_a_1.num_vars - self.var_len, 0)

    def y_vars_num(self):  # :57-59
        return max(self.f_a_2.num_vars - self.var_len, 0)

    def z_vars_num(self):  # :61-67
        return self.f_a_3.num_vars if self.f_a_3.num_vars < self.var_len else self.var_len

    def evaluate(self, point):  # :71-87
        xn, yn = self.x_vars_num(), self.y_vars_num()
        e1 = self.f_a_1.evaluate(point[: xn + yn])
        e2 = self.f_a_2.evaluate(point[xn:])
        e3 = self.f_a_3.evaluate(list(point[:xn]) + list(point[xn + yn :]))
        return e1 * e3 % self.F.p * e2 % self.F.p

    def fix_variables(self, partial_point):  # :89-118
        pp = list(partial_point)
        xn, yn = self.x_vars_num(), self.y_vars_num()
        x_y = pp[: min(xn + yn, len(pp))]
        y_z = pp[xn:] if xn <= len(pp) else []
        x_z = pp[: min(xn, len(pp))] + (pp[xn + yn :] if xn + yn <= len(pp) else [])
        return TriangleG(
            self.F,
            self.f_a_1.fix_variables(x_y),
            self.f_a_2.fix_variables(y_z),
            self.f_a_3.fix_variables(x_z),
            self.var_len,
            self.generator,
        )

    def to_univariate(self):  # :120-132
        evals = [
            sum(self.fix_variables([e]).to_evaluations()) % self.F.p
            for e in domain4_elements(self.F, self.generator)
        ]
        return SparsePoly.from_dense(self.F, interpolate_domain4(self.F, evals, self.generator))

    def num_vars(self):  # :134-136
        return self.x_vars_num() + self.y_vars_num() + self.z_vars_num()

    def to_evaluations(self):  # :138-165
        p = self.F.p
        e1, e2, e3 = self.f_a_1.evals, self.f_a_2.evals, self.f_a_3.evals
        xn, yn, zn = self.x_vars_num(), self.y_vars_num(), self.z_vars_num()
        res = []
        for x in range(1 << xn):
            for y in range(1 << yn):
                for z in range(1 << zn):
                    res.append(e1[_idx(y, x, xn)] * e2[_idx(z, y, yn)] % p * e3[_idx(z, x, xn)] % p)
        return res


# --------------------------------------------------------------------------
# gkr-protocol/src/round_polynomial.rs, circuit.rs, lib.rs
# --------------------------------------------------------------------------


class GkrW(SumCheckPolynomial):
    """gkr-protocol/src/round_polynomial.rs:23-119."""

    def __init__(self, F: Field, add_i: DenseMLE, mul_i: DenseMLE, w_b: DenseMLE, w_c: DenseMLE, generator: int = 2):
        self.F = F
        self.add_i, self.mul_i, self.w_b, self.w_c = add_i, mul_i, w_b, w_c
        self.generator = generator

    def evaluate(self, point):  # :48-57
        p = self.F.p
        k = self.w_b.num_vars
        b, c = point[:k], point[k:]
        add_e = self.add_i.evaluate(point)
        mul_e = self.mul_i.evaluate(point)
        wb = self.w_b.evaluate(b)
        wc = self.w_c.evaluate(c)
        return (add_e * (wb + wc) + mul_e * (wb * wc)) % p

    def fix_variables(self, partial_point):  # :59-76
        pp = list(partial_point)
        k = self.w_b.num_vars
        b_partial = pp[: min(k, len(pp))]
        c_partial = pp[k:] if k <= len(pp) else []
        return GkrW(
            self.F,
            self.add_i.fix_variables(pp),
            self.mul_i.fix_variables(pp),
            self.w_b.fix_variables(b_partial),
            self.w_c.fix_variables(c_partial),
            self.generator,
        )

    def to_univariate(self):  # :78-90
        evals = [
            sum(self.fix_variables([e]).to_evaluations()) % self.F.p
            for e in domain4_elements(self.F, self.generator)
        ]
        return SparsePoly.from_dense(self.F, interpolate_domain4(self.F, evals, self.generator))

    def num_vars(self):  # :92-94
        return self.add_i.num_vars

    def to_evaluations(self):  # :96-118
        p = self.F.p
        k = self.w_b.num_vars
        add, mul = self.add_i.evals, self.mul_i.evals
        res = []
        for b_idx, wb in enumerate(self.w_b.evals):
            for c_idx, wc in enumerate(self.w_c.evals):
                bc = _idx(c_idx, b_idx, k)
                res.append((add[bc] * (wb + wc) + mul[bc] * (wb * wc)) % p)
        return res


ADD, MUL = "add", "mul"


class Circuit:
    """gkr-protocol/src/circuit.rs:70-212.  layers[0] = output layer; gate = (type, (in0, in1))."""

    def __init__(self, layers: Sequence[Sequence[Tuple[str, Tuple[int, int]]]], num_inputs: int):
        self.layers = [list(l) for l in layers]
        self.num_inputs = num_inputs

    def num_vars_at(self, layer: int) -> Optional[int]:  # :86-96
        if layer < len(self.layers):
            n = len(self.layers[layer])
        elif layer == len(self.layers):
            n = self.num_inputs
        else:
            return None
        return (n & -n).bit_length() - 1 if n else 64  # trailing_zeros

    def evaluate(self, F: Field, inp: Sequence[int]) -> List[List[int]]:  # :99-124
        p = F.p
        layers = [list(inp)]
        cur = list(inp)
        for layer in reversed(self.layers):
            cur = [
                (cur[i0] + cur[i1]) % p if t == ADD else (cur[i0] * cur[i1]) % p
                for t, (i0, i1) in layer
            ]
            layers.append(cur)
        layers.reverse()
        return layers

    def add_i(self, i, a, b, c) -> bool:  # :127-131
        t, (i0, i1) = self.layers[i][a]
        return t == ADD and i0 == b and i1 == c

    def mul_i(self, i, a, b, c) -> bool:  # :134-138
        t, (i0, i1) = self.layers[i][a]
        return t == MUL and i0 == b and i1 == c

    def wiring_tables(self, F: Field, i: int) -> Tuple[DenseMLE, DenseMLE]:
        # :152-178 / gkr-protocol/src/lib.rs:385-413 (c outer, b middle, a inner)
        kc = self.num_vars_at(i)
        kn = self.num_vars_at(i + 1)
        add, mul = [], []
        for c in range(1 << kn):
            for b in range(1 << kn):
                for a in range(1 << kc):
                    add.append(1 if self.add_i(i, a, b, c) else 0)
                    mul.append(1 if self.mul_i(i, a, b, c) else 0)
        nv = kc + 2 * kn
        return DenseMLE(F, nv, add), DenseMLE(F, nv, mul)

    def add_i_ext(self, F: Field, r_i, i) -> DenseMLE:  # :152-181
        return self.wiring_tables(F, i)[0].fix_variables(r_i)

    def mul_i_ext(self, F: Field, r_i, i) -> DenseMLE:  # :183-212
        return self.wiring_tables(F, i)[1].fix_variables(r_i)


def circuit_from_book() -> Circuit:
    """gkr-protocol/src/circuit.rs:215-253."""
    return Circuit(
        [
            [(MUL, (0, 1)), (MUL, (2, 3))],
            [(MUL, (0, 0)), (MUL, (1, 1)), (MUL, (1, 2)), (MUL, (3, 3))],
        ],
        4,
    )


def three_layer_circuit() -> Circuit:
    """gkr-protocol/src/lib.rs:488-504."""
    return Circuit(
        [
            [(ADD, (0, 1)), (ADD, (2, 3))],
            [(ADD, (0, 1)), (ADD, (2, 3)), (ADD, (4, 5)), (ADD, (6, 7))],
        ],
        8,
    )


def line(F: Field, b: Sequence[int], c: Sequence[int]) -> List[SparsePoly]:
    """gkr-protocol/src/lib.rs:278-284."""
    return [SparsePoly.from_coefficients_slice(F, [(0, bi), (1, (ci - bi) % F.p)]) for bi, ci in zip(b, c)]


def restrict_poly(F: Field, b: Sequence[int], c: Sequence[int], mle: DenseMLE) -> SparsePoly:
    """gkr-protocol/src/lib.rs:291-321."""
    p = F.p
    k = [(ci - bi) % p for bi, ci in zip(b, c)]
    res = SparsePoly.zero(F)
    for i, ev in enumerate(mle.evals):
        poly = SparsePoly.from_coefficients_vec(F, [(0, ev)])
        for bit in range(mle.num_vars):
            bp = SparsePoly.from_coefficients_vec(F, [(0, b[bit]), (1, k[bit])])
            if i & (1 << bit) == 0:
                # (&DensePolynomial[1] - &b).into(): dense subtraction then Dense->Sparse
                dense = [0, 0]
                for d, cf in bp.coeffs:
                    dense[d] = cf
                dense = [(1 - dense[0]) % p, (-dense[1]) % p]
                while dense and dense[-1] == 0:
                    dense.pop()
                bp = SparsePoly.from_dense(F, dense)
            poly = poly.mul(bp)
        res = res + poly
    return res


class GkrProver:
    """gkr-protocol/src/lib.rs:324-474 (message enums flattened to tuples)."""

    def __init__(self, F: Field, circuit: Circuit, inp: Sequence[int], generator: int = 2):
        self.F = F
        self.circuit = circuit
        self.layers = circuit.evaluate(F, inp)
        self.i = 0
        self.prover: Optional[Prover] = None
        self.w: Optional[DenseMLE] = None
        self.r: List[int] = []
        self.generator = generator

    def start_protocol(self):  # :363-367
        return ("Begin", list(self.layers[0]))

    def start_round(self, i: int, r_i: Sequence[int]):  # :373-436
        F = self.F
        kn = self.circuit.num_vars_at(i + 1)
        w_b = DenseMLE(F, kn, self.layers[i + 1])
        self.w = w_b.clone()
        w_c = w_b.clone()
        add_i, mul_i = self.circuit.wiring_tables(F, i)
        add_i = add_i.fix_variables(r_i)
        mul_i = mul_i.fix_variables(r_i)
        num_vars = add_i.num_vars
        assert add_i.num_vars == mul_i.num_vars == 2 * w_b.num_vars
        self.i = i
        self.prover = Prover(GkrW(F, add_i, mul_i, w_b, w_c, self.generator))
        self.r = []
        return ("StartSumCheck", self.prover.c_1(), i, num_vars)

    def round_msg(self, j: int):  # :439-456
        if j == 2 * self.circuit.num_vars_at(self.i + 1) - 1:
            half = len(self.r) // 2
            b, c = self.r[:half], self.r[half:]
            q = restrict_poly(self.F, b, c, self.w)
            p = self.prover.round(self.r[j - 1], j)
            return ("FinalRoundMessage", p, q)
        point = 1 if j == 0 else self.r[j - 1]
        return ("SumCheckProverMessage", self.prover.round(point, j))

    def receive_verifier_msg(self, msg):  # :459-468
        if msg[0] == "SumCheckRoundResult":
            kind, val = msg[1]
            assert kind == "JthRound"
            self.r.append(val)

    def c_1(self):
        return self.prover.c_1()


class GkrVerifier:
    """gkr-protocol/src/lib.rs:38-218.  ``rng.draw()`` stands in for F::rand(rng)."""

    def __init__(self, F: Field, circuit: Circuit):
        self.F = F
        self.circuit = circuit
        self.r: List[List[int]] = []
        self.m: List[int] = []
        self.state = None

    def receive_prover_msg(self, msg, rng):  # :177-207
        F, p = self.F, self.F.p
        kind = msg[0]
        if kind == "Begin":
            outs = msg[1]
            k0 = self.circuit.num_vars_at(0)
            d = DenseMLE(F, k0, outs)
            r_zero = [rng.draw() for _ in range(k0)]
            self.r = [r_zero]
            self.m = [d.evaluate(r_zero)]
            return ("R", list(r_zero))
        if kind == "StartSumCheck":  # :89-105
            _, c_1, rnd, num_vars = msg
            add_i = self.circuit.add_i_ext(F, self.r[-1], rnd)
            mul_i = self.circuit.mul_i_ext(F, self.r[-1], rnd)
            v = Verifier(num_vars, None, F)
            v.set_c_1(c_1)
            self.state = {"bc": [], "verifier": v, "add_i": add_i, "mul_i": mul_i}
            return ("RoundStarted", rnd)
        if kind == "SumCheckProverMessage":  # :121-137
            res = self.state["verifier"].round(msg[1], rng)
            if res[0] == "JthRound":
                self.state["bc"].append(res[1])
            return ("SumCheckRoundResult", res)
        if kind == "FinalRoundMessage":  # :139-174
            _, pp, q = msg
            bc = self.state["bc"]
            q0, q1 = q.evaluate(0), q.evaluate(1)
            ev = (self.state["add_i"].evaluate(bc) * (q0 + q1) + self.state["mul_i"].evaluate(bc) * q0 * q1) % p
            assert ev == pp.evaluate(bc[-1]), (ev, pp.evaluate(bc[-1]))
            r = rng.draw()
            half = len(bc) // 2
            ln = line(F, bc[:half], bc[half:])
            r_next = [e.evaluate(r) for e in ln]
            self.r.append(r_next)
            self.m.append(q.evaluate(r))
            return ("R", list(r_next))
        raise ValueError(kind)

    def final_random_point(self, rng):  # :108-119
        pt = rng.draw()
        self.state["bc"].append(pt)
        return ("SumCheckRoundResult", ("JthRound", pt))

    def check_input(self, inp: Sequence[int]) -> bool:  # :210-217
        w = DenseMLE(self.F, (len(inp)).bit_length() - 1, inp)
        return w.evaluate(self.r[-1]) == self.m[-1]


# --------------------------------------------------------------------------
# fiat-shamir/src/lib.rs  + [ARK] DefaultFieldHasher<Sha256, 128>
# --------------------------------------------------------------------------


def expand_message_xmd(msg: bytes, dst: bytes, n: int, block_size: int) -> bytes:
    """[ARK] ExpanderXmd::expand with SHA-256: RFC 9380 expand_message_xmd EXCEPT that
    Z_pad has ``block_size`` = len_per_base_elem zero bytes (ark's field), not 64."""
    b_len = 32
    ell = (n + b_len - 1) // b_len
    assert ell <= 255 and n < (1 << 16)
    dst_prime = dst + bytes([len(dst)])
    z_pad = bytes(block_size)
    lib_str = n.to_bytes(2, "big")
    b0 = hashlib.sha256(z_pad + msg + lib_str + b"\x00" + dst_prime).digest()
    bi = hashlib.sha256(b0 + b"\x01" + dst_prime).digest()
    out = bi
    for i in range(2, ell + 1):
        bi = hashlib.sha256(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dst_prime).digest()
        out += bi
    return out[:n]


def hash_to_field(F: Field, msg: bytes, dst: bytes = b"") -> int:
    """[ARK] DefaultFieldHasher<Sha256,128>::hash_to_field::<1>(msg)[0] with H::new(dst)
    (fiat-shamir/src/lib.rs:78 uses the empty DST)."""
    L = (F.bits + 128 + 7) // 8
    uniform = expand_message_xmd(msg, dst, L, L)
    return int.from_bytes(uniform, "big") % F.p


def prover_g_1(F: Field, prover: Prover) -> bytes:
    """fiat-shamir/src/lib.rs:45-53: (c_1, round(F::one(), 0)).serialize_uncompressed."""
    poly = prover.round(1, 0)
    return ser_field(F, prover.c_1()) + poly.serialize()


def generate_transcript(F: Field, prover: Prover) -> List[bytes]:
    """fiat-shamir/src/lib.rs:75-98."""
    g_1 = prover_g_1(F, prover)
    hash_input = bytearray(g_1)
    g = [g_1]
    for j in range(1, prover.num_vars()):
        r_j = hash_to_field(F, bytes(hash_input))
        g_j = prover.round(r_j, j).serialize()
        hash_input += g_j
        g.append(g_j)
    return g


def verify_transcript(F: Field, transcript: Sequence[bytes], verifier: Verifier) -> bool:
    """fiat-shamir/src/lib.rs:123-143 + InteractiveVerifier impl :151-171."""
    hash_input = bytearray()
    for j, gj in enumerate(transcript):
        hash_input += gj
        r_j = hash_to_field(F, bytes(hash_input))
        rng = RandNums([r_j])
        if j == 0:
            c_1 = int.from_bytes(gj[: F.ser_bytes], "little")
            poly, _ = SparsePoly.deserialize(F, gj, F.ser_bytes)
            verifier.set_c_1(c_1)
            verifier.round(poly, rng)
            continue
        poly, _ = SparsePoly.deserialize(F, gj, 0)
        res = verifier.round(poly, rng)
        if res[0] == "FinalRound" and not res[1]:
            return False
    return True


# --------------------------------------------------------------------------
# synthetic input generator shared with the CUDA engine and the C oracle
# (counter-based, splitmix64(seed, index) -> mod p)
# --------------------------------------------------------------------------

_M64 = (1 << 64) - 1


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def synth_element(F: Field, seed: int, index: int) -> int:
    """Canonical value of synthetic table entry ``index`` for ``seed``.

    N limbs of splitmix64((seed * 2^40 + index) * N + limb), top limb masked to the
    modulus bit-length, then reduced by conditional subtraction of p (value < 2p).
    The result is stored AS the Montgomery-form limbs (the table is uniform either way).
    """
    N = F.n_limbs
    base = ((seed << 40) + index) * N
    v = 0
    for l in range(N):
        v |= splitmix64((base + l) & _M64) << (64 * l)
    v &= (1 << F.bits) - 1
    if v >= F.p:
        v -= F.p
    return v
