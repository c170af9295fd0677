"""ctypes binding of oracle/oracle.c (the plain-C CPU oracle).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by
thaler_study_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class _Field(C.Structure):
    _fields_ = [
        ("n", C.c_uint32),
        ("p", C.c_uint64 * 4),
        ("inv", C.c_uint64),
        ("one", C.c_uint64 * 4),
        ("r2", C.c_uint64 * 4),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def int_to_limbs(x: int, n: int) -> List[int]:
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def limbs_to_int(l: Sequence[int]) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


class CField:
    """A field descriptor + numpy helpers.  Arrays are uint64[count, n_limbs], Montgomery form."""

    def __init__(self, p: int):
        self.p = p
        self.n = (p.bit_length() + 63) // 64
        self.f = _Field()
        mod = (C.c_uint64 * 4)(*int_to_limbs(p, 4))
        rc = lib().orc_field_init(C.byref(self.f), C.c_uint32(self.n), mod)
        if rc != 0:
            raise ValueError("bad modulus")

    # ---- conversions -----------------------------------------------------
    def pack(self, vals: Sequence[int]) -> np.ndarray:
        """canonical ints -> raw limb array (NOT Montgomery)."""
        a = np.zeros((len(vals), self.n), dtype=np.uint64)
        for i, v in enumerate(vals):
            for l in range(self.n):
                a[i, l] = (int(v) >> (64 * l)) & 0xFFFFFFFFFFFFFFFF
        return a

    def unpack(self, a: np.ndarray) -> List[int]:
        a = a.reshape(-1, self.n)
        return [limbs_to_int(row) for row in a.tolist()]

    def to_mont(self, vals: Sequence[int]) -> np.ndarray:
        raw = self.pack([int(v) % self.p for v in vals])
        out = np.empty_like(raw)
        lib().orc_to_mont(C.byref(self.f), _ptr(out), _ptr(raw), C.c_size_t(len(vals)))
        return out

    def from_mont(self, a: np.ndarray) -> List[int]:
        a = np.ascontiguousarray(a.reshape(-1, self.n))
        out = np.empty_like(a)
        lib().orc_from_mont(C.byref(self.f), _ptr(out), _ptr(a), C.c_size_t(a.shape[0]))
        return self.unpack(out)

    # ---- path functions ----------------------------------------------------
    def synth(self, seed: int, start: int, count: int) -> np.ndarray:
        out = np.empty((count, self.n), dtype=np.uint64)
        lib().orc_synth_fill(C.byref(self.f), C.c_uint64(seed), C.c_uint64(start), C.c_size_t(count), _ptr(out))
        return out

    def fix_variable(self, tab: np.ndarray, r: np.ndarray, threads: int = 1) -> np.ndarray:
        tab = np.ascontiguousarray(tab)
        out = np.empty((tab.shape[0] // 2, self.n), dtype=np.uint64)
        lib().orc_fix_variable(C.byref(self.f), _ptr(tab), C.c_size_t(tab.shape[0]), _ptr(np.ascontiguousarray(r)), _ptr(out), C.c_int(threads))
        return out

    def _tabs(self, tabs):
        tabs = [np.ascontiguousarray(t) for t in tabs]
        arr = (C.POINTER(C.c_uint64) * len(tabs))(*[_ptr(t) for t in tabs])
        return tabs, arr

    def product_sum(self, tabs, threads: int = 1) -> np.ndarray:
        tabs, arr = self._tabs(tabs)
        out = np.empty((1, self.n), dtype=np.uint64)
        lib().orc_product_sum(C.byref(self.f), C.c_uint32(len(tabs)), arr, C.c_size_t(tabs[0].shape[0]), _ptr(out), C.c_int(threads))
        return out

    def product_round_evals(self, tabs, npts: int, threads: int = 1) -> np.ndarray:
        tabs, arr = self._tabs(tabs)
        out = np.empty((npts, self.n), dtype=np.uint64)
        lib().orc_product_round_evals(C.byref(self.f), C.c_uint32(len(tabs)), arr, C.c_size_t(tabs[0].shape[0]), C.c_uint32(npts), _ptr(out), C.c_int(threads))
        return out

    def product_prove(self, tabs, challenges: np.ndarray, npts: int, threads: int = 1, want_c1: bool = True):
        """Prover::new + v rounds; returns (c_1, evals[v, npts, n]).  Input tables are not modified."""
        tabs, arr = self._tabs(tabs)
        v = int(tabs[0].shape[0]).bit_length() - 1
        c1 = np.zeros((1, self.n), dtype=np.uint64)
        ev = np.empty((v, npts, self.n), dtype=np.uint64)
        ch = np.ascontiguousarray(challenges.reshape(-1, self.n)) if v > 1 else np.zeros((1, self.n), dtype=np.uint64)
        lib().orc_product_prove(C.byref(self.f), C.c_uint32(len(tabs)), C.c_uint32(v), arr, _ptr(ch), C.c_uint32(npts),
                                _ptr(c1) if want_c1 else None, _ptr(ev), C.c_int(threads))
        return c1, ev

    def mle_vsbw(self, evals: np.ndarray, r: np.ndarray) -> np.ndarray:
        out = np.empty((1, self.n), dtype=np.uint64)
        lib().orc_mle_vsbw(C.byref(self.f), _ptr(np.ascontiguousarray(evals)), C.c_uint32(r.shape[0]), _ptr(np.ascontiguousarray(r)), _ptr(out))
        return out

    def mle_cti(self, evals: np.ndarray, r: np.ndarray) -> np.ndarray:
        out = np.empty((1, self.n), dtype=np.uint64)
        lib().orc_mle_cti(C.byref(self.f), _ptr(np.ascontiguousarray(evals)), C.c_uint32(r.shape[0]), _ptr(np.ascontiguousarray(r)), _ptr(out))
        return out

    def mle_evaluate_le(self, evals: np.ndarray, point: np.ndarray) -> np.ndarray:
        out = np.empty((1, self.n), dtype=np.uint64)
        lib().orc_mle_evaluate_le(C.byref(self.f), _ptr(np.ascontiguousarray(evals)), C.c_uint32(point.shape[0]), _ptr(np.ascontiguousarray(point)), _ptr(out))
        return out

    def triangle_sum(self, f1, f2, f3, xn, yn, zn) -> np.ndarray:
        out = np.empty((1, self.n), dtype=np.uint64)
        lib().orc_triangle_sum(C.byref(self.f), _ptr(np.ascontiguousarray(f1)), _ptr(np.ascontiguousarray(f2)), _ptr(np.ascontiguousarray(f3)),
                               C.c_uint32(xn), C.c_uint32(yn), C.c_uint32(zn), _ptr(out))
        return out

    def triangle_round_eval_at(self, f1, f2, f3, xn, yn, zn, e: np.ndarray) -> np.ndarray:
        out = np.empty((1, self.n), dtype=np.uint64)
        lib().orc_triangle_round_eval_at(C.byref(self.f), _ptr(np.ascontiguousarray(f1)), _ptr(np.ascontiguousarray(f2)), _ptr(np.ascontiguousarray(f3)),
                                         C.c_uint32(xn), C.c_uint32(yn), C.c_uint32(zn), _ptr(np.ascontiguousarray(e)), _ptr(out))
        return out

    def gkrw_sum(self, add, mul, wb, wc, bn, cn) -> np.ndarray:
        out = np.empty((1, self.n), dtype=np.uint64)
        lib().orc_gkrw_sum(C.byref(self.f), _ptr(np.ascontiguousarray(add)), _ptr(np.ascontiguousarray(mul)), _ptr(np.ascontiguousarray(wb)),
                           _ptr(np.ascontiguousarray(wc)), C.c_uint32(bn), C.c_uint32(cn), _ptr(out))
        return out

    def gkrw_round_eval_at(self, add, mul, wb, wc, bn, cn, e: np.ndarray) -> np.ndarray:
        out = np.empty((1, self.n), dtype=np.uint64)
        lib().orc_gkrw_round_eval_at(C.byref(self.f), _ptr(np.ascontiguousarray(add)), _ptr(np.ascontiguousarray(mul)), _ptr(np.ascontiguousarray(wb)),
                                     _ptr(np.ascontiguousarray(wc)), C.c_uint32(bn), C.c_uint32(cn), _ptr(np.ascontiguousarray(e)), _ptr(out))
        return out


def max_threads() -> int:
    return int(lib().orc_max_threads())
