"""Mirror of gkr_protocol::{Circuit, Prover, Verifier} (gkr-protocol/src/{circuit,lib}.rs) on the B200 engine.

The prover -- the hot part: circuit evaluation, the two k-round sum-checks per layer over gate-list wiring, the line
restriction -- runs in libsumcheck_b200.so (csrc/gkr.cuh, csrc/gkr_engine.inc).  The verifier's per-layer work is the
reference's: the sum-check Verifier (C++ host layer) plus, in final_round_message, add~/mul~ at (b*, c*) evaluated from
the gate list on the device (the reference's dense tables have 2^(k_i + 2 k_{i+1}) entries) and a handful of field
operations on canonical integers.  Messages are tuples mirroring the reference's enums (lib.rs:222-275).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import api
from ._lib import check, lib, u8p, u32p

ADD, MUL = 0, 1


class Circuit:
    """gkr_protocol::circuit::Circuit: layers[0] is the output layer; a gate is (type, (in0, in1))."""

    def __init__(self, F: api.Field, layers: Sequence[Sequence[Tuple[int, Tuple[int, int]]]], num_inputs: int):
        self.F = F
        self.layer_sizes = [len(l) for l in layers]
        self.num_inputs = num_inputs
        sizes = np.array(self.layer_sizes, dtype=np.uint32)
        types = np.array([t for l in layers for t, _ in l], dtype=np.uint8)
        in0 = np.array([i[0] for l in layers for _, i in l], dtype=np.uint32)
        in1 = np.array([i[1] for l in layers for _, i in l], dtype=np.uint32)
        self._h = C.c_void_p()
        check(lib.scb_circuit_create(F._h, len(layers), sizes.ctypes.data_as(u32p), types.ctypes.data_as(u8p), in0.ctypes.data_as(u32p),
                                     in1.ctypes.data_as(u32p), num_inputs, C.byref(self._h)))

    @staticmethod
    def from_arrays(F: api.Field, layer_sizes: Sequence[int], types: np.ndarray, in0: np.ndarray, in1: np.ndarray, num_inputs: int) -> "Circuit":
        self = Circuit.__new__(Circuit)
        self.F, self.layer_sizes, self.num_inputs = F, list(layer_sizes), num_inputs
        sizes = np.array(self.layer_sizes, dtype=np.uint32)
        types, in0, in1 = (np.ascontiguousarray(types, dtype=np.uint8), np.ascontiguousarray(in0, dtype=np.uint32),
                           np.ascontiguousarray(in1, dtype=np.uint32))
        self._h = C.c_void_p()
        check(lib.scb_circuit_create(F._h, len(self.layer_sizes), sizes.ctypes.data_as(u32p), types.ctypes.data_as(u8p),
                                     in0.ctypes.data_as(u32p), in1.ctypes.data_as(u32p), num_inputs, C.byref(self._h)))
        return self

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.scb_circuit_free(self._h)
            self._h = None

    def layers_len(self) -> int:
        return len(self.layer_sizes)

    def num_vars_at(self, layer: int) -> int:
        o = C.c_uint32()
        check(lib.scb_circuit_num_vars_at(self._h, layer, C.byref(o)))
        return o.value

    def wiring_eval(self, layer: int, r_i: Sequence[int], b: Sequence[int], c: Sequence[int]) -> Tuple[int, int]:
        """(add~_i(r_i, b, c), mul~_i(r_i, b, c))."""
        F = self.F
        a_out, m_out = np.zeros((1, F.n), dtype=np.uint64), np.zeros((1, F.n), dtype=np.uint64)
        pad = lambda v: F.to_mont(list(v)) if len(v) else np.zeros((1, F.n), dtype=np.uint64)
        check(lib.scb_circuit_wiring_eval(self._h, layer, api._p64(pad(r_i)), api._p64(pad(b)), api._p64(pad(c)), api._p64(a_out), api._p64(m_out)))
        return F.from_mont(a_out)[0], F.from_mont(m_out)[0]


class GkrProver:
    """gkr_protocol::Prover (gkr-protocol/src/lib.rs:324-474)."""

    def __init__(self, circuit: Circuit, inp):
        self.F, self.circuit = circuit.F, circuit
        m = api._as_mont(self.F, inp)
        self._h = C.c_void_p()
        check(lib.scb_gkr_prover_new(circuit._h, api._p64(m), m.shape[0], C.byref(self._h)))
        self.i, self.k = 0, 0
        self.r: List[int] = []

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.scb_gkr_prover_free(self._h)
            self._h = None

    def layer(self, i: int) -> api.DenseMultilinearExtension:
        h = C.c_void_p()
        check(lib.scb_gkr_prover_layer(self._h, i, C.byref(h)))
        return api.DenseMultilinearExtension(self.F, h)

    def start_protocol(self):  # lib.rs:363-367
        return ("Begin", self.layer(0).to_evaluations())

    def start_round(self, i: int, r_i: Sequence[int]):  # lib.rs:373-436
        F = self.F
        c1 = np.zeros((1, F.n), dtype=np.uint64)
        nv = C.c_uint32()
        pt = F.to_mont(list(r_i)) if len(r_i) else np.zeros((1, F.n), dtype=np.uint64)
        check(lib.scb_gkr_prover_start_round(self._h, i, api._p64(pt), api._p64(c1), C.byref(nv)))
        self.i, self.k, self.r = i, nv.value // 2, []
        self._c_1 = F.from_mont(c1)[0]
        return ("StartSumCheck", self._c_1, i, nv.value)

    def c_1(self) -> int:
        return self._c_1

    def _round_poly(self, j: int) -> api.SparsePolynomial:
        F = self.F
        ev = np.zeros((3, F.n), dtype=np.uint64)
        r_prev = F.elem(self.r[j - 1]) if j else np.zeros((1, F.n), dtype=np.uint64)
        check(lib.scb_gkr_prover_round_evals(self._h, j, api._p64(r_prev), api._p64(ev)))
        return api.evals_to_univariate_mont(F, api.KIND_GKR_W, ev)

    def round_msg(self, j: int):  # lib.rs:439-456
        F = self.F
        if j == 2 * self.k - 1:
            p = self._round_poly(j)
            qe = np.zeros((self.k + 1, F.n), dtype=np.uint64)
            n = C.c_uint32()
            check(lib.scb_gkr_prover_restrict_evals(self._h, api._p64(F.elem(self.r[j])), api._p64(qe), self.k + 1, C.byref(n)))
            q = api.evals_to_univariate_mont(F, api.KIND_GKR_W, qe)  # restrict_poly's polynomial as the unique interpolant, zero terms dropped (the reference can keep explicit zero terms: DESIGN.md section 5)
            return ("FinalRoundMessage", p, q)
        return ("SumCheckProverMessage", self._round_poly(j))

    def receive_verifier_msg(self, msg) -> None:  # lib.rs:459-468
        if msg[0] == "SumCheckRoundResult":
            kind, val = msg[1]
            assert kind == "JthRound"
            self.r.append(val)

    def prove_layer(self, i: int, r_i: Sequence[int], challenges: Sequence[int]):
        """The whole proof of layer i when the verifier's 2k public-coin challenges are known up front: the messages
        start_round / round_msg(0..2k-1) would produce, from ONE library call (kernels launched back to back, two host
        waits).  Returns (StartSumCheck message, raw device results); ``layer_messages`` turns the latter into the
        reference's message sequence."""
        F = self.F
        k = self.circuit.num_vars_at(i + 1)
        assert len(challenges) == 2 * k
        pt = F.to_mont(list(r_i)) if len(r_i) else np.zeros((1, F.n), dtype=np.uint64)
        rs = F.to_mont(list(challenges))
        c1 = np.zeros((1, F.n), dtype=np.uint64)
        ev = np.zeros((2 * k * 3, F.n), dtype=np.uint64)
        qe = np.zeros((k + 1, F.n), dtype=np.uint64)
        nv = C.c_uint32()
        check(lib.scb_gkr_prover_prove_layer(self._h, i, api._p64(pt), api._p64(rs), 2 * k, api._p64(c1), api._p64(ev), api._p64(qe), k + 1, C.byref(nv)))
        self.i, self.k, self.r = i, k, list(challenges)
        self._c_1 = F.from_mont(c1)[0]
        return ("StartSumCheck", self._c_1, i, nv.value), (ev, qe)

    def layer_messages(self, raw):
        """Raw sums of ``prove_layer`` -> [SumCheckProverMessage x (2k-1), FinalRoundMessage] (lib.rs:439-456)."""
        F, k = self.F, self.k
        ev, qe = raw
        polys = [api.evals_to_univariate_mont(F, api.KIND_GKR_W, ev[3 * j:3 * j + 3]) for j in range(2 * k)]
        q = api.evals_to_univariate_mont(F, api.KIND_GKR_W, qe)
        return [("SumCheckProverMessage", pj) for pj in polys[:-1]] + [("FinalRoundMessage", polys[-1], q)]


class GkrVerifier:
    """gkr_protocol::Verifier (gkr-protocol/src/lib.rs:38-218); ``rng.draw()`` stands in for F::rand(rng)."""

    def __init__(self, circuit: Circuit):
        self.F, self.circuit = circuit.F, circuit
        self.r: List[List[int]] = []
        self.m: List[int] = []
        self.state = None

    def receive_prover_msg(self, msg, rng):  # lib.rs:177-207
        F, p = self.F, self.F.p
        kind = msg[0]
        if kind == "Begin":
            outs = msg[1]
            k0 = self.circuit.num_vars_at(0)
            d = api.DenseMultilinearExtension.from_evaluations_slice(F, k0, outs)
            r_zero = [rng.draw() for _ in range(k0)]
            self.r, self.m = [r_zero], [d.evaluate(r_zero)]
            return ("R", list(r_zero))
        if kind == "StartSumCheck":  # lib.rs:89-105 (add_i / mul_i stay a gate list; evaluated in final_round_message)
            _, c_1, rnd, num_vars = msg
            v = api.Verifier(num_vars, None, F)
            v.set_c_1(c_1)
            self.state = {"bc": [], "verifier": v, "round": rnd}
            return ("RoundStarted", rnd)
        if kind == "SumCheckProverMessage":  # lib.rs:121-137
            res = self.state["verifier"].round(msg[1], rng)
            if res[0] == "JthRound":
                self.state["bc"].append(res[1])
            return ("SumCheckRoundResult", res)
        if kind == "FinalRoundMessage":  # lib.rs:139-174
            _, pp, q = msg
            bc = self.state["bc"]
            half = len(bc) // 2
            q0, q1 = q.evaluate(0), q.evaluate(1)
            add_e, mul_e = self.circuit.wiring_eval(self.state["round"], self.r[-1], bc[:half], bc[half:])
            ev = (add_e * (q0 + q1) + mul_e * q0 * q1) % p
            assert ev == pp.evaluate(bc[-1]), (ev, pp.evaluate(bc[-1]))  # lib.rs:157 assert_eq!
            r = rng.draw()
            r_next = [(b + r * (c - b)) % p for b, c in zip(bc[:half], bc[half:])]  # line(b, c) evaluated at r, lib.rs:160-164
            self.r.append(r_next)
            self.m.append(q.evaluate(r))
            return ("R", list(r_next))
        raise ValueError(kind)

    def final_random_point(self, rng):  # lib.rs:108-119
        pt = rng.draw()
        self.state["bc"].append(pt)
        return ("SumCheckRoundResult", ("JthRound", pt))

    def check_input(self, inp) -> bool:  # lib.rs:210-217
        m = api._as_mont(self.F, inp)
        w = api.DenseMultilinearExtension.from_evaluations_slice(self.F, int(m.shape[0]).bit_length() - 1, m)
        return w.evaluate(self.r[-1]) == self.m[-1]
