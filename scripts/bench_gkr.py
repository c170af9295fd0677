#!/usr/bin/env python
"""BASELINE configs[4], second half: GKR on a layered circuit of width 2^20 and depth 16 (random gates, random wiring).
The reference's dense formulation needs 2^60-entry wiring tables per layer and cannot run this; the gate-list prover's
messages are the same polynomials (tests/test_gpu_gkr.py).  Checks: the reference's verifier logic accepts every
layer and check_input holds.  Prints one JSON object."""
import argparse, json, os, random, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import thaler_study_b200 as T
from thaler_study_b200.gkr import Circuit, GkrProver, GkrVerifier

T.options_from_env()  # harness opt-in: SCB_<OPTION>=n -> scb_set_option

ap = argparse.ArgumentParser()
ap.add_argument("--width-bits", type=int, default=20)
ap.add_argument("--depth", type=int, default=16)
ap.add_argument("--modulus", type=int, default=1572869)
a = ap.parse_args()
p, wb, depth = a.modulus, a.width_bits, a.depth
F = T.Field(p)
S = 1 << wb
rng = np.random.default_rng(2024)
t0 = time.perf_counter()
circ = Circuit.from_arrays(F, [S] * depth, rng.integers(0, 2, size=S * depth, dtype=np.uint8), rng.integers(0, S, size=S * depth, dtype=np.uint32),
                           rng.integers(0, S, size=S * depth, dtype=np.uint32), S)
t_circuit = time.perf_counter() - t0
inp = F.to_mont([int(x) % p for x in rng.integers(0, 2**62, size=S)])


class Rng:
    def __init__(self, seed):
        self.r = random.Random(seed)

    def draw(self):
        return self.r.randrange(p)


def run(verify=True):
    rnd = Rng(1)
    torch.cuda.synchronize()
    t = {"evaluate": 0.0, "prover": 0.0, "verifier": 0.0}
    t0 = time.perf_counter()
    prover = GkrProver(circ, inp)
    torch.cuda.synchronize()
    t["evaluate"] = time.perf_counter() - t0
    begin = prover.start_protocol()
    verifier = GkrVerifier(circ)
    kind, r_i = verifier.receive_prover_msg(begin, rnd)
    for i in range(depth):
        t0 = time.perf_counter()
        msg = prover.start_round(i, r_i)
        t["prover"] += time.perf_counter() - t0
        nv = 2 * circ.num_vars_at(i + 1)
        verifier.receive_prover_msg(msg, rnd)
        for j in range(nv - 1):
            t0 = time.perf_counter()
            pm = prover.round_msg(j)
            t["prover"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            vm = verifier.receive_prover_msg(pm, rnd)
            t["verifier"] += time.perf_counter() - t0
            prover.receive_verifier_msg(vm)
        prover.receive_verifier_msg(verifier.final_random_point(rnd))
        t0 = time.perf_counter()
        pm = prover.round_msg(nv - 1)
        t["prover"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        kind, r_i = verifier.receive_prover_msg(pm, rnd)
        t["verifier"] += time.perf_counter() - t0
    t0 = time.perf_counter()
    ok = verifier.check_input(inp)
    t["verifier"] += time.perf_counter() - t0
    return ok, t


class Replay:
    def __init__(self, values):
        self.values, self.pos = list(values), 0

    def draw(self):
        v = self.values[self.pos]
        self.pos += 1
        return v


def run_batched():
    """Public coins known up front: one library call per layer (scb_gkr_prover_prove_layer); the reference's verifier
    logic then replays the messages against the same challenges."""
    rnd = Rng(1)
    torch.cuda.synchronize()
    t = {"prover": 0.0, "format": 0.0, "verifier": 0.0}
    prover = GkrProver(circ, inp)
    verifier = GkrVerifier(circ)
    kind, r_i = verifier.receive_prover_msg(prover.start_protocol(), rnd)
    for i in range(depth):
        k = circ.num_vars_at(i + 1)
        ch = [rnd.draw() for _ in range(2 * k)]
        t0 = time.perf_counter()
        start, raw = prover.prove_layer(i, r_i, ch)
        t["prover"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        msgs = prover.layer_messages(raw)
        t["format"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        replay = Replay(ch)
        verifier.receive_prover_msg(start, replay)
        for m in msgs[:-1]:
            verifier.receive_prover_msg(m, replay)
        verifier.final_random_point(replay)
        kind, r_i = verifier.receive_prover_msg(msgs[-1], rnd)
        t["verifier"] += time.perf_counter() - t0
    return verifier.check_input(inp), t


ok, _ = run()
T.launch_count(reset=True)
ok2, t = run()
launches = T.launch_count()
okb, _ = run_batched()
T.launch_count(reset=True)
okb2, tb = run_batched()
launches_b = T.launch_count()
gates = S * depth
print(json.dumps({"config": f"configs[4]b: GKR, layered circuit width 2^{wb}, depth {depth}, random add/mul gates and wiring, field bits {F.bits}",
                  "verified": bool(ok and ok2), "gates": gates, "sumcheck_rounds": depth * 2 * wb,
                  "circuit_upload_and_csr_s": t_circuit, "evaluate_ms": t["evaluate"] * 1e3, "prover_ms": t["prover"] * 1e3,
                  "verifier_ms": t["verifier"] * 1e3, "prover_Mgates_per_s": gates / t["prover"] / 1e6, "gpu_launches": launches,
                  "batched": {"note": "challenges of a layer handed over up front (scb_gkr_prover_prove_layer): each phase is one "
                                      "cooperative launch, the k+1 line evaluations share one pass per 8 points, one host wait per layer", "verified": bool(okb and okb2), "prover_ms": tb["prover"] * 1e3,
                              "format_ms": tb["format"] * 1e3, "verifier_ms": tb["verifier"] * 1e3,
                              "prover_Mgates_per_s": gates / tb["prover"] / 1e6, "gpu_launches": launches_b}}))
