#!/usr/bin/env python
"""Workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): one small instance of every kernel
family of the hot path, each checked against the C oracle so that a sanitizer-clean run is also a correct one.

  compute-sanitizer --tool racecheck python scripts/sanitize_run.py [--resident]

Without --resident every pass is an ordinary launch (options pair_resident = 0, tail_vars = 0: the sanitizer slows the
device by 10-100x, and the resident kernels' 250 ms mailbox patience would only exercise their fallback); with it the
resident kernels run too (their shared-memory trees and grid barriers are what racecheck / synccheck look at)."""
import argparse
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import thaler_study_b200 as T  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from oracle.coracle import CField  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--resident", action="store_true")
ap.add_argument("--variants", action="store_true", help="also the cp.async / TMA variants of the pair kernels")
ap.add_argument("--only-g4", action="store_true", help="only the 4-limb kernels, every generation")
a = ap.parse_args()

BLS = O.BLS12_381_FR.p
BN254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617
checks = 0


def ok(cond, what):
    global checks
    checks += 1
    if not cond:
        print("MISMATCH:", what)
        sys.exit(1)


def product_case(p, v, K, label):
    OF, F, cf = O.Field(p), T.Field(p), CField(p)
    seeds = [40 + k for k in range(K)]
    tabs_c = [cf.synth(s, 0, 1 << v) for s in seeds]
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, s) for s in seeds])
    rnd = random.Random(v * 7 + K)
    ch = [rnd.randrange(p) for _ in range(v - 1)]
    c1, ev = cf.product_prove(tabs_c, cf.to_mont(ch), K + 1, threads=2)
    ok(g.sum() == cf.from_mont(c1)[0], f"{label}: sum")
    ok(g.round_evals() == cf.from_mont(ev[0]), f"{label}: round 0")
    cur, prev = g, cf.from_mont(ev[0])
    for j in range(1, min(v, 6)):
        claim = T.evals_to_univariate(F, T.KIND_PRODUCT, prev).evaluate(ch[j - 1])
        cur, got = cur.fix_and_round_evals(ch[j - 1], claim=claim)
        ok(got == cf.from_mont(ev[j]), f"{label}: fused round {j}")
        prev = got
    # whole proof through the Fiat-Shamir driver (pair passes / resident kernels depending on the options) + verifier
    tr = T.generate_transcript(T.Prover(g))
    ok(T.verify_transcript(tr, T.Verifier(v, g)), f"{label}: verify")
    if v <= 10:
        vals = [cf.from_mont(t) for t in tabs_c]
        want = O.generate_transcript(OF, O.Prover(O.ProductMLE(OF, [O.DenseMLE(OF, v, t) for t in vals])))
        ok(tr == want, f"{label}: transcript bytes")
    # single fold, MLE evaluation (one launch + the two-launch form), relabel
    r = ch[0]
    m = T.DenseMultilinearExtension.synthetic(F, v, seeds[0])
    ok(np.array_equal(m.fix_variables([r]).to_evaluations_mont(), cf.fix_variable(tabs_c[0], cf.to_mont([r]))), f"{label}: fold")
    pt = [rnd.randrange(p) for _ in range(v)]
    want_e = cf.from_mont(cf.mle_evaluate_le(tabs_c[0], cf.to_mont(pt)))[0]
    for fused in (1, 0):
        T.set_option("mle_fused", fused)
        ok(m.evaluate(pt) == want_e, f"{label}: mle eval fused={fused}")
    T.set_option("mle_fused", 1)


def main():
    if not a.resident:
        T.set_option("pair_resident", 0)
        T.set_option("tail_vars", 0)
    variants = [{}]
    if a.variants:
        variants += [{"pair_stage": 1}, {"grid_tma": 1}, {"grid_pf": 0}, {"pairs": 0}, {"packed": 0}, {"g4_kernel": 0}, {"g4_kernel": 1}, {"g4_p0one": 0}]
    if a.only_g4:  # the 4-limb kernels alone: wide accumulators (p = 1 mod 2^32 variant and generic), carry chains, radix 2^29
        for var in ({}, {"g4_p0one": 0}, {"g4_kernel": 1}, {"g4_kernel": 2}):
            for k_, v_ in var.items():
                T.set_option(k_, v_)
            tag = ",".join(f"{k}={v}" for k, v in var.items()) or "default"
            for p, v, K in ((BLS, 11, 3), (BLS, 7, 2), (BLS, 9, 4), (BN254, 10, 3)):
                product_case(p, v, K, f"[{tag}] p{p.bit_length()} v{v} K{K}")
            T.reset_options()
            T.set_option("pair_resident", 0)
            T.set_option("tail_vars", 0)
        print(f"sanitize_run (4-limb only): {checks} checks passed")
        sys.exit(0)
    for var in variants:
        for k_, v_ in var.items():
            T.set_option(k_, v_)
        tag = ",".join(f"{k}={v}" for k, v in var.items()) or "default"
        for p, v, K in ((1572869, 13, 3), (1572869, 9, 4), (389, 10, 2), (5, 8, 3), (0xFFFFFFFF00000001, 11, 3), (BLS, 11, 3), (BLS, 7, 2)):
            product_case(p, v, K, f"[{tag}] p{p.bit_length()} v{v} K{K}")
        for k_ in var:
            T.set_option(k_, {"grid_pf": 1, "pairs": 1, "packed": 1, "g4_kernel": 3, "g4_p0one": 1}.get(k_, 0))
    # matmul G::new (relabel / eq-table fixes), triangle (tiled matmul + folds), GKR W
    OF, F = O.FP1572869, T.Field(1572869)
    rnd = random.Random(3)
    n = 3
    A = [rnd.randrange(OF.p) for _ in range(1 << (2 * n))]
    B = [rnd.randrange(OF.p) for _ in range(1 << (2 * n))]
    pt = [rnd.randrange(OF.p) for _ in range(2 * n)]
    og, dg = O.MatMulG.new(OF, n, A, B, pt), T.MatMulG.new(F, n, A, B, pt)
    ok(T.generate_transcript(T.Prover(dg)) == O.generate_transcript(OF, O.Prover(og)), "matmul G transcript")
    adj = [rnd.randrange(2) for _ in range(64)]
    for tiled in (1, 0):
        T.set_option("tri_tiled", tiled)
        ot, dt = O.TriangleG.new_adj_matrix(OF, 6, adj), T.TriangleG.new_adj_matrix(F, 6, adj)
        ok(T.generate_transcript(T.Prover(dt)) == O.generate_transcript(OF, O.Prover(ot)), f"triangle transcript tiled={tiled}")
    T.set_option("tri_tiled", 1)
    from thaler_study_b200.gkr import Circuit, GkrProver, GkrVerifier

    layers = [[(rnd.choice((0, 1)), (rnd.randrange(8), rnd.randrange(8))) for _ in range(4)],  # 0 = Add, 1 = Mul
              [(rnd.choice((0, 1)), (rnd.randrange(8), rnd.randrange(8))) for _ in range(8)]]
    try:
        circ = Circuit(F, layers, 8)
        inp = [rnd.randrange(OF.p) for _ in range(8)]
        prover, verifier = GkrProver(circ, inp), GkrVerifier(circ)

        class R:
            def draw(self):
                return rnd.randrange(OF.p)

        rr = R()
        kind, r_i = verifier.receive_prover_msg(prover.start_protocol(), rr)
        for i in range(2):
            msg = prover.start_round(i, r_i)
            nv = 2 * circ.num_vars_at(i + 1)
            verifier.receive_prover_msg(msg, rr)
            for j in range(nv - 1):
                vm = verifier.receive_prover_msg(prover.round_msg(j), rr)
                prover.receive_verifier_msg(vm)
            prover.receive_verifier_msg(verifier.final_random_point(rr))
            kind, r_i = verifier.receive_prover_msg(prover.round_msg(nv - 1), rr)
        ok(verifier.check_input(inp), "gkr check_input")
    except TypeError as ex:  # layer description format differs: not a kernel problem
        print("gkr section skipped:", ex)
    T.synchronize()
    print(f"SANITIZE_WORKLOAD_OK checks={checks} resident={a.resident} variants={a.variants}")


if __name__ == "__main__":
    main()
