#!/usr/bin/env python
"""BASELINE.json configs[0..3] on one B200: each is run through the public API at its full size, checked through
a size-independent property (verifier acceptance, c_1 against an independent integer computation) and timed.
Prints one JSON object per config (committed under profiles/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import thaler_study_b200 as T  # noqa: E402

P21 = 1572869          # the reference's largest field (triangle-counting/src/lib.rs:272-277)
P28 = 268435361        # 28-bit prime, p = 1 mod 4 (two-adicity >= 2 like the reference's fields), > 6*triangles at n = 1024
BLS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def is_prime(n):
    if n < 2:
        return False
    for q in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


assert is_prime(P28) and P28 % 4 == 1


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), min(ts)


def cfg1(p=P21):
    F = T.Field(p)
    v = 20
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 11)])
    tr = T.generate_transcript(T.Prover(g))
    ok = T.verify_transcript(tr, T.Verifier(v, g))
    t_p, _ = timed(lambda: T.generate_transcript(T.Prover(g)))
    t_v, _ = timed(lambda: T.verify_transcript(tr, T.Verifier(v, g)))
    return {"config": "configs[0]: prover+verifier, random 20-variable multilinear (ProductMLE<1>)", "field_bits": F.bits, "verified": ok,
            "prover_ms": t_p * 1e3, "verifier_ms": t_v * 1e3, "prover_Melem_s": (1 << v) / t_p / 1e6}


def cfg2(p):
    F = T.Field(p)
    v = 24
    m = T.DenseMultilinearExtension.synthetic(F, v, 21)
    rng = np.random.default_rng(1)
    r = [int(x) % p for x in rng.integers(0, 2**62, size=v)]
    be = m.evaluate_be(r)
    le = m.evaluate(list(reversed(r)))
    # independent check: v successive folds (LSB-first) must give the same element
    folded = m.fix_variables(list(reversed(r))).to_evaluations()[0]
    t_dev, t_min = timed(lambda: m.evaluate_be(r), reps=10)
    # device-side span (first kernel start .. last kernel end), excluding the Python int <-> limb conversions
    evs = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m.evaluate_be(r)
        b.record()
        torch.cuda.synchronize()
        evs.append(a.elapsed_time(b))
    t_ev = sorted(evs)[len(evs) // 2] * 1e-3
    host = m.to_evaluations_mont()
    t_e2e, _ = timed(lambda: T.vsbw_multilinear_from_evaluations(F, host, r), reps=3, warm=1)
    nbytes = (1 << v) * 8 * F.n
    # raw C-ABI call with the point already in Montgomery limbs (what a Rust caller issues): no Python int <-> limb conversion
    from thaler_study_b200._lib import check, lib, u64p
    pt = F.to_mont(r)
    out = np.zeros((1, F.n), dtype=np.uint64)
    raw = []
    for _ in range(30):
        t0 = time.perf_counter()
        check(lib.scb_mle_evaluate_be(m._h, pt.ctypes.data_as(u64p), v, out.ctypes.data_as(u64p)))
        raw.append(time.perf_counter() - t0)
    t_raw = sorted(raw)[len(raw) // 2]
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # 4-limb fields: the integer pipe binds (64 wide multiply-adds per entry + row reductions at 9.31 T/s, profiles/r02_mont29.md)
    t_int = (1 << v) * 90 / 9.31e12 if F.n == 4 else 0.0
    roof = {"bound": "hbm", "achieved": nbytes / t_raw / 1e9, "peak": peak, "unit": "GB/s", "frac": nbytes / t_raw / 1e9 / peak, "traffic": None,
            "peak_source": peak_src, "call_ms": t_raw * 1e3, "t_hbm_bound_ms": nbytes / (peak * 1e9) * 1e3, "t_integer_bound_ms": t_int * 1e3 or None,
            "frac_of_slower_bound": max(nbytes / (peak * 1e9), t_int) / t_raw,
            "note": "whole synchronous C-ABI call (launch + kernel + wait), not the kernel alone; an empty launch + wait costs 9.5 us on these boxes"}
    return {"config": "configs[1]: MLE evaluation, 2^24 evals, random point (vsbw order), eq table by doubling", "field_bits": F.bits, "roofline": roof,
            "checks": {"be==le(reversed)": be == le, "eq_table==24_folds": be == folded},
            "device_resident_ms": t_dev * 1e3, "device_resident_GBs": nbytes / t_dev / 1e9, "best_ms": t_min * 1e3,
            "device_span_ms(cuda events)": t_ev * 1e3, "device_span_GBs": nbytes / t_ev / 1e9,
            "e2e_host_evals_ms": t_e2e * 1e3, "e2e_GBs": nbytes / t_e2e / 1e9, "algorithmic_bytes": nbytes}


def cfg3(p=P21):
    F = T.Field(p)
    n_bits, n = 10, 1024
    rng = np.random.default_rng(3)
    a = rng.integers(0, p, size=(n, n), dtype=np.int64)
    b = rng.integers(0, p, size=(n, n), dtype=np.int64)
    a_m, b_m = F.to_mont(a.reshape(-1).tolist()), F.to_mont(b.reshape(-1).tolist())
    i, j = 517, 33
    point = [(i >> t) & 1 for t in range(n_bits)] + [(j >> t) & 1 for t in range(n_bits)]
    g = T.MatMulG.new(F, n_bits, a_m, b_m, point)
    want = int(sum(int(x) * int(y) for x, y in zip(a[i, :], b[:, j])) % p)  # (A*B)[i][j], matrix-multiplication/src/lib.rs:339-340
    c1 = T.Prover(g).c_1()
    rpoint = [int(x) % p for x in rng.integers(0, 2**62, size=2 * n_bits)]
    g2 = T.MatMulG.new(F, n_bits, a_m, b_m, rpoint)
    tr = T.generate_transcript(T.Prover(g2))
    ok = T.verify_transcript(tr, T.Verifier(n_bits, g2))
    t_setup, _ = timed(lambda: T.MatMulG.new(F, n_bits, a_m, b_m, rpoint), reps=3, warm=1)
    t_prove, _ = timed(lambda: T.generate_transcript(T.Prover(g2)))
    return {"config": "configs[2]: matrix-multiplication sum-check, n = 1024", "field_bits": F.bits,
            "checks": {"c_1==(A*B)[i][j]": c1 == want, "verified_random_point": ok},
            "G_new_setup_ms(host matrices: H2D 16 MB + one-pass eq-table fixes)": t_setup * 1e3, "sumcheck_10_rounds_ms": t_prove * 1e3}


def cfg4(p=P28):
    F = T.Field(p)
    n_bits, n = 10, 1024
    rng = np.random.default_rng(4)
    up = np.triu(rng.integers(0, 2, size=(n, n), dtype=np.int64), 1)
    adj = up + up.T
    tri6 = int(np.trace(np.linalg.matrix_power(adj.astype(np.float64), 3)))  # = 6 * triangles (exact in float64 here)
    a64 = adj.astype(np.int64)
    tri6_exact = int(((a64 @ a64) * a64).sum())
    g = T.TriangleG.new_adj_matrix(F, 2 * n_bits, adj.reshape(-1).astype(bool).tolist())
    t0 = time.perf_counter()
    prover = T.Prover(g)
    c1 = prover.c_1()
    tr = T.generate_transcript(prover)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    ok = T.verify_transcript(tr, T.Verifier(3 * n_bits, g))
    t_prove, _ = timed(lambda: T.generate_transcript(T.Prover(g)), reps=3, warm=1)
    return {"config": "configs[3]: triangle-counting sum-check, random 1024-node graph (30 rounds)", "field_bits": F.bits, "modulus": p,
            "checks": {"c_1==6*triangles": c1 == tri6_exact % p and tri6_exact < p, "float_check": tri6 == tri6_exact, "verified": ok},
            "triangles": tri6_exact // 6, "prove_ms": t_prove * 1e3, "first_call_ms": t_first * 1e3,
            "roofline": {"bound": "integer + latency", "t_integer_bound_ms": float(n) ** 3 / 9.31e12 * 1e3,
                         "note": "n^3 wide multiply-adds of the one field matmul (M = f2 f1, tri.cuh) at the measured 9.31 T/s multiplier peak; the matmul launch takes "
                                 "0.216 ms (profiles/r02_launches_triangle_mle.csv), the other ~0.8 ms are 30 Fiat-Shamir rounds of 3-4 dependent small launches each "
                                 "(~9.5 us launch + wait floor per synchronous call, profiles/r02_launch_latency.jsonl)",
                         "frac_of_integer_bound": float(n) ** 3 / 9.31e12 / t_prove}}


if __name__ == "__main__":
    out = [cfg1(), cfg2(P21), cfg2(BLS), cfg3(), cfg4()]
    for o in out:
        print(json.dumps(o))
