#!/usr/bin/env python
"""The reference's own criterion bench, in its shape (matrix-multiplication/benches/mm_benchmark.rs:62-98): group
`prover/prove/{num_vars}` for p = 2..15 over F_5 -- n = 2^p random matrices, G::new(p, a, b, point = bits of (2, 2)) in
UNTIMED set-up, then each iteration = g.clone() + Prover::new + num_vars x Prover::round with fresh random challenges;
throughput unit = num_vars (criterion's Throughput::Elements(num_vars)).  Here the iteration runs on the engine through
the Prover mirror (scb_prover_new / scb_prover_round), and beside it the CPU port (oracle.c, one thread = the
reference's own threading) does the same iteration.  Also reports F_1572869.  One JSON line per (field, p).

Matrices of n = 2^15 would be 2^30 entries each (the reference builds them too: 8 GiB per matrix at 8 B); p is capped
at SWEEP_MAX_P (default 13: 2^26-entry matrices) to keep the set-up within host memory and a minute of wall time."""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import thaler_study_b200 as T  # noqa: E402
from oracle.coracle import CField  # noqa: E402

MAX_P = int(os.environ.get("SWEEP_MAX_P", "13"))
SAMPLES = 10  # criterion sample_size(10)


def u32_to_boolean_vec(v, bits):
    return [(v >> i) & 1 for i in range(bits)]


for modulus in (5, 1572869):
    F, cf = T.Field(modulus), CField(modulus)
    rnd = random.Random(modulus)
    rng = np.random.default_rng(modulus)
    for p in range(2, MAX_P + 1):
        n = 1 << p
        a = rng.integers(0, modulus, size=n * n, dtype=np.uint64).reshape(-1, 1)
        b = rng.integers(0, modulus, size=n * n, dtype=np.uint64).reshape(-1, 1)
        a_m = np.ascontiguousarray(cf.to_mont(a[:, 0].tolist())) if n * n <= 1 << 16 else None
        if a_m is None:  # large matrices: Montgomery form by one modular multiplication with R (vectorised)
            R = pow(2, 64, modulus)
            a_m = (a * np.uint64(R) % np.uint64(modulus)).astype(np.uint64)
            b_m = (b * np.uint64(R) % np.uint64(modulus)).astype(np.uint64)
        else:
            b_m = np.ascontiguousarray(cf.to_mont(b[:, 0].tolist()))
        point = u32_to_boolean_vec(2, p) + u32_to_boolean_vec(2, p)
        g = T.MatMulG.new(F, p, a_m, b_m, point)      # untimed set-up, as in the reference
        num_vars = g.num_vars()
        assert num_vars == p

        def iteration():
            prover = T.Prover(g.clone())
            r_j = 1
            for j in range(num_vars):
                prover.round(r_j, j)
                r_j = rnd.randrange(modulus)

        iteration()
        torch.cuda.synchronize()
        ts = []
        for _ in range(SAMPLES):
            t0 = time.perf_counter()
            iteration()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        # the CPU port on the same tables (f_a, f_b after G::new), one thread
        fa, fb = g.table(0).to_evaluations_mont(), g.table(1).to_evaluations_mont()
        ch = cf.to_mont([rnd.randrange(modulus) for _ in range(max(num_vars - 1, 1))])
        t0 = time.perf_counter()
        for _ in range(SAMPLES):
            cf.product_prove([fa, fb], ch, 3, threads=1)
        cpu = (time.perf_counter() - t0) / SAMPLES
        med = sorted(ts)[len(ts) // 2]
        print(json.dumps({"group": "prover/prove", "field": f"F_{modulus}", "num_vars": num_vars, "matrix_n": n, "samples": SAMPLES,
                          "engine_us_per_iter_median": round(med * 1e6, 1), "engine_elements_per_s(num_vars/iter)": round(num_vars / med, 1),
                          "cpu_port_1_thread_us_per_iter": round(cpu * 1e6, 1),
                          "note": "sum-check over p variables (2^p-entry tables): launch-latency bound on the device at these sizes"}), flush=True)
        del g
