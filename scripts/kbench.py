#!/usr/bin/env python
"""Kernel micro-benchmark (GPU box): times the round-message kernel and the fused fold+message kernel alone with
CUDA events and prints achieved algorithmic GB/s.  Tuning knobs come from the environment (SCB_UNROLL, SCB_BPS)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import thaler_study_b200 as T

ap = argparse.ArgumentParser()
ap.add_argument("--vars", type=int, default=28)
ap.add_argument("--tables", type=int, default=3)
ap.add_argument("--modulus", type=int, default=1572869)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
F = T.Field(a.modulus)
E = 8 * F.n
tabs = [T.DenseMultilinearExtension.synthetic(F, a.vars, 100 + k) for k in range(a.tables)]
g = T.ProductMLE.new(tabs)
d_out = torch.empty([a.tables + 1, F.n], dtype=torch.int64, device="cuda")
def timeit(fn):
    ts = []
    for i in range(3 + a.iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = fn(); e.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(s.elapsed_time(e))
        del r
    return sum(ts) / len(ts), min(ts)
n = 1 << a.vars
m_avg, m_min = timeit(lambda: g.round_evals_device(d_out.data_ptr()))
f_avg, f_min = timeit(lambda: g.fix_and_round_evals_device(12345 % a.modulus, d_out.data_ptr()))
print(json.dumps({"env": {k: os.environ.get(k) for k in ("SCB_UNROLL", "SCB_BPS")}, "vars": a.vars, "K": a.tables, "p_bits": F.bits,
                  "round_evals_ms": m_avg, "round_evals_GBs": a.tables * n * E / m_avg / 1e6,
                  "fold_round_ms": f_avg, "fold_round_GBs": 1.5 * a.tables * n * E / f_avg / 1e6, "fold_round_min_ms": f_min}))
