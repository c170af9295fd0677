#!/usr/bin/env python
"""Kernel micro-benchmark (GPU box): times the round-message kernel and the fused fold+message kernels alone with
CUDA events and prints achieved GB/s (bytes the kernel's own layout must move).  Tuning knobs come from the
environment (SCB_UNROLL, SCB_BPS, SCB_BPS32)."""
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
K, n = a.tables, 1 << a.vars
tabs = [T.DenseMultilinearExtension.synthetic(F, a.vars, 100 + k) for k in range(K)]
g = T.ProductMLE.new(tabs)
d_out = torch.empty([K + 1, F.n], dtype=torch.int64, device="cuda")
r = 12345 % a.modulus


def timeit(fn):
    ts = []
    for i in range(3 + a.iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); res = fn(); e.record(); torch.cuda.synchronize()
        if i >= 3:
            ts.append(s.elapsed_time(e))
        del res
    return sum(ts) / len(ts)


out = {"env": {k: os.environ.get(k) for k in ("SCB_UNROLL", "SCB_BPS", "SCB_BPS32")}, "vars": a.vars, "K": K, "p_bits": F.bits}
t = timeit(lambda: g.round_evals_device(d_out.data_ptr()))
out["round_evals"] = {"ms": t, "GBs": K * n * E / t / 1e6}
t = timeit(lambda: g.fix_and_round_evals_device(r, d_out.data_ptr()))
out["fold_round_u64_u64"] = {"ms": t, "GBs": 1.5 * K * n * E / t / 1e6}
if F.policy == 0:
    gp = g.clone().allow_packed(True)
    t = timeit(lambda: gp.fix_and_round_evals_device(r, d_out.data_ptr()))
    out["fold_round_u64_u32"] = {"ms": t, "GBs": K * n * (E + 2) / t / 1e6}
    g1 = gp.fix_and_round_evals_device(r, d_out.data_ptr())  # 2^(v-1) packed entries
    t = timeit(lambda: g1.fix_and_round_evals_device(r, d_out.data_ptr()))
    out["fold_round_u32_u32"] = {"ms": t, "GBs": K * (n // 2) * (4 + 2) / t / 1e6}
print(json.dumps(out))
