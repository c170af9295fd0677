#!/usr/bin/env python
"""Times the stand-alone pair-pass kernels (csrc/pairs.cuh) with CUDA events: k_grid_sp (pass A) and k_pair_pass_sp on
u64 and on packed u32 input.  Usage: python scripts/kbench_pairs.py [vars] [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import thaler_study_b200 as T

v = int(sys.argv[1]) if len(sys.argv) > 1 else 28
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(os.environ.get("REPS", "5"))
F = T.Field(1572869)
g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 0xB200 + k) for k in range(K)])
T.synchronize()


def timed(fn, n=reps):
    fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts), out


ms, _ = timed(lambda: g.grid_evals())
print(f"k_grid_sp<{K},u64>      2^{v}: {ms:.4f} ms  {K * (1 << v) * 8 / ms / 1e6:.0f} GB/s")
ms, (g2, _) = timed(lambda: g.pair_pass(12345, 67890))
print(f"k_pair_pass_sp<{K},u64> 2^{v}: {ms:.4f} ms  {K * ((1 << v) * 8 + (1 << (v - 2)) * 4) / ms / 1e6:.0f} GB/s")
ms, _ = timed(lambda: g2.grid_evals())
print(f"k_grid_sp<{K},u32>      2^{v-2}: {ms:.4f} ms  {K * (1 << (v - 2)) * 4 / ms / 1e6:.0f} GB/s")
ms, (g3, _) = timed(lambda: g2.pair_pass(12345, 67890))
print(f"k_pair_pass_sp<{K},u32> 2^{v-2}: {ms:.4f} ms  {K * ((1 << (v - 2)) * 4 + (1 << (v - 4)) * 4) / ms / 1e6:.0f} GB/s")
ms, _ = timed(lambda: g.round_evals())
print(f"k_round_evals<{K}>      2^{v}: {ms:.4f} ms  {K * (1 << v) * 8 / ms / 1e6:.0f} GB/s")
