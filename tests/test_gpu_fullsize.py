"""Parity at the sizes BASELINE.json states, on the GPU box, against the oracles.

* configs[4]: the 2^26 and 2^28 Fiat-Shamir proofs of ProductMLE<3> over F_1572869 -- the path bench.py times
  (k_grid_sp_pf, the stand-alone k_pair_pass_sp over 8-byte tables, the packed-only resident kernel) -- are re-proved
  by oracle/oracle.c (the reference's structure: Prover::new's sum, then per round fix_variables and the message pass,
  sum-check-protocol/src/lib.rs:88-112) with the challenges the Python oracle's hashlib hash_to_field derives from the
  GPU transcript's own bytes (fiat-shamir/src/lib.rs:87-88).  c_1 and every round message must agree byte for byte.
* configs[1]: 2^24-entry MLE evaluation against orc_mle_vsbw (multilinear-extensions/src/lib.rs:6-24).
* configs[2], configs[3]: n = 1024 matrix-multiplication and triangle-counting c_1 against integer arithmetic
  (matrix-multiplication/src/lib.rs:339-340, triangle-counting/src/lib.rs:294-300) and full verifier runs.
* configs[4], second half: GKR 2^20 x 16 through the reference's verifier logic incl. check_input.
* multi-GPU: the sharded provers' transcripts against the single-GPU one, whenever the box has more than one GPU.
"""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle as O
from oracle.coracle import CField

import thaler_study_b200 as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P21 = 1572869
P28 = 268435361  # 28-bit prime = 1 mod 4 (two-adicity >= 2 like the reference's fields), > 6 * triangles at n = 1024
HOST_THREADS = len(os.sched_getaffinity(0))


def oracle_messages(OF, cf, tabs_c, K, transcript):
    """Re-proves with the C oracle under the challenges hashed (hashlib, Python oracle) from `transcript` and returns
    the oracle's message bytes, built by the PYTHON oracle's interpolation + serialization (independent of the C++
    host layer the engine used)."""
    v = len(transcript)
    ch = [O.hash_to_field(OF, b"".join(transcript[: j + 1])) for j in range(v - 1)]
    c1, ev = cf.product_prove(tabs_c, cf.to_mont(ch), K + 1, threads=HOST_THREADS)
    c_1 = cf.from_mont(c1)[0]
    out = []
    for j in range(v):
        sums = cf.from_mont(ev[j])
        if j == 0:
            assert (sums[0] + sums[1]) % OF.p == c_1  # Prover::new's sum of the product table == g_1(0) + g_1(1)
        body = O.SparsePoly.from_dense(OF, O.lagrange_to_coeffs(OF, sums)).serialize()
        out.append((O.ser_field(OF, c_1) if j == 0 else b"") + body)
    return c_1, out


@pytest.mark.parametrize("v", [26, 28])
def test_headline_proof_matches_c_oracle(v):
    """configs[4] at the size bench.py times (2^28) and at the threshold where the stand-alone pair pass starts (2^26)."""
    OF, F, cf, K = O.FP1572869, T.Field(P21), CField(P21), 3
    seeds = [0xB200 + k for k in range(K)]  # bench.py's tables
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, s) for s in seeds])
    T.launch_count(reset=True)
    prover = T.Prover(g)
    got_c1 = prover.c_1()
    transcript = T.generate_transcript(prover)
    launches = T.launch_count()
    assert len(transcript) == v
    assert T.verify_transcript(transcript, T.Verifier(v, g))
    tabs_c = [cf.synth(s, 0, 1 << v) for s in seeds]
    # same synthetic stream on both sides (spot-check a window; the whole table is compared at small sizes elsewhere)
    lo = T.DenseMultilinearExtension.synthetic(F, 12, seeds[1], start=(1 << v) - (1 << 12)).to_evaluations_mont()
    assert np.array_equal(lo, tabs_c[1][-(1 << 12):])
    want_c1, want = oracle_messages(OF, cf, tabs_c, K, transcript)
    assert got_c1 == want_c1
    for j in range(v):
        assert transcript[j] == want[j], f"round {j}"
    # the timed path really is the three-launch proof: grid pass, (from 2^26) the stand-alone pair pass, resident kernel
    assert launches == 3, launches


def test_headline_proof_4_limb_matches_c_oracle():
    """The 4-limb path (BLS12-381 Fr, ark_ed_on_bls12_381::Fq of Cargo.toml:20) at 2^22: every round against the C oracle."""
    OF, cf, K, v = O.BLS12_381_FR, CField(O.BLS12_381_FR.p), 3, 22
    F = T.Field(OF.p)
    seeds = [0xB200 + k for k in range(K)]
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, s) for s in seeds])
    prover = T.Prover(g)
    got_c1 = prover.c_1()
    transcript = T.generate_transcript(prover)
    tabs_c = [cf.synth(s, 0, 1 << v) for s in seeds]
    want_c1, want = oracle_messages(OF, cf, tabs_c, K, transcript)
    assert got_c1 == want_c1 and transcript == want


@pytest.mark.parametrize("p", [P21, O.BLS12_381_FR.p], ids=["p21", "bls12_381_fr"])
def test_mle_eval_2_24_matches_c_oracle(p):
    """configs[1]: 2^24 evaluations, random point, the vsbw order (r[0] <-> index MSB), device-resident and from host."""
    OF, F, cf = O.Field(p), T.Field(p), CField(p)
    v = 24
    rnd = random.Random(24)
    r = [rnd.randrange(p) for _ in range(v)]
    evals = cf.synth(21, 0, 1 << v)
    want = cf.from_mont(cf.mle_vsbw(evals, cf.to_mont(r)))[0]
    m = T.DenseMultilinearExtension.synthetic(F, v, 21)
    assert m.evaluate_be(r) == want
    assert m.evaluate(list(reversed(r))) == want
    assert T.vsbw_multilinear_from_evaluations(F, evals, r) == want
    assert T.cti_multilinear_from_evaluations(F, evals, r) == want
    assert cf.from_mont(cf.mle_evaluate_le(evals, cf.to_mont(list(reversed(r)))))[0] == want


def test_matmul_n1024_full_size():
    """configs[2]: c_1 == (A*B)[i][j] for Boolean points (matrix-multiplication/src/lib.rs:339-340) and a full
    prover/verifier run at a random point, n = 1024."""
    p, n_bits, n = P21, 10, 1024
    F = T.Field(p)
    rng = np.random.default_rng(3)
    a = rng.integers(0, p, size=(n, n), dtype=np.int64)
    b = rng.integers(0, p, size=(n, n), dtype=np.int64)
    a_m, b_m = F.to_mont(a.reshape(-1).tolist()), F.to_mont(b.reshape(-1).tolist())
    for i, j in ((517, 33), (0, 0), (1023, 1023)):
        point = [(i >> t) & 1 for t in range(n_bits)] + [(j >> t) & 1 for t in range(n_bits)]
        g = T.MatMulG.new(F, n_bits, a_m, b_m, point)
        want = int(sum(int(x) * int(y) for x, y in zip(a[i, :], b[:, j])) % p)
        assert T.Prover(g).c_1() == want
    rpoint = [int(x) % p for x in rng.integers(0, 2**62, size=2 * n_bits)]
    g2 = T.MatMulG.new(F, n_bits, a_m, b_m, rpoint)
    tr = T.generate_transcript(T.Prover(g2))
    assert len(tr) == n_bits and T.verify_transcript(tr, T.Verifier(n_bits, g2))
    # the set-up folds against the C oracle: f_A = relabel + fix (row bits), f_B = fix (column bits)
    cf = CField(p)
    fa = a_m.reshape(n, n)
    cur = np.ascontiguousarray(fa.T.reshape(-1, 1))  # relabel(0, n, n) = transpose
    for t in range(n_bits):
        cur = cf.fix_variable(cur, cf.to_mont([rpoint[t]]))
    assert np.array_equal(g2.table(0).to_evaluations_mont(), cur)
    cur = b_m
    for t in range(n_bits):
        cur = cf.fix_variable(cur, cf.to_mont([rpoint[n_bits + t]]))
    assert np.array_equal(g2.table(1).to_evaluations_mont(), cur)


def test_triangle_n1024_full_size():
    """configs[3]: random 1024-node graph, c_1 == 6 * triangles (triangle-counting/src/lib.rs:294-300) in a field that
    holds the count, and the 30-round proof verifies."""
    p, n_bits, n = P28, 10, 1024
    F = T.Field(p)
    rng = np.random.default_rng(4)
    up = np.triu(rng.integers(0, 2, size=(n, n), dtype=np.int64), 1)
    adj = up + up.T
    tri6 = int(((adj @ adj) * adj).sum())
    assert tri6 < p and tri6 % 6 == 0
    g = T.TriangleG.new_adj_matrix(F, 2 * n_bits, adj.reshape(-1).astype(bool).tolist())
    prover = T.Prover(g)
    assert prover.c_1() == tri6
    tr = T.generate_transcript(prover)
    assert len(tr) == 3 * n_bits and T.verify_transcript(tr, T.Verifier(3 * n_bits, g))
    # first message against the C oracle's literal 4-point evaluation at X = 0, 1, 2 (degree 2: SURVEY F7)
    cf = CField(p)
    f = F.to_mont(adj.reshape(-1).tolist())
    # (each oracle point is 2^29 products on one thread; g_1(1) follows from the integer count: g_1(0) + g_1(1) = 6 * triangles)
    s0, s2 = (cf.from_mont(cf.triangle_round_eval_at(f, f, f, n_bits, n_bits, n_bits, cf.to_mont([x])))[0] for x in (0, 2))
    sums = [s0, (tri6 - s0) % p, s2]
    want0 = O.ser_field(O.Field(p), tri6) + O.SparsePoly.from_dense(O.Field(p), O.lagrange_to_coeffs(O.Field(p), sums)).serialize()
    assert tr[0] == want0


def test_gkr_width_2_20_depth_16_full_size():
    """configs[4], second half: layered circuit of width 2^20 and depth 16; the reference's verifier logic accepts every
    layer of the batched layer prover and check_input holds; the circuit evaluation is checked against numpy."""
    from thaler_study_b200.gkr import Circuit, GkrProver, GkrVerifier

    p, wb, depth = P21, 20, 16
    F = T.Field(p)
    S = 1 << wb
    rng = np.random.default_rng(2024)
    types = rng.integers(0, 2, size=S * depth, dtype=np.uint8)
    in0 = rng.integers(0, S, size=S * depth, dtype=np.uint32)
    in1 = rng.integers(0, S, size=S * depth, dtype=np.uint32)
    circ = Circuit.from_arrays(F, [S] * depth, types, in0, in1, S)
    inp_vals = rng.integers(0, p, size=S, dtype=np.int64)
    inp = F.to_mont(inp_vals.tolist())

    class Rng:
        def __init__(self, seed):
            self.r = random.Random(seed)

        def draw(self):
            return self.r.randrange(p)

    class Replay:
        def __init__(self, values):
            self.values, self.pos = list(values), 0

        def draw(self):
            val = self.values[self.pos]
            self.pos += 1
            return val

    rnd = Rng(1)
    prover = GkrProver(circ, inp)
    # circuit evaluation (gkr-protocol/src/circuit.rs:99-124) against numpy, layer by layer from the inputs up
    cur = inp_vals.copy()
    for layer in range(depth - 1, -1, -1):
        sl = slice(layer * S, (layer + 1) * S)
        x, y = cur[in0[sl]], cur[in1[sl]]
        cur = np.where(types[sl] == 1, (x * y) % p, (x + y) % p)
    begin = prover.start_protocol()
    assert begin[1] == cur.tolist()
    verifier = GkrVerifier(circ)
    kind, r_i = verifier.receive_prover_msg(begin, rnd)
    for i in range(depth):
        k = circ.num_vars_at(i + 1)
        ch = [rnd.draw() for _ in range(2 * k)]
        start, raw = prover.prove_layer(i, r_i, ch)
        msgs = prover.layer_messages(raw)
        replay = Replay(ch)
        verifier.receive_prover_msg(start, replay)
        for m in msgs[:-1]:
            verifier.receive_prover_msg(m, replay)
        verifier.final_random_point(replay)
        kind, r_i = verifier.receive_prover_msg(msgs[-1], rnd)
    assert verifier.check_input(inp)


def test_sharded_provers_match_single_gpu_when_the_box_has_more_gpus():
    """Runs scripts/mgpu_check.py (NCCL and NVLink-P2P sharded provers == single GPU, byte for byte) on every GPU of the
    box.  One GPU: the same script at world size 1 exercises the sharded entry points with a trivial exchange."""
    n = T.device_count()
    world = 1
    while world * 2 <= min(n, 8):
        world *= 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "mgpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU_OK" in out.stdout, out.stdout[-3000:]
