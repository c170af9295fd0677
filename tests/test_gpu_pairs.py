"""Two sum-check rounds per pass over the tables (csrc/pairs.cuh, small-prime fields): the (K+1)^2 grid sums against the
Python oracle's definition, the two-variable fold against two oracle folds, and the transcripts of the resident pair
kernel byte for byte against one launch per round and against the oracle."""
import os
import random
import subprocess
import sys

import pytest

from oracle import pyoracle as O

pytestmark = pytest.mark.gpu

T = pytest.importorskip("thaler_study_b200")

SP_FIELDS = [O.FP5, O.FP389, O.FP1572869, O.Field(268435399)]  # the reference's moduli + the largest prime below 2^28


def fid(f):
    return f"p{f.p}"


def oracle_grid(OF, tabs, K):
    p, NP = OF.p, K + 1
    n = len(tabs[0])
    H = [[0] * NP for _ in range(NP)]
    for g in range(n // 4):
        for a in range(NP):
            for b in range(NP):
                pr = 1
                for k in range(K):
                    c = tabs[k][4 * g:4 * g + 4]
                    v0 = (c[0] + a * (c[1] - c[0])) % p
                    v1 = (c[2] + a * (c[3] - c[2])) % p
                    pr = pr * ((v0 + b * (v1 - v0)) % p) % p
                H[a][b] = (H[a][b] + pr) % p
    return H


@pytest.mark.parametrize("OF", SP_FIELDS, ids=fid)
@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_grid_evals_and_pair_pass_vs_oracle(OF, K):
    F = T.Field(OF.p)
    rnd = random.Random(1000 * K + OF.p % 997)
    for v in (2, 3, 4, 5, 8, 11):
        tabs = [[rnd.randrange(OF.p) for _ in range(1 << v)] for _ in range(K)]
        g = T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, v, t) for t in tabs])
        assert g.grid_evals() == oracle_grid(OF, tabs, K), (v, K)
        if v >= 4:
            ra, rb = rnd.randrange(OF.p), rnd.randrange(OF.p)
            folded = [O.DenseMLE(OF, v, t).fix_variables([ra, rb]).evals for t in tabs]
            g2, H2 = g.pair_pass(ra, rb)
            assert g2.num_vars() == v - 2
            assert H2 == oracle_grid(OF, folded, K), (v, K)
            for k in range(K):
                assert g2.table(k).to_evaluations() == folded[k]
            if v >= 6:  # packed input this time
                rc, rd = rnd.randrange(OF.p), rnd.randrange(OF.p)
                folded2 = [O.DenseMLE(OF, v - 2, t).fix_variables([rc, rd]).evals for t in folded]
                g3, H3 = g2.pair_pass(rc, rd)
                assert H3 == oracle_grid(OF, folded2, K)
                assert g3.grid_evals() == H3
                assert [g3.table(k).to_evaluations() for k in range(K)] == folded2


@pytest.mark.parametrize("OF", SP_FIELDS[:3], ids=fid)
@pytest.mark.parametrize("kind", ["product1", "product2", "product3", "product4", "matmul"])
def test_pair_transcripts_vs_oracle(OF, kind):
    """generate_transcript routes small-prime product polynomials through the pair kernel: bytes == oracle, every v."""
    F = T.Field(OF.p)
    rnd = random.Random(5)
    for v in (2, 3, 4, 5, 6, 7, 10):
        if kind == "matmul":
            a = [rnd.randrange(OF.p) for _ in range(1 << v)]
            b = [rnd.randrange(OF.p) for _ in range(1 << v)]
            og = O.MatMulG(OF, O.DenseMLE(OF, v, a), O.DenseMLE(OF, v, b))
            dg = T.MatMulG.from_tables(T.DenseMultilinearExtension.from_evaluations_vec(F, v, a), T.DenseMultilinearExtension.from_evaluations_vec(F, v, b))
        else:
            K = int(kind[-1])
            vals = [[rnd.randrange(OF.p) for _ in range(1 << v)] for _ in range(K)]
            og = O.ProductMLE(OF, [O.DenseMLE(OF, v, t) for t in vals])
            dg = T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, v, t) for t in vals])
        want = O.generate_transcript(OF, O.Prover(og))
        got = T.generate_transcript(T.Prover(dg))
        assert got == want, (kind, v)
        assert T.verify_transcript(got, T.Verifier(v, dg))


def test_pair_kernel_matches_one_round_per_pass():
    """SCB_PAIRS=0 (one round per pass) and the default must print identical transcripts at sizes where the grid-wide
    barrier, the packed layouts and the solo endgame all take part."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import thaler_study_b200 as T; T.options_from_env()\n"
        "for p, v, K in ((1572869, 20, 3), (1572869, 19, 3), (1572869, 17, 4), (1572869, 16, 2), (389, 15, 1), (5, 14, 3), (268435399, 18, 3), (1572869, 5, 3), (1572869, 4, 2)):\n"
        "    F = T.Field(p)\n"
        "    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 70 + k) for k in range(K)])\n"
        "    print(b''.join(T.generate_transcript(T.Prover(g))).hex())\n"
        "    a = T.DenseMultilinearExtension.synthetic(F, v, 7); b = T.DenseMultilinearExtension.synthetic(F, v, 8)\n"
        "    print(b''.join(T.generate_transcript(T.Prover(T.MatMulG.from_tables(a, b)))).hex())\n"
        # accumulator head-room: the largest p below 2^28, > 16 thread-iterations per pass (so the periodic fold runs)
        # and the stored-word pattern [0, 0, p-1, p-1] that maximises every lazy grid value (5p-3 for K = 3)
        "import numpy as np, hashlib\n"
        "p = 268435399; F = T.Field(p)\n"
        "for v, K in ((24, 3), (23, 4), (23, 2), (22, 1)):\n"
        "    pat = np.tile(np.array([0, 0, p - 1, p - 1], dtype=np.uint64), (1 << v) // 4)\n"
        "    g = T.ProductMLE.new([T.DenseMultilinearExtension.from_evaluations_vec(F, v, pat) for _ in range(K)])\n"
        "    print(hashlib.sha256(b''.join(T.generate_transcript(T.Prover(g)))).hexdigest())\n"
        "    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 30 + k) for k in range(K)])\n"
        "    print(hashlib.sha256(b''.join(T.generate_transcript(T.Prover(g)))).hexdigest())\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    outs = []
    for env_add in ({"SCB_PAIRS": "0", "SCB_TAIL_VARS": "0"}, {"SCB_PAIRS": "0"}, {}, {"SCB_PAIR_BPS": "1"}, {"SCB_PAIR_RESIDENT": "0"},
                    {"SCB_PAIR_STAGE": "1"}, {"SCB_GRID_TMA": "1"}, {"SCB_GRID_PF": "0"}, {"SCB_PAIR_PIPE": "0"}):
        env = dict(os.environ, **env_add)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600))
    for o in outs:
        assert o.returncode == 0, o.stderr[-2000:]
    assert len(outs[0].stdout.split()) == 26
    assert len(set(o.stdout for o in outs)) == 1


def test_w21_triples_match_the_8_byte_first_pass():
    """Option pair_w21 (K = 3 tables over a field of at most 21 bits, first pair pass as its own launch): Prover::new's grid
    pass also writes word i = A[i] | B[i] << 21 | C[i] << 42 and the pair pass reads those words instead of the caller's
    8-byte tables (csrc/pairs.cuh).  Transcripts must not depend on it -- switched off, and on in its four variants (three /
    two CTAs per SM, nested folds / the bilinear two-variable fold; the default is the last) --
    and the stats must show that the triple kernels really ran exactly where they apply (not for K != 3, not for a
    28-bit field).  Includes tables of p - 1 everywhere (all 21 bits of every field of a word set)."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import ctypes as C, hashlib, numpy as np\n"
        "import thaler_study_b200 as T; T.options_from_env()\n"
        "def w21():\n"
        "    a, b = C.c_uint64(), C.c_uint64()\n"
        "    T.lib.scb_grid_pass_stats(None, None, C.byref(a), C.byref(b)); return a.value, b.value\n"
        "for p, v, K, applies in ((1572869, 20, 3, 1), (1572869, 17, 3, 1), (1572869, 13, 3, 1), (1572869, 12, 3, 1), (5, 14, 3, 1), (389, 15, 3, 1),\n"
        "                         (1572869, 11, 3, 0), (1572869, 16, 2, 0), (1572869, 16, 4, 0), (268435399, 16, 3, 0)):\n"
        "    F = T.Field(p)\n"
        "    T.lib.scb_resident_stats_reset()\n"
        "    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 90 + k) for k in range(K)])\n"
        "    tr = T.generate_transcript(T.Prover(g))\n"
        "    assert T.verify_transcript(tr, T.Verifier(v, g))\n"
        "    used = w21()\n"
        "    want = (1, 1) if (applies and T.get_option('pair_w21') != 0) else (0, 0)\n"
        "    assert used == want, (p, v, K, used, want)\n"
        "    print(hashlib.sha256(b''.join(tr)).hexdigest())\n"
        "p = 1572869; F = T.Field(p)\n"
        "for pat in ([p - 1], [0, p - 1, p - 1, 0, 1]):\n"
        "    tabs = [T.DenseMultilinearExtension.from_evaluations_vec(F, 18, np.resize(np.array(pat[k %% len(pat):] + pat[:k %% len(pat)], dtype=np.uint64), 1 << 18)) for k in range(3)]\n"
        "    g = T.ProductMLE.new(tabs)\n"
        "    tr = T.generate_transcript(T.Prover(g))\n"
        "    assert T.verify_transcript(tr, T.Verifier(18, g))\n"
        "    print(hashlib.sha256(b''.join(tr)).hexdigest())\n"
        # default threshold (2^26) again: with the triples the first pass is a launch of its own from 2^24 entries (pair_w21_alone)
        "T.set_option('pair_first_alone', 26); T.lib.scb_resident_stats_reset()\n"
        "g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, 24, 60 + k) for k in range(3)])\n"
        "tr = T.generate_transcript(T.Prover(g))\n"
        "assert T.verify_transcript(tr, T.Verifier(24, g))\n"
        "assert w21() == ((1, 1) if T.get_option('pair_w21') != 0 else (0, 0)), w21()\n"
        "print(hashlib.sha256(b''.join(tr)).hexdigest())\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    outs = []
    for env_add in ({"SCB_PAIR_W21": "0"}, {"SCB_PAIR_W21": "1"}, {"SCB_PAIR_W21": "2"}, {"SCB_PAIR_W21": "4"}, {}):
        env = dict(os.environ, SCB_PAIR_FIRST_ALONE="12", **env_add)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600))
    for o in outs:
        assert o.returncode == 0, o.stderr[-2000:]
    assert len(outs[0].stdout.split()) == 13
    assert len(set(o.stdout for o in outs)) == 1


# ----------------------------------------------------------------------------- the resident-kernel entry points, directly
def _mont1(F, x):
    return int(F.to_mont([x])[0, 0])


@pytest.mark.parametrize("OF,v,K,max_rounds", [(O.FP1572869, 16, 3, 0), (O.FP1572869, 17, 2, 5), (O.Field(0xFFFFFFFF00000001), 15, 3, 0),
                                               (O.FP389, 9, 4, 0), (O.BLS12_381_FR, 15, 2, 7)], ids=lambda x: str(getattr(x, "bits", x)))
def test_resident_rounds_callback_api(OF, v, K, max_rounds):
    """scb_poly_resident_rounds with a caller-supplied challenge callback (what a Rust generate_transcript with its own
    hashing would do): the sums handed to the callback and the folded polynomial it returns equal the ones
    scb_poly_fix_and_round_evals produces round by round with the same challenges."""
    import ctypes as C

    import numpy as np

    from thaler_study_b200 import _lib
    from thaler_study_b200._lib import check, lib

    F = T.Field(OF.p)
    rnd = random.Random(v * 131 + K)
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 400 + k) for k in range(K)])
    n_rounds = v - 1 if max_rounds == 0 else max_rounds
    ch = [rnd.randrange(OF.p) for _ in range(n_rounds + 1)]
    want, cur = [], g
    for j in range(n_rounds):
        cur, ev = cur.fix_and_round_evals(ch[j])
        want.append(ev)
    got = []
    N, npts = F.n, K + 1

    def cb(user, rnd_idx, evals, next_out):
        arr = np.ctypeslib.as_array(evals, shape=(npts * N,)).copy().reshape(npts, N)
        got.append(F.from_mont(arr))
        nxt = F.elem(ch[rnd_idx + 1])
        for i in range(N):
            next_out[i] = int(nxt.reshape(-1)[i])
        return 0

    done = C.c_uint32()
    folded = C.c_void_p()
    check(lib.scb_poly_resident_rounds(g._h, T.api._p64(F.elem(ch[0])), npts, max_rounds, _lib.ROUND_CB(cb), None, C.byref(done), C.byref(folded)))
    assert done.value == n_rounds and got == want
    out = type(g)(F, folded)
    assert out.num_vars() == v - n_rounds
    assert [out.table(k).to_evaluations() for k in range(K)] == [cur.table(k).to_evaluations() for k in range(K)]


@pytest.mark.parametrize("OF,v,K,max_passes", [(O.FP1572869, 16, 3, 0), (O.FP1572869, 15, 2, 3), (O.Field(268435399), 12, 4, 0), (O.FP5, 7, 1, 0)],
                         ids=lambda x: str(getattr(x, "bits", x)))
def test_resident_pairs_callback_api(OF, v, K, max_passes):
    """scb_poly_resident_pairs with a caller-supplied callback: every grid equals scb_poly_pair_pass's (and the
    oracle's definition through it), the line of an odd tail equals the ordinary round sums."""
    import ctypes as C

    import numpy as np

    from thaler_study_b200 import _lib
    from thaler_study_b200._lib import check, lib

    F = T.Field(OF.p)
    rnd = random.Random(v * 17 + K)
    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 500 + k) for k in range(K)])
    all_passes = (v - 1) // 2
    n_passes = all_passes if max_passes == 0 else max_passes
    pairs = [(rnd.randrange(OF.p), rnd.randrange(OF.p)) for _ in range(n_passes + 1)]
    want, cur = [], g
    for t in range(n_passes):
        if cur.num_vars() >= 4:
            cur, H = cur.pair_pass(*pairs[t])
            want.append([x for row in H for x in row])
        else:  # three variables: fold two, one left -> its line sums
            cur = cur.fix_variables(list(pairs[t]))
            want.append(cur.round_evals())
    got = []
    npts = K + 1

    def cb(user, pass_idx, n_vals, vals, next_pair):
        arr = np.ctypeslib.as_array(vals, shape=(n_vals,)).copy().reshape(n_vals, 1)
        got.append(F.from_mont(arr))
        a, b = pairs[pass_idx + 1]
        next_pair[0] = _mont1(F, a)
        next_pair[1] = _mont1(F, b)
        return 0

    done = C.c_uint32()
    folded = C.c_void_p()
    check(lib.scb_poly_resident_pairs(g._h, T.api._p64(F.elem(pairs[0][0])), T.api._p64(F.elem(pairs[0][1])), max_passes, _lib.PAIR_CB(cb), None,
                                      C.byref(done), C.byref(folded)))
    assert done.value == n_passes and got == want
    out = type(g)(F, folded)
    assert out.num_vars() == v - 2 * n_passes
    if out.num_vars() >= 1:
        assert [out.table(k).to_evaluations() for k in range(K)] == [cur.table(k).to_evaluations() for k in range(K)]
