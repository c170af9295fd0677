"""GPU tests of the packed uint32 intermediate tables (small-prime policy, packed.cuh): same field elements and
transcript bytes as ark's 8-byte layout, and every public entry point still works on a packed handle."""
import ctypes as C
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle as O
from oracle.coracle import CField

import thaler_study_b200 as T
from thaler_study_b200._lib import check, lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("p", [5, 389, 1572869])
@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_packed_children_equal_plain_children(p, K):
    if K >= p:
        pytest.skip("degree >= characteristic")
    OF, F, cf = O.Field(p), T.Field(p), CField(p)
    rnd = random.Random(p * 10 + K)
    for v in (2, 3, 4, 7, 12):
        tabs_c = [cf.synth(300 + k, 0, 1 << v) for k in range(K)]
        mk = lambda: T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 300 + k) for k in range(K)])
        plain, packed = mk(), mk()
        check(lib.scb_poly_allow_packed(packed._h, 1))
        cur_c = tabs_c
        for j in range(1, v):
            r = rnd.randrange(p)
            plain, ev_a = plain.fix_and_round_evals(r)
            packed, ev_b = packed.fix_and_round_evals(r)  # 64->32 on the first round, 32->32 afterwards
            cur_c = [cf.fix_variable(t, cf.to_mont([r])) for t in cur_c]
            assert ev_a == ev_b == cf.from_mont(cf.product_round_evals(cur_c, K + 1)), (v, j)
            # the packed handle behaves like a plain one through every other entry point
            for k in range(K):
                assert np.array_equal(packed.table(k).to_evaluations_mont(), cur_c[k])
            assert packed.num_vars() == v - j
            assert packed.sum() == plain.sum()
            assert packed.round_evals() == ev_a
            pt = [rnd.randrange(p) for _ in range(v - j)]
            assert packed.evaluate(pt) == plain.evaluate(pt)
            if v - j >= 1:
                assert packed.fix_variables(pt[:1]).to_evaluations() == plain.fix_variables(pt[:1]).to_evaluations()
            assert packed.to_evaluations() == plain.to_evaluations()


def test_transcript_independent_of_packing_and_tail():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import thaler_study_b200 as T; T.options_from_env()\n"
        "for p, v, K in ((1572869, 19, 3), (389, 16, 2), (5, 15, 4)):\n"
        "    F = T.Field(p)\n"
        "    g = T.ProductMLE.new([T.DenseMultilinearExtension.synthetic(F, v, 70 + k) for k in range(K)])\n"
        "    print(b''.join(T.generate_transcript(T.Prover(g))).hex())\n"
    ) % ROOT
    outs = []
    for packed, tail in (("0", "0"), ("1", "0"), ("0", "14"), ("1", "14"), ("1", "6")):
        env = dict(os.environ, SCB_PACKED=packed, SCB_TAIL_VARS=tail)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600))
    for o in outs:
        assert o.returncode == 0, o.stderr[-2000:]
    assert len(set(o.stdout for o in outs)) == 1 and len(outs[0].stdout.split()) == 3
    # anchor: the 2^16 / F_389 proof against the C oracle's round sums through the host transcript object
    OF, cf, F = O.FP389, CField(389), T.Field(389)
    v, K = 16, 2
    tabs = [cf.synth(70 + k, 0, 1 << v) for k in range(K)]
    tr = T.Transcript(F, T.KIND_PRODUCT)
    cur = tabs
    r = tr.absorb_round_mont(cf.product_round_evals(cur, K + 1)[None])
    for j in range(1, v):
        cur = [cf.fix_variable(t, r.copy()) for t in cur]
        r = tr.absorb_round_mont(cf.product_round_evals(cur, K + 1)[None])
    assert b"".join(tr.messages()).hex() == outs[0].stdout.split()[1]
