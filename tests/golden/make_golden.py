#!/usr/bin/env python
"""Generates tests/golden/*.json with the Python oracle (oracle/pyoracle.py).

The reference is Rust on un-vendored arkworks and cannot run in this image, so these vectors are NOT outputs of the
reference binary: they are (a) the known-answer values the reference's own tests assert, copied as data with their
file:line (reference_kats.json), and (b) transcripts / round sums produced by the oracle that is pinned to (a)
(oracle_transcripts.json).  Re-run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as O  # noqa: E402

kats = {
    "mle_example_from_book": {
        "cite": "multilinear-extensions/src/lib.rs:76-120", "modulus": 5, "evals": [1, 2, 1, 4],
        "table": [[1, 2, 3, 4, 0], [1, 4, 2, 0, 3], [1, 1, 1, 1, 1], [1, 3, 0, 2, 4], [1, 0, 4, 3, 2]],
    },
    "matrix_test_from_book": {
        "cite": "matrix-multiplication/src/lib.rs:202-243", "modulus": 5, "a": [[0, 1], [2, 0]], "b": [[1, 0], [0, 4]], "c": [[0, 4], [2, 0]],
    },
    "triangle_simple_matrix": {
        "cite": "triangle-counting/src/lib.rs:224-266", "modulus": 389,
        "adj": [[0, 1, 1, 0], [1, 0, 1, 0], [1, 1, 0, 0], [0, 0, 0, 0]], "c_1": 6,
    },
    "gkr_restrict_poly": {
        "cite": "gkr-protocol/src/lib.rs:506-548", "modulus": 389, "b": [2, 4], "c": [3, 2], "evals": [0, 0, 2, 5], "coeffs": [32, 385, 383],
    },
    "gkr_circuit_from_book": {
        "cite": "gkr-protocol/src/circuit.rs:258-284", "input": [3, 2, 3, 1], "layers": [[36, 6], [9, 4, 6, 1], [3, 2, 3, 1]],
    },
}
json.dump(kats, open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1)

rnd = random.Random(0x601D)
cases = []
FIELDS = [5, 389, 1572869, (1 << 61) - 1, 0xFFFFFFFF00000001, O.BLS12_381_FR.p]
for p in FIELDS:
    F = O.Field(p)
    for kind, K in (("product", 1), ("product", 2), ("product", 3), ("product", 4), ("matmul_g", 2)):
        if K >= p:
            continue
        for v in (1, 3, 6):
            tables = [[rnd.randrange(p) for _ in range(1 << v)] for _ in range(K)]
            mles = [O.DenseMLE(F, v, t) for t in tables]
            g = O.MatMulG(F, mles[0], mles[1]) if kind == "matmul_g" else O.ProductMLE(F, mles)
            prover = O.Prover(g)
            transcript = O.generate_transcript(F, O.Prover(g))
            point = [rnd.randrange(p) for _ in range(v)]
            cases.append({
                "modulus": str(p), "kind": kind, "num_vars": v, "tables": [[str(x) for x in t] for t in tables],
                "c_1": str(prover.c_1()), "round0_evals": [str(x) for x in g.round_evals()],
                "transcript_hex": [m.hex() for m in transcript], "point": [str(x) for x in point], "evaluate": str(g.evaluate(point)),
                "mle_eval_be_table0": str(O.vsbw_multilinear_from_evaluations(F, tables[0], point)),
            })
# triangle / GKR-W (fields with a size-4 FFT domain, like the reference's)
for p in (389, 1572869):
    F = O.Field(p)
    for n in (1, 2):
        size = 1 << n
        m = [[0] * size for _ in range(size)]
        for i in range(size):
            for j in range(i + 1, size):
                m[i][j] = m[j][i] = rnd.randrange(2)
        flat = sum(m, [])
        g = O.TriangleG.new_adj_matrix(F, 2 * n, [bool(x) for x in flat])
        cases.append({"modulus": str(p), "kind": "triangle_g", "num_vars": 3 * n, "adj": flat, "c_1": str(O.Prover(g).c_1()),
                      "transcript_hex": [x.hex() for x in O.generate_transcript(F, O.Prover(g))]})
    k = 2
    tabs = [[rnd.randrange(p) for _ in range(1 << nv)] for nv in (2 * k, 2 * k, k, k)]
    w = O.GkrW(F, *[O.DenseMLE(F, nv, t) for nv, t in zip((2 * k, 2 * k, k, k), tabs)])
    cases.append({"modulus": str(p), "kind": "gkr_w", "num_vars": 2 * k, "tables": [[str(x) for x in t] for t in tabs],
                  "c_1": str(O.Prover(w).c_1()), "transcript_hex": [x.hex() for x in O.generate_transcript(F, O.Prover(w))]})
# hash_to_field vectors
h2f = []
for p in FIELDS:
    for msg in (b"", b"abc", bytes(range(100))):
        h2f.append({"modulus": str(p), "msg_hex": msg.hex(), "out": str(O.hash_to_field(O.Field(p), msg))})
json.dump({"generator": "tests/golden/make_golden.py (oracle/pyoracle.py); parity unpinned by the reference, see DESIGN.md section 5",
           "cases": cases, "hash_to_field": h2f}, open(os.path.join(HERE, "oracle_transcripts.json"), "w"), indent=0)
print(len(cases), "transcript cases,", len(h2f), "hash vectors")
