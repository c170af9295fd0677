"""world_size-2/4 gloo test (CPU) of the sharded prover's host logic: slab layout, per-round all-gather of the
partial sums, modular combine, Fiat-Shamir chain, consolidation.  The per-rank compute engine is a TEST DOUBLE
backed by the Python oracle (the product engine is CUDA-only); the driver code under test is
thaler_study_b200/distributed.py and the C++ transcript state machine."""
import os
import random
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle as O  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_oracle_engine(T, OF, F, kind, tables):
    from thaler_study_b200.distributed import LocalEngine

    class OracleEngine(LocalEngine):
        def __init__(self, tables):
            self.F, self.kind = F, kind
            self.tabs = [O.DenseMLE(OF, len(t).bit_length() - 1, t) for t in tables]
            self.n_points = len(tables) + 1

        def _poly(self):
            return O.ProductMLE(OF, self.tabs)

        def num_vars(self):
            return self.tabs[0].num_vars

        def new_buffer(self, shape):
            return torch.empty(list(shape), dtype=torch.int64)

        def _write(self, out, evals):
            m = np.array([[((OF.to_mont(e)) >> (64 * l)) & 0xFFFFFFFFFFFFFFFF for l in range(F.n)] for e in evals], dtype=np.uint64)
            out.copy_(torch.from_numpy(m.view(np.int64)))

        def round_evals(self, out):
            self._write(out, self._poly().round_evals())

        def fix_and_round_evals(self, r_mont, out):
            r = OF.from_mont(sum(int(x) << (64 * l) for l, x in enumerate(r_mont.reshape(-1).tolist())))
            self.tabs = [t.fix_variables([r]) for t in self.tabs]
            self._write(out, self._poly().round_evals())

        def slabs(self):
            res = []
            for t in self.tabs:
                m = np.array([[(OF.to_mont(e) >> (64 * l)) & 0xFFFFFFFFFFFFFFFF for l in range(F.n)] for e in t.evals], dtype=np.uint64)
                res.append(torch.from_numpy(m.view(np.int64)).clone())
            return res

        def from_slabs(self, tables):
            vals = []
            for t in tables:
                a = t.numpy().view(np.uint64)
                vals.append([OF.from_mont(sum(int(x) << (64 * l) for l, x in enumerate(row))) for row in a.tolist()])
            return OracleEngine(vals)

    return OracleEngine(tables)


def _worker(rank, world, port, p, v, K, consolidate_at, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import thaler_study_b200 as T
        from thaler_study_b200.distributed import prove_sharded

        OF, F = O.Field(p), T.Field(p)
        rnd = random.Random(1234)
        full = [[rnd.randrange(p) for _ in range(1 << v)] for _ in range(K)]
        slab = (1 << v) // world
        local = [t[rank * slab : (rank + 1) * slab] for t in full]
        eng = make_oracle_engine(T, OF, F, T.KIND_PRODUCT, local)
        c_1, msgs = prove_sharded(eng, consolidate_at=consolidate_at)
        want = O.generate_transcript(OF, O.Prover(O.ProductMLE(OF, [O.DenseMLE(OF, v, t) for t in full])))
        want_c1 = sum(O.ProductMLE(OF, [O.DenseMLE(OF, v, t) for t in full]).to_evaluations()) % p
        q.put((rank, msgs == want and c_1 == want_c1))
    except Exception as ex:  # report instead of leaving the parent to time out
        q.put((rank, f"{type(ex).__name__}: {ex}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,p,v,K,cat", [(2, 1572869, 6, 3, 2), (2, 5, 5, 2, 1), (4, 389, 6, 2, 2), (2, O.BLS12_381_FR.p, 4, 3, 8)])
def test_sharded_transcript_equals_single_prover(world, p, v, K, cat):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, p, v, K, cat, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
