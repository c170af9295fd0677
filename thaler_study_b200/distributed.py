"""Sharded multi-GPU sum-check prover: one process per GPU, torch.distributed for the plumbing.

Every table is split by its TOP log2(G) index bits: rank g owns the contiguous slab [g*2^lv, (g+1)*2^lv).  ark folds
variable 0 = index LSB first, so adjacent pairs never straddle ranks and no table data moves while a slab has more
than ``consolidate_at`` variables.  Per round each rank reduces its slab to (d+1) partial field sums on the device;
the only exchange is one all-gather of those (d+1)*E bytes per rank; every rank then adds the G rows mod p, derives
the message polynomial and the Fiat-Shamir challenge (identical on all ranks, so nothing is broadcast) and launches
the next fused fold+message kernel.  When the slabs are down to 2^consolidate_at entries they are all-gathered once
and the remaining rounds run replicated on every rank with no further communication.

The transcript is bit-identical to the single-GPU proof of the concatenated tables (tests/test_distributed_gloo.py
checks this over gloo with an oracle-backed engine; tests/test_gpu_parity.py on NCCL when >1 GPU is present).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import api
from ._lib import check, lib


class LocalEngine:
    """What the sharded driver needs from the per-rank compute engine.  All tensors are torch.int64 views of
    Montgomery limbs with shape [..., n_limbs] on the engine's device."""

    F: api.Field
    kind: int
    n_points: int

    def num_vars(self) -> int:  # local variables left in this rank's slab
        raise NotImplementedError

    def round_evals(self, out) -> None:  # out: int64[n_points, n_limbs]
        raise NotImplementedError

    def fix_and_round_evals(self, r_mont: np.ndarray, out) -> None:
        raise NotImplementedError

    def slabs(self) -> list:  # current tables as int64[2^num_vars, n_limbs] tensors
        raise NotImplementedError

    def from_slabs(self, tables: list) -> "LocalEngine":  # engine over (gathered) tables
        raise NotImplementedError

    def new_buffer(self, shape: Sequence[int]):
        raise NotImplementedError


class CudaProductEngine(LocalEngine):
    """ProductMLE<K> / matrix_multiplication::G slab on this rank's GPU (libsumcheck_b200.so kernels)."""

    def __init__(self, poly: api.SumCheckPolynomial, keepalive=None):
        import torch

        self._torch = torch
        self.poly = poly
        self.F = poly.F
        self.kind = poly.kind
        self.n_points = poly.n_points
        self._keep = keepalive
        check(lib.scb_poly_allow_packed(poly._h, 1))  # this engine owns its private clone of the tables
        o = C.c_uint32()
        check(lib.scb_poly_n_tables(poly._h, C.byref(o)))
        self.n_tables = o.value

    def num_vars(self) -> int:
        return self.poly.num_vars()

    def new_buffer(self, shape):
        return self._torch.empty(list(shape), dtype=self._torch.int64, device="cuda")

    def round_evals(self, out) -> None:
        check(lib.scb_poly_round_evals_device(self.poly._h, self.n_points, C.c_void_p(out.data_ptr())))

    def fix_and_round_evals(self, r_mont: np.ndarray, out) -> None:
        h = C.c_void_p()
        check(lib.scb_poly_fix_and_round_evals_device(self.poly._h, api._p64(r_mont), self.n_points, C.byref(h), C.c_void_p(out.data_ptr())))
        self.poly = type(self.poly)(self.F, h)

    def slabs(self) -> list:
        out = []
        for k in range(self.n_tables):
            m = self.poly.table(k)
            t = self.new_buffer([1 << m.num_vars, self.F.n])
            check(lib.scb_mle_copy_to_device(m._h, C.c_void_p(t.data_ptr())))
            out.append(t)
        return out

    def from_slabs(self, tables: list) -> "CudaProductEngine":
        nv = int(tables[0].shape[0]).bit_length() - 1
        mles = [api.DenseMultilinearExtension.from_device(self.F, nv, t.data_ptr(), copy=False) for t in tables]
        if self.kind == api.KIND_MATMUL_G:
            poly = api.MatMulG.from_tables(mles[0], mles[1])
        else:
            poly = api.ProductMLE.new(mles)
        return CudaProductEngine(poly, keepalive=tables)  # the borrowed tensors must outlive the handles


def _all_gather(dist, group, out, inp):
    if hasattr(dist, "all_gather_into_tensor"):
        dist.all_gather_into_tensor(out.view(-1), inp.view(-1), group=group)  # rank-order concatenation
    else:  # pragma: no cover
        chunks = list(out.chunk(dist.get_world_size(group)))
        dist.all_gather(chunks, inp, group=group)


def prove_sharded(engine: LocalEngine, group=None, consolidate_at: int = 16) -> Tuple[int, List[bytes]]:
    """Fiat-Shamir sum-check proof of the polynomial whose tables are the rank-order concatenation of every rank's
    slabs.  Returns (c_1, [g_1 bytes, g_2 bytes, ...]) -- fiat_shamir::generate_transcript's output
    (fiat-shamir/src/lib.rs:75-98) -- identically on every rank."""
    import torch.distributed as dist

    G = dist.get_world_size(group) if dist.is_initialized() else 1
    assert G & (G - 1) == 0, "the number of ranks must be a power of two (tables shard by their top variables)"
    lg = G.bit_length() - 1
    consolidate_at = max(1, consolidate_at)
    F, npts = engine.F, engine.n_points
    v = engine.num_vars() + lg
    transcript = api.Transcript(F, engine.kind)

    def gather_tables(eng: LocalEngine) -> LocalEngine:
        gathered = []
        for t in eng.slabs():
            full = eng.new_buffer([G * t.shape[0], t.shape[1]])
            _all_gather(dist, group, full, t.contiguous())
            gathered.append(full)
        return eng.from_slabs(gathered)

    sharded = G > 1
    if sharded and engine.num_vars() <= consolidate_at:
        engine, sharded = gather_tables(engine), False
    mine = engine.new_buffer([npts, F.n])
    parts = engine.new_buffer([G, npts, F.n]) if sharded else None

    def exchange_and_absorb() -> np.ndarray:
        if sharded:
            _all_gather(dist, group, parts, mine)
            host = parts.cpu().numpy().view(np.uint64)
        else:
            host = mine.cpu().numpy().view(np.uint64)[None]
        return transcript.absorb_round_mont(host).copy()

    engine.round_evals(mine)
    r = exchange_and_absorb()
    for _ in range(1, v):
        if sharded and engine.num_vars() <= consolidate_at:
            engine, sharded = gather_tables(engine), False
        engine.fix_and_round_evals(r, mine)
        r = exchange_and_absorb()
    return transcript.c_1(), transcript.messages()


# ---------------------------------------------------------------------------------------------------------------------
# NVLink peer-memory exchange (the default on GPUs): the per-round all-gather + modular sum happens INSIDE the round
# kernel -- its finishing thread posts the rank's partial sums into every peer's window with P2P stores, waits for the
# peers' posts and adds the rows -- and the whole round loop (hash chain, consolidation, resident tail kernel) runs in
# the C++ host layer.  torch.distributed is only used to hand the 64-byte IPC handles around at set-up.
# ---------------------------------------------------------------------------------------------------------------------
class Peers:
    """scb_peers: this rank's exchange window + IPC mappings of every peer's window."""

    def __init__(self, group=None, gather_bytes: int = 32 << 20):
        import torch
        import torch.distributed as dist

        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # the ranks of one box share its host cores (pack threads of the narrowing upload); the library does not read
        # the environment, so the driver tells it
        from ._lib import set_option
        import os

        set_option("local_ranks", max(1, int(os.environ.get("LOCAL_WORLD_SIZE", self.world))))
        self._h = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        # every rank takes part in every collective below even if a local step fails; the outcome is agreed on last
        err = None
        rc = lib.scb_peers_create(self.rank, self.world, gather_bytes, C.byref(self._h), handle)
        if rc != 0:
            err = (lib.scb_last_error() or b"").decode()
        if self.world > 1:
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device="cuda")
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(allh, mine, group=group)
            if err is None:
                buf = (C.c_uint8 * (self.world * 64)).from_buffer_copy(bytes(allh.cpu().tolist()))
                if lib.scb_peers_connect(self._h, buf) != 0:
                    err = (lib.scb_last_error() or b"").decode()
            flag = torch.tensor([0 if err is None else 1], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            if flag.item() != 0:
                self.close()
                raise RuntimeError(f"peer windows unavailable on at least one rank ({err or 'error on another rank'})")
        elif err is not None:
            raise RuntimeError(err)

    def close(self):
        if getattr(self, "_h", None) and lib is not None:  # at interpreter shutdown the module globals may be gone already
            lib.scb_peers_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def prove_sharded_p2p(slab_poly: api.SumCheckPolynomial, peers: Peers, consolidate_at: int = 16) -> Tuple[int, List[bytes]]:
    """Same contract as prove_sharded, with the exchange fused into the round kernels over NVLink peer memory."""
    prover = api.Prover.__new__(api.Prover)
    prover.F = slab_poly.F
    prover._h = C.c_void_p()
    check(lib.scb_prover_new_sharded(slab_poly._h, peers._h, peers.world, consolidate_at, C.byref(prover._h)))
    c_1 = prover.c_1()
    return c_1, api.generate_transcript(prover)


# ---------------------------------------------------------------------------------------------------------------------
# Sharded MLE evaluation and the verifier's final oracle query for a sharded polynomial ("MLE-eval shards
# trivially: each GPU dots its slab with its slice of the eq table; one exchange of E bytes").
# ---------------------------------------------------------------------------------------------------------------------
def mle_evaluate_sharded(slab: api.DenseMultilinearExtension, peers: Peers, point: Sequence[int], big_endian: bool = False) -> int:
    """Evaluation of the table whose rank-order concatenation of slabs is the full table, at a point with
    num_vars(slab) + log2(world) coordinates (LSB-first like ark's evaluate, or the multilinear-extensions order with
    ``big_endian``).  One launch per rank (scb_mle_evaluate_sharded); same value on every rank."""
    F = slab.F
    pt = F.to_mont(list(point))
    out = np.zeros((1, F.n), dtype=np.uint64)
    check(lib.scb_mle_evaluate_sharded(slab._h, peers._h, api._p64(pt), len(point), 1 if big_endian else 0, api._p64(out)))
    return F.from_mont(out)[0]


def poly_evaluate_sharded(slab_poly: api.SumCheckPolynomial, peers: Peers, point: Sequence[int]) -> int:
    """SumCheckPolynomial::evaluate of a sharded product polynomial: the product of its tables' evaluations
    (matrix-multiplication/src/lib.rs:96-101)."""
    assert slab_poly.kind in (api.KIND_PRODUCT, api.KIND_MATMUL_G), "only product polynomials shard"
    o = C.c_uint32()
    check(lib.scb_poly_n_tables(slab_poly._h, C.byref(o)))
    val = 1
    for k in range(o.value):
        val = val * mle_evaluate_sharded(slab_poly.table(k), peers, point) % slab_poly.F.p
    return val


def parse_message(F: api.Field, msg: bytes, first: bool):
    """One transcript message -> (c_1 or None, SparsePolynomial): [ARK] serialize_uncompressed of (c_1, g_1) or g_j
    (fiat-shamir/src/lib.rs:48-50,58): canonical little-endian field elements, u64 lengths and degrees."""
    sb, off, c_1 = F.ser_bytes, 0, None
    if first:
        c_1 = int.from_bytes(msg[:sb], "little")
        off = sb
    n = int.from_bytes(msg[off:off + 8], "little")
    off += 8
    terms = []
    for _ in range(n):
        d = int.from_bytes(msg[off:off + 8], "little")
        cf = int.from_bytes(msg[off + 8:off + 8 + sb], "little")
        if cf >= F.p:
            raise ValueError("Codec error")
        terms.append((d, cf))
        off += 8 + sb
    if off != len(msg) or (c_1 is not None and c_1 >= F.p):
        raise ValueError("Codec error")
    return c_1, api.SparsePolynomial(F, terms)


def verify_transcript_sharded(transcript: Sequence[bytes], slab_poly: api.SumCheckPolynomial, peers: Peers) -> bool:
    """fiat_shamir::verify_transcript (fiat-shamir/src/lib.rs:123-143) over Verifier::round
    (sum-check-protocol/src/lib.rs:278-330, with the strict final-round link check) for a polynomial that only exists
    as slabs: the hash chain and the round checks are host arithmetic on every rank; the final oracle query g(r) is the
    sharded evaluation above.  Same verdict on every rank."""
    F = slab_poly.F
    lg = peers.world.bit_length() - 1
    n = slab_poly.num_vars() + lg
    if len(transcript) != n:
        return False
    claim, rs, sofar = None, [], b""
    for j, msg in enumerate(transcript):
        c_1, g_j = parse_message(F, msg, j == 0)
        sofar += msg
        r_j = F.hash_to_field(sofar)
        if j == 0:
            claim = c_1
        if (g_j.evaluate(0) + g_j.evaluate(1)) % F.p != claim:
            return False  # Error::ProverClaimMismatch
        claim = g_j.evaluate(r_j)
        rs.append(r_j)
    return claim == poly_evaluate_sharded(slab_poly, peers, rs)
