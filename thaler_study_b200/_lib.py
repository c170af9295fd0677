"""ctypes loader for libsumcheck_b200.so (the C ABI declared in include/sumcheck_b200.h).

The library is built in-tree by ``thaler_study_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is no fallback: if the shared object is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libsumcheck_b200.so")

SCB_OK, SCB_EINVAL, SCB_ENOMEM, SCB_ECUDA, SCB_ENCCL, SCB_EVERIFY, SCB_ENOPOLY = 0, -1, -2, -3, -4, -5, -6

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)
# scb_round_cb: int (*)(void* user, uint32_t round, const uint64_t* evals, uint64_t* next_challenge_out)
PAIR_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
ROUND_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, u64p, u64p)

# name -> (restype, argtypes); every symbol include/sumcheck_b200.h declares
SIGNATURES = {
    "scb_last_error": (C.c_char_p, []),
    "scb_version": (C.c_char_p, []),
    "scb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "scb_set_stream": (C.c_int, [vp]),
    "scb_synchronize": (C.c_int, []),
    "scb_launch_count": (C.c_int, [u64p, C.c_int]),
    "scb_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
    "scb_get_option": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64)]),
    "scb_option_name": (C.c_char_p, [C.c_uint32]),
    "scb_reset_options": (None, []),
    "scb_field_create": (C.c_int, [C.c_uint32, u64p, vpp]),
    "scb_field_free": (None, [vp]),
    "scb_field_n_limbs": (C.c_int, [vp, u32p]),
    "scb_field_modulus_bits": (C.c_int, [vp, u32p]),
    "scb_field_policy": (C.c_int, [vp, u32p]),
    "scb_field_to_mont": (C.c_int, [vp, u64p, u64p, C.c_size_t]),
    "scb_field_from_mont": (C.c_int, [vp, u64p, u64p, C.c_size_t]),
    "scb_mle_from_host": (C.c_int, [vp, C.c_uint32, u64p, vpp]),
    "scb_mle_from_device": (C.c_int, [vp, C.c_uint32, vp, C.c_int, vpp]),
    "scb_mle_synthetic": (C.c_int, [vp, C.c_uint32, C.c_uint64, C.c_uint64, vpp]),
    "scb_mle_clone": (C.c_int, [vp, vpp]),
    "scb_mle_free": (None, [vp]),
    "scb_mle_num_vars": (C.c_int, [vp, u32p]),
    "scb_mle_device_ptr": (C.c_int, [vp, vpp]),
    "scb_mle_fix_variables": (C.c_int, [vp, u64p, C.c_uint32, vpp]),
    "scb_mle_evaluate": (C.c_int, [vp, u64p, C.c_uint32, u64p]),
    "scb_mle_evaluate_be": (C.c_int, [vp, u64p, C.c_uint32, u64p]),
    "scb_mle_evaluate_many": (C.c_int, [vp, u64p, C.c_uint32, C.c_uint32, u64p]),
    "scb_mle_relabel": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, vpp]),
    "scb_mle_to_evaluations": (C.c_int, [vp, u64p, C.c_size_t]),
    "scb_mle_copy_to_device": (C.c_int, [vp, vp]),
    "scb_vsbw_multilinear_from_evaluations": (C.c_int, [vp, u64p, C.c_size_t, u64p, C.c_uint32, u64p]),
    "scb_cti_multilinear_from_evaluations": (C.c_int, [vp, u64p, C.c_size_t, u64p, C.c_uint32, u64p]),
    "scb_poly_product": (C.c_int, [vpp, C.c_uint32, vpp]),
    "scb_poly_matmul_g": (C.c_int, [vp, vp, vpp]),
    "scb_poly_product_from_host": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.POINTER(u64p), vpp]),
    "scb_host_pack_stats": (C.c_int, [u64p, u64p, u64p]),
    "scb_host_pack_selftest": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]),
    "scb_poly_matmul_g_new": (C.c_int, [vp, C.c_uint32, u64p, u64p, u64p, vpp]),
    "scb_poly_triangle_g_new": (C.c_int, [vp, C.c_uint32, u8p, vpp]),
    "scb_poly_gkr_w": (C.c_int, [vp, vp, vp, vp, vpp]),
    "scb_poly_clone": (C.c_int, [vp, vpp]),
    "scb_poly_free": (None, [vp]),
    "scb_poly_kind_of": (C.c_int, [vp, u32p]),
    "scb_poly_n_tables": (C.c_int, [vp, u32p]),
    "scb_poly_table": (C.c_int, [vp, C.c_uint32, vpp]),
    "scb_poly_n_points": (C.c_int, [vp, u32p]),
    "scb_poly_evaluate": (C.c_int, [vp, u64p, C.c_uint32, u64p]),
    "scb_poly_fix_variables": (C.c_int, [vp, u64p, C.c_uint32, vpp]),
    "scb_poly_num_vars": (C.c_int, [vp, u32p]),
    "scb_poly_to_evaluations": (C.c_int, [vp, u64p, C.c_size_t]),
    "scb_poly_to_univariate": (C.c_int, [vp, u64p, u64p, C.c_uint32, u32p]),
    "scb_poly_round_evals": (C.c_int, [vp, C.c_uint32, u64p]),
    "scb_poly_sum": (C.c_int, [vp, u64p]),
    "scb_poly_fix_and_round_evals": (C.c_int, [vp, u64p, C.c_uint32, vpp, u64p]),
    "scb_poly_fix_and_round_evals_claim": (C.c_int, [vp, u64p, u64p, C.c_uint32, vpp, u64p]),
    "scb_poly_round_evals_device": (C.c_int, [vp, C.c_uint32, vp]),
    "scb_poly_fix_and_round_evals_device": (C.c_int, [vp, u64p, C.c_uint32, vpp, vp]),
    "scb_poly_allow_packed": (C.c_int, [vp, C.c_int]),
    "scb_poly_tail_rounds": (C.c_int, [vp, u64p, C.c_uint32, ROUND_CB, vp, u32p]),
    "scb_poly_grid_evals": (C.c_int, [vp, u64p]),
    "scb_poly_pair_pass": (C.c_int, [vp, u64p, u64p, C.POINTER(vp), u64p]),
    "scb_poly_resident_pairs": (C.c_int, [vp, u64p, u64p, C.c_uint32, PAIR_CB, vp, u32p, C.POINTER(vp)]),
    "scb_resident_stats": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_double), u32p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint32]),
    "scb_pair_pass_stats": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "scb_grid_pass_stats": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "scb_resident_stats_reset": (None, []),
    "scb_poly_resident_rounds": (C.c_int, [vp, u64p, C.c_uint32, C.c_uint32, ROUND_CB, vp, u32p, C.POINTER(vp)]),
    "scb_evals_to_univariate": (C.c_int, [vp, C.c_uint32, u64p, C.c_uint32, u64p, u64p, C.c_uint32, u32p]),
    "scb_unipoly_serialize": (C.c_int, [vp, u64p, u64p, C.c_uint32, u8p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "scb_unipoly_evaluate": (C.c_int, [vp, u64p, u64p, C.c_uint32, u64p, u64p]),
    "scb_hash_to_field": (C.c_int, [vp, u8p, C.c_size_t, u64p]),
    "scb_prover_new": (C.c_int, [vp, vpp]),
    "scb_prover_free": (None, [vp]),
    "scb_prover_c_1": (C.c_int, [vp, u64p]),
    "scb_prover_num_vars": (C.c_int, [vp, u32p]),
    "scb_prover_round": (C.c_int, [vp, u64p, C.c_uint32, u64p, u64p, C.c_uint32, u32p]),
    "scb_verifier_new": (C.c_int, [vp, C.c_uint32, vp, vpp]),
    "scb_verifier_free": (None, [vp]),
    "scb_verifier_set_c_1": (C.c_int, [vp, u64p]),
    "scb_verifier_round": (C.c_int, [vp, u64p, u64p, C.c_uint32, u64p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "scb_fs_generate_transcript": (C.c_int, [vp, u8p, C.c_size_t, C.POINTER(C.c_size_t), u64p]),
    "scb_fs_verify_transcript": (C.c_int, [vp, u8p, C.c_size_t, u64p, C.c_uint32, C.POINTER(C.c_int)]),
    "scb_circuit_create": (C.c_int, [vp, C.c_uint32, u32p, u8p, u32p, u32p, C.c_uint32, vpp]),
    "scb_circuit_free": (None, [vp]),
    "scb_circuit_num_vars_at": (C.c_int, [vp, C.c_uint32, u32p]),
    "scb_circuit_wiring_eval": (C.c_int, [vp, C.c_uint32, u64p, u64p, u64p, u64p, u64p]),
    "scb_gkr_prover_new": (C.c_int, [vp, u64p, C.c_size_t, vpp]),
    "scb_gkr_prover_free": (None, [vp]),
    "scb_gkr_prover_layer": (C.c_int, [vp, C.c_uint32, vpp]),
    "scb_gkr_prover_start_round": (C.c_int, [vp, C.c_uint32, u64p, u64p, u32p]),
    "scb_gkr_prover_round_evals": (C.c_int, [vp, C.c_uint32, u64p, u64p]),
    "scb_gkr_prover_restrict_evals": (C.c_int, [vp, u64p, u64p, C.c_uint32, u32p]),
    "scb_gkr_prover_prove_layer": (C.c_int, [vp, C.c_uint32, u64p, u64p, C.c_uint32, u64p, u64p, u64p, C.c_uint32, u32p]),
    "scb_peers_create": (C.c_int, [C.c_uint32, C.c_uint32, C.c_size_t, vpp, u8p]),
    "scb_peers_gather_capacity": (C.c_int, [vp, C.POINTER(C.c_size_t)]),
    "scb_peers_connect": (C.c_int, [vp, u8p]),
    "scb_peers_free": (None, [vp]),
    "scb_peers_set_current": (C.c_int, [vp]),
    "scb_peers_gather_poly": (C.c_int, [vp, vp, vpp]),
    "scb_prover_new_sharded": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, vpp]),
    "scb_mle_evaluate_sharded": (C.c_int, [vp, vp, u64p, C.c_uint32, C.c_int, u64p]),
    "scb_transcript_new": (C.c_int, [vp, C.c_uint32, vpp]),
    "scb_transcript_free": (None, [vp]),
    "scb_transcript_absorb_round": (C.c_int, [vp, u64p, C.c_uint32, C.c_uint32, u64p]),
    "scb_transcript_c_1": (C.c_int, [vp, u64p]),
    "scb_transcript_bytes": (C.c_int, [vp, u8p, C.c_size_t, C.POINTER(C.c_size_t), u64p, C.c_uint32, u32p]),
}


def _load():
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C thaler_study_b200/csrc).  thaler_study_b200 has no CPU fallback."
        )
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class ScbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class ProverClaimMismatch(ScbError):
    """sum_check_protocol::Error::ProverClaimMismatch"""


class NoPolySet(ScbError):
    """sum_check_protocol::Error::NoPolySet"""


def check(rc: int) -> None:
    if rc == SCB_OK:
        return
    msg = (lib.scb_last_error() or b"").decode()
    if rc == SCB_EVERIFY:
        raise ProverClaimMismatch(rc, msg)
    if rc == SCB_ENOPOLY:
        raise NoPolySet(rc, msg)
    raise ScbError(rc, msg)


# ---------------------------------------------------------------- options (scb_set_option; csrc/options.hpp)
def set_option(name: str, value: int) -> None:
    check(lib.scb_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    out = C.c_int64()
    check(lib.scb_get_option(name.encode(), C.byref(out)))
    return out.value


def option_names():
    names, i = [], 0
    while True:
        n = lib.scb_option_name(i)
        if not n:
            return names
        names.append(n.decode())
        i += 1


def reset_options() -> None:
    lib.scb_reset_options()


def options_from_env(environ=None) -> dict:
    """Harness helper (tests, scripts, bench.py): applies ``SCB_<NAME>=<int>`` variables as scb_set_option calls.
    The LIBRARY never reads the environment; only a caller that asks for it here does."""
    environ = os.environ if environ is None else environ
    applied = {}
    for n in option_names():
        v = environ.get("SCB_" + n.upper())
        if v is not None and v.strip().lstrip("-").isdigit():
            set_option(n, int(v))
            applied[n] = int(v)
    return applied
