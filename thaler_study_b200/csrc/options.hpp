// options.hpp -- the library's tuning / diagnostic switches.  They are set through the C ABI only
// (scb_set_option / scb_get_option, include/sumcheck_b200.h): the library never reads the process environment, so a
// host application's variables cannot change what a call does.  Every switch selects between code paths that produce
// the same field elements and transcript bytes (tests/test_gpu_pairs.py, tests/test_gpu_parity.py run them against
// each other); defaults are the measured best on B200.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstring>

namespace scb {

#define SCB_OPTION_LIST(X)                                                                                                  \
    X(packed, 1)               /* prover keeps its private folded tables as packed uint32 (small-prime fields) */           \
    X(pairs, 1)                /* two rounds per pass over the tables (pairs.cuh) */                                        \
    X(pair_resident, 1)        /* pair passes inside one resident kernel; 0: one ordinary launch per pass (ncu) */          \
    X(pair_first_alone, 26)    /* tables of >= 2^n entries: the pass over 8-byte tables is its own launch; 0: never */     \
    X(pair_w21, 3)             /* K = 3, p < 2^21, first pair pass alone: the grid pass writes 21-bit triples for it; that pass: 1 = 3 CTAs/SM, 2 = prefetching (2 CTAs/SM), 3 / 4 = the same with the bilinear two-variable fold (3: measured best); 0: off */ \
    X(pair_w21_alone, 24)      /* ... with the triples the first pair pass is its own launch from 2^n entries already (2^25: 0.62 -> 0.54 ms per proof) */ \
    X(pair_stage, 0)           /* cp.async staging in the pair kernels (measured slower) */                                 \
    X(pair_pipe, 1)            /* resident pair kernel: loads pipelined across tables */                                    \
    X(pair_bps, 0)             /* resident pair kernel: cap on CTAs per SM; 0: occupancy calculator */                      \
    X(grid_pf, 1)              /* grid kernel with the register double buffer */                                            \
    X(grid_tma, 0)             /* grid kernel with the cp.async.bulk + mbarrier ring (measured not faster) */               \
    X(tail_vars, 14)           /* single-CTA resident tail for tables of <= 2^n entries; 0: no resident kernels */          \
    X(persist_vars, 15)        /* grid-wide resident kernel from 2^n entries; 0: off */                                     \
    X(persist_max_generic, 22) /* ... up to 2^n entries for the integer-bound policies */                                   \
    X(persist_trace, 0)        /* print per-round device / host turn-around times of the resident kernels to stderr */     \
    X(bps, 0)                  /* grid-stride kernels: CTAs per SM override; 0: per-kernel default */                       \
    X(bps32, 8)                /* fused fold+message kernel on packed input: CTAs per SM */                                 \
    X(qp32, 4)                 /* fused fold+message kernel on packed input: quads per thread-iteration */                  \
    X(eq_split, 1)             /* eq tables as outer products of shared-memory sub-tables */                                \
    X(mle_lb, 0)               /* MLE evaluation: index bits of the low (shared-memory) eq table; 0: default */             \
    X(mle_u, 0)                /* MLE evaluation: 1 = one group per thread-iteration */                                     \
    X(mle_fused, 1)            /* MLE evaluation as ONE launch (eq sub-tables built per CTA in shared memory) */            \
    X(mle_rows_multi, 1)       /* several evaluations of one table (GKR restrict_poly): row-wise kernel for one-limb fields */ \
    X(mle_rows_bps, 2)         /* ... its CTAs per SM (fewer CTAs: fewer partials for the last CTA to add up) */               \
    X(gkr_multi, 1)            /* GKR layer with challenges up front: up to 4 rounds per pass, no barriers (k_pqs_multi) */  \
    X(gkr_persist, 1)          /* GKR layer phases as one cooperative launch each */                                        \
    X(gkr_tail, 1)             /* GKR layer with challenges up front: the last 11 rounds of a phase in one CTA (k_pqs_tail) */ \
    X(gkr_scatter, 1)          /* GKR phase tables for small-prime fields: thread per gate + 64-bit integer atomics (gkr.cuh) */ \
    X(g4_kernel, 3)            /* 4-limb fused fold+message with a claim: 3 = carry chains + unreduced last products in 544-bit        \
                                  shared-memory accumulators, round 0 included (g4.cuh, K >= 2); 1 = carry chains, one point fewer      \
                                  (14.7 ms per 2^28 x 3 launch); 2 = radix-2^29 lazy carries (g29.cuh: 20.8 ms); 0 = round 1's (18.2 ms) */ \
    X(g4_p0one, 1)             /* g4_kernel 3: use the p = 1 (mod 2^32) variant when the modulus allows (0: always the generic one) */   \
    X(g4_blocks, 2)            /* 4-limb kernels of generations 1-2: variant compiled for 2 or 3 resident CTAs per SM */     \
    X(g4_blocks4, 0)           /* g4_kernel 3, K = 3: 0 = measured defaults (fold kernel 3 CTAs per SM, round 0 two), 1..3 force */ \
    X(tri_tiled, 1)            /* triangle x-phase as a shared-memory tiled field matmul */                                 \
    X(host_pack, 1)            /* narrowing upload of host tables (upload_engine.inc) */                                    \
    X(host_pack_threads, 0)    /* pack threads; 0: hardware threads / local_ranks */                                        \
    X(host_pack_min_vars, 22)  /* narrowing upload from 2^n entries */                                                      \
    X(host_pack_chunk_log2, 20)                                                                                             \
    X(host_pack_raw, 1)        /* device-side narrowing lane; 2: also from pageable memory (tests) */                       \
    X(host_pack_raw_slots, 3)  /* chunks the device-side lane keeps in flight (1..8) */                                     \
    X(host_pack_wire, 21)      /* 21: three 21-bit entries per 64-bit word when p < 2^21; 32: uint32 */                     \
    X(host_pack_prefetch, 4096) /* pack threads prefetch this many bytes ahead of their loads; 0: off (hostpack.hpp) */                     \
    X(host_pack_nt, 0)         /* streaming stores into the staging buffers */                                              \
    X(local_ranks, 1)          /* processes sharing this box's host cores (set by the sharded driver) */                    \
    X(sha_scalar, 0)           /* portable SHA-256 compression instead of the x86 SHA extensions */                         \
    X(strict_verifier, 1)      /* Verifier::round checks the round link in the final round too */    \
    X(consolidate_auto, 16)    /* sharded prover: the slab size (variables) at which consolidate_at = 0 gathers the slabs */

enum Opt : int {
#define X(name, dflt) OPT_##name,
    SCB_OPTION_LIST(X)
#undef X
        OPT_COUNT
};

struct OptionTable {
    std::atomic<int64_t> v[OPT_COUNT];
    OptionTable() { reset(); }
    void reset() {
        int i = 0;
#define X(name, dflt) v[i++].store(dflt, std::memory_order_relaxed);
        SCB_OPTION_LIST(X)
#undef X
    }
};
inline OptionTable& option_table() {
    static OptionTable t;
    return t;
}
inline int64_t opt(Opt o) { return option_table().v[o].load(std::memory_order_relaxed); }
inline int option_index(const char* name) {
    static const char* const names[OPT_COUNT] = {
#define X(name, dflt) #name,
        SCB_OPTION_LIST(X)
#undef X
    };
    for (int i = 0; i < OPT_COUNT; ++i)
        if (std::strcmp(names[i], name) == 0) return i;
    return -1;
}
inline const char* option_name(int i) {
    static const char* const names[OPT_COUNT] = {
#define X(name, dflt) #name,
        SCB_OPTION_LIST(X)
#undef X
    };
    return i >= 0 && i < OPT_COUNT ? names[i] : nullptr;
}

}  // namespace scb
