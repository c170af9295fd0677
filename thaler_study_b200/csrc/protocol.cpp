// protocol.cpp -- level 2 of the C ABI: the host half of the drop-in.  C++ mirror of
//   sum_check_protocol::{Prover, Verifier}      /root/reference/sum-check-protocol/src/lib.rs:73-117,227-331
//   SumCheckPolynomial::to_univariate           per implementor (interpolation + SparsePolynomial conventions)
//   fiat_shamir::{generate_transcript, verify_transcript}   /root/reference/fiat-shamir/src/lib.rs:44-98,123-171
// built ONLY on level 1 (engine.cu): the device produces the (d+1) round sums, everything that decides
// transcript bytes (interpolation, zero-term handling, serialization, SHA-256 hash-to-field) is host code
// here, exactly where the reference keeps it.  In a Rust build this file is replaced by the reference's own
// crates; it exists so that the path is complete and testable without a Rust toolchain.
#include <atomic>
#include <cstring>
#include <memory>
#include <vector>

#include "host/fiat_shamir.hpp"
#include "host/unipoly.hpp"
#include "internal.hpp"
#include "options.hpp"
#include "sumcheck_b200.h"

using namespace scb;

extern "C" int scb_poly_field_impl(const scb_poly* p, const FieldImpl** out);

#define ARG_TRY(cond, msg)        \
    do {                          \
        if (!(cond)) {            \
            set_error("%s", msg); \
            return SCB_EINVAL;    \
        }                         \
    } while (0)
#define RC_TRY(expr)                     \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != SCB_OK) return rc__; \
    } while (0)

static constexpr uint32_t kMaxTerms = 64;  // restrict_poly of a 2^40-gate layer has 41 points

// (d+1) sums -> the univariate::SparsePolynomial each implementor's to_univariate returns
static SparsePoly evals_to_poly(const HostField& F, uint32_t kind, const std::vector<Fe>& ev) {
    const InterpConsts& c = interp_consts(F);
    if (kind == SCB_POLY_MATMUL_G && ev.size() == 3) {
        // matrix-multiplication/src/lib.rs:124-130
        const Fe y[3] = {ev[0], ev[1], ev[2]};
        return interpolate_quadratic_012(F, c, y);
    }
    // triangle-counting/src/lib.rs:128-131, gkr-protocol/src/round_polynomial.rs:86-89 (`p.into()`),
    // and ProductMLE<K>: unique interpolant, Dense -> Sparse
    if (ev.size() < 10 && !c.basis[ev.size()].empty()) return SparsePoly::from_dense(F, lagrange_to_coeffs_cached(F, c, ev));
    return SparsePoly::from_dense(F, lagrange_to_coeffs(F, ev));
}

static int poly_out(const HostField& F, const SparsePoly& sp, uint64_t* degrees, uint64_t* coeffs, uint32_t cap, uint32_t* n_terms) {
    ARG_TRY(degrees && coeffs && n_terms, "null argument");
    ARG_TRY(sp.coeffs.size() <= cap, "term buffer too small");
    for (size_t i = 0; i < sp.coeffs.size(); ++i) {
        degrees[i] = sp.coeffs[i].first;
        F.store(sp.coeffs[i].second, coeffs + i * F.n);
    }
    *n_terms = (uint32_t)sp.coeffs.size();
    return SCB_OK;
}
static SparsePoly poly_in(const HostField& F, const uint64_t* degrees, const uint64_t* coeffs, uint32_t n_terms) {
    SparsePoly sp;
    for (uint32_t i = 0; i < n_terms; ++i) {
        Fe c;
        F.load(coeffs + (size_t)i * F.n, c);
        sp.coeffs.emplace_back(degrees[i], c);
    }
    return sp;
}

static int device_round_poly(const scb_poly* g, SparsePoly* out) {
    const FieldImpl* fi;
    RC_TRY(scb_poly_field_impl(g, &fi));
    uint32_t np = 0, kind = 0;
    RC_TRY(scb_poly_n_points(g, &np));
    RC_TRY(scb_poly_kind_of(g, &kind));
    uint64_t w[8 * kHostMaxLimbs];
    RC_TRY(scb_poly_round_evals(g, np, w));
    std::vector<Fe> ev(np);
    for (uint32_t i = 0; i < np; ++i) fi->h.load(w + (size_t)i * fi->h.n, ev[i]);
    *out = evals_to_poly(fi->h, kind, ev);
    return SCB_OK;
}

extern "C" int scb_poly_to_univariate(const scb_poly* p, uint64_t* degrees, uint64_t* coeffs, uint32_t cap_terms, uint32_t* n_terms) {
    ARG_TRY(p, "null argument");
    const FieldImpl* fi;
    RC_TRY(scb_poly_field_impl(p, &fi));
    SparsePoly sp;
    RC_TRY(device_round_poly(p, &sp));
    return poly_out(fi->h, sp, degrees, coeffs, cap_terms, n_terms);
}

extern "C" int scb_evals_to_univariate(const scb_field* f, uint32_t kind, const uint64_t* evals, uint32_t n_points, uint64_t* degrees,
                                       uint64_t* coeffs, uint32_t cap_terms, uint32_t* n_terms) {
    ARG_TRY(f && evals, "null argument");
    ARG_TRY(n_points >= 1 && n_points <= kMaxTerms, "n_points out of range");
    const HostField& F = f->impl->h;
    std::vector<Fe> ev(n_points);
    for (uint32_t i = 0; i < n_points; ++i) {
        F.load(evals + (size_t)i * F.n, ev[i]);
        ARG_TRY(F.is_canonical(ev[i]), "evaluation is not a canonical field element");
    }
    return poly_out(F, evals_to_poly(F, kind, ev), degrees, coeffs, cap_terms, n_terms);
}
extern "C" int scb_unipoly_serialize(const scb_field* f, const uint64_t* degrees, const uint64_t* coeffs, uint32_t n_terms, uint8_t* out,
                                     size_t cap, size_t* out_len) {
    ARG_TRY(f && out && out_len && ((degrees && coeffs) || n_terms == 0), "null argument");
    const HostField& F = f->impl->h;
    std::vector<uint8_t> bytes;
    poly_in(F, degrees, coeffs, n_terms).serialize(F, bytes);
    ARG_TRY(bytes.size() <= cap, "output buffer too small");
    std::memcpy(out, bytes.data(), bytes.size());
    *out_len = bytes.size();
    return SCB_OK;
}
extern "C" int scb_unipoly_evaluate(const scb_field* f, const uint64_t* degrees, const uint64_t* coeffs, uint32_t n_terms,
                                    const uint64_t* x, uint64_t* out_elem) {
    ARG_TRY(f && x && out_elem && ((degrees && coeffs) || n_terms == 0), "null argument");
    const HostField& F = f->impl->h;
    Fe xe;
    F.load(x, xe);
    F.store(poly_in(F, degrees, coeffs, n_terms).evaluate(F, xe), out_elem);
    return SCB_OK;
}
extern "C" int scb_hash_to_field(const scb_field* f, const uint8_t* msg, size_t len, uint64_t* out_elem) {
    ARG_TRY(f && out_elem && (msg || len == 0), "null argument");
    const HostField& F = f->impl->h;
    F.store(hash_to_field(F, msg, len), out_elem);
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------ Prover
static int packed_enabled() { return opt(OPT_packed) != 0; }
struct scb_prover {
    scb_poly* g = nullptr;      // g: P
    Fe c_1;                     // c_1: F
    std::vector<Fe> r;          // r: Vec<F>
    uint32_t num_vars = 0;
    const FieldImpl* fi = nullptr;
    uint32_t kind = 0, np = 0;
    // the message sent last, while the per-round path is in use: g_{j-1}(r_{j-1}) is the claim the next message must sum to
    bool have_last = false;
    SparsePoly last_msg;
    bool have_round0 = false;   // round-0 sums computed by Prover::new (they also yield c_1)
    std::vector<Fe> round0;
    // small-prime fields: Prover::new's pass accumulates the (K+1)^2 grid H[a][b] (pairs.cuh) instead of the K+1
    // line sums; it yields c_1, g_1 and -- once r_1 is known -- g_2 without another pass over the tables
    bool have_grid = false;
    std::vector<Fe> grid;       // a-major
    // sharded prover (multi-GPU): g is this rank's slab while `sharded`; afterwards the consolidated table
    scb_peers* peers = nullptr;
    uint32_t consolidate_at = 0;
    bool sharded = false;
    ~scb_prover() { scb_poly_free(g); }
};

struct PeersScope {  // makes `peers` the current exchange group of this thread for the lifetime of the scope
    explicit PeersScope(scb_peers* p) : active(p != nullptr) {
        if (active) scb_peers_set_current(p);
    }
    ~PeersScope() {
        if (active) scb_peers_set_current(nullptr);
    }
    bool active;
};
// slabs small enough: all-gather them once (P2P stores into every peer's window) and continue replicated
static int maybe_consolidate(scb_prover* p, bool force = false) {
    if (!p->sharded) return SCB_OK;
    uint32_t lv = 0;
    RC_TRY(scb_poly_num_vars(p->g, &lv));
    if (lv > p->consolidate_at && !force) return SCB_OK;
    scb_poly* full = nullptr;
    RC_TRY(scb_peers_gather_poly(p->peers, p->g, &full));
    scb_poly_free(p->g);
    p->g = full;
    p->sharded = false;
    return SCB_OK;
}

static int prover_new_impl(const scb_poly* g, scb_peers* peers, uint32_t world, uint32_t consolidate_at, scb_prover** out) {
    // Prover::new :88-97 -- c_1 = g.to_evaluations().into_iter().sum(), computed on the device
    ARG_TRY(g && out, "null argument");
    auto p = std::make_unique<scb_prover>();
    RC_TRY(scb_poly_clone(g, &p->g));
    RC_TRY(scb_poly_allow_packed(p->g, packed_enabled()));  // the prover's folded tables are private to it
    RC_TRY(scb_poly_field_impl(g, &p->fi));
    RC_TRY(scb_poly_kind_of(g, &p->kind));
    RC_TRY(scb_poly_n_points(g, &p->np));
    RC_TRY(scb_poly_num_vars(g, &p->num_vars));
    if (peers && world > 1) {
        ARG_TRY(p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G, "only product polynomials shard");
        uint32_t lg = 0;
        while ((1u << lg) < world) ++lg;
        p->num_vars += lg;  // the top log2(world) variables are the rank index
        p->peers = peers;
        // consolidate_at = 0: the library picks the threshold.  Above it every pass costs one exchange (~6 us flag
        // round-trip over NVLink, profiles/r01_pairs.md); below it the passes run replicated on all ranks with no
        // exchange but on world x the data, and the gather itself moves world x K x 4 x 2^lv bytes into every window.
        p->consolidate_at = consolidate_at < 1 ? (opt(OPT_consolidate_auto) > 1 ? (uint32_t)opt(OPT_consolidate_auto) : 16) : consolidate_at;
        if (consolidate_at < 1) {  // the library's own choice must fit the window: world x K tables x 2^lv entries of 8 n_limbs bytes
            size_t cap = 0;
            uint32_t K = 0;
            RC_TRY(scb_peers_gather_capacity(peers, &cap));
            RC_TRY(scb_poly_n_tables(g, &K));
            while (p->consolidate_at > 3 && ((size_t)world * K * 8 * p->fi->d.n << p->consolidate_at) > cap) --p->consolidate_at;
        }
        p->sharded = true;
        RC_TRY(maybe_consolidate(p.get()));
    }
    const HostField& F = p->fi->h;
    if (p->num_vars >= 2 && pair_passes_ok(p->g)) {
        // g_1(X) = H(X,0) + H(X,1) and c_1 = g_1(0) + g_1(1), from the grid pass (sharded: summed over the peer
        // GPUs by the kernel's finishing thread)
        uint64_t w[32];
        {
            // when the first pair pass will run as its own launch (the condition of scb_fs_generate_transcript), the grid pass
            // may leave it a narrower copy of the tables (21-bit triples, pairs.cuh)
            uint32_t live = 0;
            RC_TRY(scb_poly_num_vars(p->g, &live));
            uint32_t first_alone = (uint32_t)opt(OPT_pair_first_alone);
            // with the triples the stand-alone pass pays off on smaller tables (option pair_w21_alone); whether they apply to this
            // polynomial is the grid pass's decision, and scb_fs_generate_transcript follows it (poly_has_w21)
            if (first_alone != 0 && opt(OPT_pair_w21) != 0 && (uint32_t)opt(OPT_pair_w21_alone) < first_alone) first_alone = (uint32_t)opt(OPT_pair_w21_alone);
            const bool resident = opt(OPT_pair_resident) != 0;  // 0: every pass is an ordinary launch (a sharded proof consolidates first)
            const bool alone_next = !poly_is_packed(p->g) && live >= 4 &&
                                    (resident ? first_alone != 0 && live >= first_alone && (!p->sharded || live > p->consolidate_at) : !p->sharded);
            PeersScope scope(p->sharded ? p->peers : nullptr);
            RC_TRY(poly_grid_evals_ex(p->g, w, alone_next));
        }
        const uint32_t np = p->np;
        p->grid.resize((size_t)np * np);
        for (uint32_t i = 0; i < np * np; ++i) F.load(w + i, p->grid[i]);
        p->have_grid = true;
        p->round0.resize(np);
        for (uint32_t x = 0; x < np; ++x) p->round0[x] = F.add(p->grid[x * np], p->grid[x * np + 1]);
        p->have_round0 = true;
        p->c_1 = F.add(p->round0[0], p->round0[1]);
    } else if (p->num_vars >= 1) {
        // sum over the hypercube = g_1(0) + g_1(1): the round-0 message pass also yields c_1, so the
        // tables are streamed once here and not again by round(_, 0)  (same field elements either way)
        uint64_t w[8 * kHostMaxLimbs];
        {
            PeersScope scope(p->sharded ? p->peers : nullptr);
            RC_TRY(scb_poly_round_evals(p->g, p->np, w));
        }
        p->round0.resize(p->np);
        for (uint32_t i = 0; i < p->np; ++i) F.load(w + (size_t)i * F.n, p->round0[i]);
        p->have_round0 = true;
        p->c_1 = F.add(p->round0[0], p->round0[1]);
    } else {
        uint64_t w[kHostMaxLimbs];
        RC_TRY(scb_poly_sum(g, w));
        F.load(w, p->c_1);
    }
    p->r.reserve(p->num_vars);
    *out = p.release();
    return SCB_OK;
}
extern "C" int scb_prover_new(const scb_poly* g, scb_prover** out) { return prover_new_impl(g, nullptr, 1, 0, out); }
extern "C" int scb_prover_new_sharded(const scb_poly* slab, scb_peers* peers, uint32_t world, uint32_t consolidate_at, scb_prover** out) {
    ARG_TRY(peers || world <= 1, "null peers");
    return prover_new_impl(slab, peers, world, consolidate_at, out);
}
extern "C" void scb_prover_free(scb_prover* p) { delete p; }
extern "C" int scb_prover_c_1(const scb_prover* p, uint64_t* out_elem) {
    ARG_TRY(p && out_elem, "null argument");
    p->fi->h.store(p->c_1, out_elem);
    return SCB_OK;
}
extern "C" int scb_prover_num_vars(const scb_prover* p, uint32_t* out) {
    ARG_TRY(p && out, "null argument");
    *out = p->num_vars;
    return SCB_OK;
}

// Prover::round :105-112.  For j != 0 the fold and the message are ONE fused kernel.
static int prover_round(scb_prover* p, const Fe* r_prev, uint32_t j, SparsePoly* out) {
    const HostField& F = p->fi->h;
    uint64_t w[8 * kHostMaxLimbs];
    if (j != 0) {
        uint64_t rw[kHostMaxLimbs];
        F.store(*r_prev, rw);
        p->r.push_back(*r_prev);
        RC_TRY(maybe_consolidate(p));
        scb_poly* next = nullptr;
        PeersScope scope(p->sharded ? p->peers : nullptr);  // per-round exchange inside the kernel while sharded
        if (p->have_last) {  // g_j(0) + g_j(1) = g_{j-1}(r_{j-1}) is known before the pass: one point fewer to accumulate
            uint64_t cw[kHostMaxLimbs];
            F.store(p->last_msg.evaluate(F, *r_prev), cw);
            RC_TRY(scb_poly_fix_and_round_evals_claim(p->g, rw, cw, p->np, &next, w));
        } else {
            RC_TRY(scb_poly_fix_and_round_evals(p->g, rw, p->np, &next, w));
        }
        scb_poly_free(p->g);
        p->g = next;
    } else if (p->have_round0) {
        *out = evals_to_poly(F, p->kind, p->round0);
        p->last_msg = *out;
        p->have_last = true;
        return SCB_OK;
    } else {
        RC_TRY(scb_poly_round_evals(p->g, p->np, w));
    }
    std::vector<Fe> ev(p->np);
    for (uint32_t i = 0; i < p->np; ++i) F.load(w + (size_t)i * F.n, ev[i]);
    *out = evals_to_poly(F, p->kind, ev);
    p->last_msg = *out;
    p->have_last = true;
    return SCB_OK;
}
extern "C" int scb_prover_round(scb_prover* p, const uint64_t* r_prev, uint32_t j, uint64_t* degrees, uint64_t* coeffs,
                                uint32_t cap_terms, uint32_t* n_terms) {
    ARG_TRY(p && (r_prev || j == 0), "null argument");
    Fe r;
    if (j != 0) p->fi->h.load(r_prev, r);
    SparsePoly sp;
    RC_TRY(prover_round(p, &r, j, &sp));
    return poly_out(p->fi->h, sp, degrees, coeffs, cap_terms, n_terms);
}

// ------------------------------------------------------------------------------------------ Verifier
struct scb_verifier {
    std::shared_ptr<FieldImpl> f;
    uint32_t n = 0;                    // n: usize
    Fe c_1;                            // c_1: F
    std::vector<SparsePoly> g_part;    // g_part
    std::vector<Fe> r;                 // r
    scb_poly* g = nullptr;             // g: Option<P>
    ~scb_verifier() { scb_poly_free(g); }
};

extern "C" int scb_verifier_new(const scb_field* f, uint32_t n, const scb_poly* g, scb_verifier** out) {
    ARG_TRY(f && out, "null argument");
    auto v = std::make_unique<scb_verifier>();
    v->f = f->impl;
    v->n = n;
    if (g) RC_TRY(scb_poly_clone(g, &v->g));
    *out = v.release();
    return SCB_OK;
}
extern "C" void scb_verifier_free(scb_verifier* v) { delete v; }
extern "C" int scb_verifier_set_c_1(scb_verifier* v, const uint64_t* c_1) {
    ARG_TRY(v && c_1, "null argument");
    v->f->h.load(c_1, v->c_1);
    return SCB_OK;
}

// Verifier::round :278-330
static int verifier_round(scb_verifier* v, const SparsePoly& g_j, const Fe& r_j, int* final_round, int* accepted) {
    const HostField& F = v->f->h;
    *final_round = 0;
    *accepted = 0;
    // The reference's last-round branch (:298-310) checks g_n(r_n) == g(r) but NOT g_n(0) + g_n(1) == g_{n-1}(r_{n-1}),
    // and for n == 1 its first-round branch (:284-297) returns before the oracle is ever evaluated: a prover can send
    // self-consistent g_1..g_{n-1} for a false c_1 followed by the honest g_n and be accepted.  With option
    // strict_verifier (default 1) both checks are made; honest transcripts and every byte of them are unchanged.
    // strict_verifier = 0 is the reference's literal behaviour (tests/test_host_abi.py shows the difference).
    const bool strict = opt(OPT_strict_verifier) != 0;
    auto final_check = [&](const SparsePoly& g_last, const Fe& r_last) -> int {
        if (!v->g) {
            set_error("verifier has no oracle access to the polynomial");
            return SCB_ENOPOLY;
        }
        std::vector<uint64_t> pt((size_t)v->n * F.n);
        for (uint32_t i = 0; i < v->n; ++i) F.store(v->r[i], &pt[(size_t)i * F.n]);
        uint64_t w[kHostMaxLimbs];
        RC_TRY(scb_poly_evaluate(v->g, pt.data(), v->n, w));  // g.evaluate(&self.r) on the device (K4, LSB-first)
        Fe oracle;
        F.load(w, oracle);
        *final_round = 1;
        *accepted = g_last.evaluate(F, r_last) == oracle ? 1 : 0;
        return SCB_OK;
    };
    if (strict && v->r.size() >= (size_t)v->n) {
        set_error("verifier has already run its %u rounds", v->n);
        return SCB_EINVAL;
    }
    if (v->r.empty()) {  // first round :284-297
        Fe evaluation = F.add(g_j.evaluate(F, F.zero()), g_j.evaluate(F, F.one()));
        if (v->c_1 != evaluation) {
            set_error("prover claim mismatches evaluation (start)");
            return SCB_EVERIFY;
        }
        v->g_part.push_back(g_j);
        v->r.push_back(r_j);
        if (strict && v->n == 1) return final_check(g_j, r_j);
        return SCB_OK;
    } else if (v->r.size() == (size_t)v->n - 1) {  // last round :298-310
        if (strict) {
            Fe prev = v->g_part.back().evaluate(F, v->r.back());
            Fe evaluation = F.add(g_j.evaluate(F, F.zero()), g_j.evaluate(F, F.one()));
            if (prev != evaluation) {
                set_error("prover claim mismatches evaluation (round %zu)", v->r.size());
                return SCB_EVERIFY;
            }
        }
        v->r.push_back(r_j);
        return final_check(g_j, r_j);
    } else {  // j-th round :311-329
        Fe prev = v->g_part.back().evaluate(F, v->r.back());
        Fe evaluation = F.add(g_j.evaluate(F, F.zero()), g_j.evaluate(F, F.one()));
        if (prev != evaluation) {
            set_error("prover claim mismatches evaluation (round %zu)", v->r.size());
            return SCB_EVERIFY;
        }
        v->g_part.push_back(g_j);
        v->r.push_back(r_j);
        return SCB_OK;
    }
}
extern "C" int scb_verifier_round(scb_verifier* v, const uint64_t* degrees, const uint64_t* coeffs, uint32_t n_terms, const uint64_t* r_j,
                                  int* final_round, int* accepted) {
    ARG_TRY(v && r_j && final_round && accepted && ((degrees && coeffs) || n_terms == 0), "null argument");
    const HostField& F = v->f->h;
    Fe r;
    F.load(r_j, r);
    return verifier_round(v, poly_in(F, degrees, coeffs, n_terms), r, final_round, accepted);
}

// ------------------------------------------------------------------------------------------ fiat-shamir
// Small tables are finished by the single-CTA tail kernel, large ones by the grid-wide resident kernel; the engine
// decides (resident_rounds_ok, engine.cu).
struct TailCtx {
    const HostField* F;
    uint32_t kind;
    std::vector<uint8_t>* hash_input;
    FsChain* chain;
    uint64_t* offsets;
    uint32_t j;   // protocol round of the tail's round 0
    uint32_t np;  // sums per round
    std::vector<uint64_t> used;  // challenges consumed by the tail's folds, in order (n limbs each)
};
// A resident kernel that loses lock-step with the host (a profiler that serialises kernel and host, or a host thread
// that was descheduled for longer than the kernel's 250 ms patience) costs a quarter of a second before the per-round
// fallback takes over.  After such a failure the resident kernels are skipped for a number of proofs that doubles with
// every consecutive failure (8, 16, ... 4096): a profiler pays the stall a handful of times per run, a one-off hiccup
// costs eight slower proofs.  Sharded provers must take the same path on every rank, so they disable for good.
static std::atomic<uint32_t> g_resident_skip{0}, g_resident_backoff{8};
static bool g_resident_off = false;
static bool resident_allowed(bool begin_proof) {
    if (g_resident_off) return false;
    uint32_t s = g_resident_skip.load();
    if (s == 0) return true;
    if (begin_proof) g_resident_skip.store(s - 1);
    return false;
}
static void resident_failed(bool sharded) {
    if (sharded) g_resident_off = true;
    const uint32_t b = g_resident_backoff.load();
    g_resident_skip.store(b);
    g_resident_backoff.store(b < 4096 ? b * 2 : b);
}
static void resident_succeeded() { g_resident_backoff.store(8); }
// one tail round on the host: sums -> message polynomial -> bytes -> hash chain -> next challenge
static int tail_round_cb(void* user, uint32_t t, const uint64_t* evals, uint64_t* next_r) {
    TailCtx* tc = (TailCtx*)user;
    const HostField& F = *tc->F;
    std::vector<Fe> ev(tc->np);
    for (size_t i = 0; i < ev.size(); ++i) F.load(evals + i * F.n, ev[i]);
    const size_t before = tc->hash_input->size();
    evals_to_poly(F, tc->kind, ev).serialize(F, *tc->hash_input);
    tc->chain->absorb(tc->hash_input->data() + before, tc->hash_input->size() - before);
    tc->offsets[tc->j + t + 1] = tc->hash_input->size();
    F.store(tc->chain->challenge(), next_r);
    tc->used.insert(tc->used.end(), next_r, next_r + F.n);
    return SCB_OK;
}

// Two rounds per pass (pairs.cuh): host half.  From a grid H (a-major, (K+1)^2 values) the next two messages are
// g(X) = H(X,0) + H(X,1) and, with r = hash(transcript so far), g'(Y) = H(r, Y): column Y of the grid interpolated
// in `a` on the nodes 0..K and evaluated at r -- exact field arithmetic, hence the reference's field elements.
struct PairCtx {
    const HostField* F;
    uint32_t kind, np;
    std::vector<uint8_t>* hash_input;
    FsChain* chain;
    uint64_t* offsets;
    uint32_t msgs;                // messages emitted so far (g_1 .. g_msgs)
    std::vector<uint64_t> used;   // challenges derived so far, in order (one limb each)

    void emit(const std::vector<Fe>& ev) {
        const size_t before = hash_input->size();
        evals_to_poly(*F, kind, ev).serialize(*F, *hash_input);
        chain->absorb(hash_input->data() + before, hash_input->size() - before);
        offsets[++msgs] = hash_input->size();
    }
    Fe next_challenge() {
        Fe r = chain->challenge();
        uint64_t w[kHostMaxLimbs];
        F->store(r, w);
        used.push_back(w[0]);
        return r;
    }
    void emit_first(const std::vector<Fe>& H) {
        std::vector<Fe> ev(np);
        for (uint32_t x = 0; x < np; ++x) ev[x] = F->add(H[x * np], H[x * np + 1]);
        emit(ev);
    }
    // g'(Y) = H(r, Y) = sum_a L_a(r) H[a][Y] with the Lagrange basis on the nodes 0..np-1: L_a(r) =
    // prod_{c != a} (r - c) / prod_{c != a} (a - c).  The np weights are computed once per message (prefix/suffix
    // products, denominators inverted once per proof), then every column costs np multiplications.
    std::vector<Fe> den_inv;  // 1 / prod_{c != a} (a - c)
    void emit_second(const std::vector<Fe>& H, const Fe& r) {
        if (den_inv.empty()) {
            den_inv.resize(np);
            for (uint32_t a = 0; a < np; ++a) {
                Fe d = F->one();
                for (uint32_t c = 0; c < np; ++c)
                    if (c != a) d = F->mul(d, F->sub(F->from_u64(a), F->from_u64(c)));
                den_inv[a] = F->inverse(d);
            }
        }
        Fe diff[8], pre[9], suf[9], w[8];
        for (uint32_t c = 0; c < np; ++c) diff[c] = F->sub(r, F->from_u64(c));
        pre[0] = F->one();
        for (uint32_t c = 0; c < np; ++c) pre[c + 1] = F->mul(pre[c], diff[c]);
        suf[np] = F->one();
        for (uint32_t c = np; c-- > 0;) suf[c] = F->mul(suf[c + 1], diff[c]);
        for (uint32_t a = 0; a < np; ++a) w[a] = F->mul(F->mul(pre[a], suf[a + 1]), den_inv[a]);
        std::vector<Fe> ev(np);
        for (uint32_t y = 0; y < np; ++y) {
            Fe acc = F->zero();
            for (uint32_t a = 0; a < np; ++a) acc = F->add(acc, F->mul(w[a], H[a * np + y]));
            ev[y] = acc;
        }
        emit(ev);
    }
};
static int pair_pass_cb(void* user, uint32_t, uint32_t n_vals, const uint64_t* vals, uint64_t* next_pair) {
    PairCtx* pc = (PairCtx*)user;
    const HostField& F = *pc->F;
    std::vector<Fe> v(n_vals);
    for (uint32_t i = 0; i < n_vals; ++i) F.load(vals + i, v[i]);
    next_pair[0] = next_pair[1] = 0;
    if (n_vals == pc->np) {  // a single variable was left: its line sums are the last message
        pc->emit(v);
        return SCB_OK;
    }
    pc->emit_first(v);
    const Fe ra = pc->next_challenge();
    pc->emit_second(v, ra);
    const Fe rb = pc->next_challenge();
    uint64_t w[kHostMaxLimbs] = {0};
    F.store(ra, w);
    next_pair[0] = w[0];
    F.store(rb, w);
    next_pair[1] = w[0];
    return SCB_OK;
}

extern "C" int scb_fs_generate_transcript(scb_prover* p, uint8_t* out, size_t cap, size_t* out_len, uint64_t* offsets) {
    // fiat-shamir/src/lib.rs:75-98 with InteractiveProver for Prover (:44-66)
    ARG_TRY(p && out && out_len && offsets, "null argument");
    const HostField& F = p->fi->h;
    std::vector<uint8_t> hash_input;  // == concatenation of all messages so far
    FsChain chain(F);                 // SHA-256 state over the same bytes, advanced incrementally
    SparsePoly sp;
    Fe dummy;
    // g_1 = (c_1, round(F::one(), 0)).serialize_uncompressed()
    RC_TRY(prover_round(p, &dummy, 0, &sp));
    F.serialize(p->c_1, hash_input);
    sp.serialize(F, hash_input);
    chain.absorb(hash_input.data(), hash_input.size());
    offsets[0] = 0;
    offsets[1] = hash_input.size();
    const bool product = p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G;
    uint32_t j = 1;
    const bool resident_ok = resident_allowed(true);  // one decision per proof
    if (p->have_grid && resident_ok && p->num_vars >= 2 && pair_passes_ok(p->g)) {
        // two rounds per pass: g_2 comes from Prover::new's grid, then every pass of the resident kernel folds two
        // variables and returns the grid for the next two messages.  Invariant at the top of the loop: p->g is folded
        // by used[0 .. size-2), the last two challenges are the pair the next pass folds by, and the messages for
        // those two variables are already out.
        PairCtx pc{&F, p->kind, p->np, &hash_input, &chain, offsets, 1, {}};
        const Fe r1 = pc.next_challenge();
        pc.emit_second(p->grid, r1);
        int rc = SCB_OK;
        size_t base = 0;  // challenges already folded into p->g
        if (p->num_vars > 2) pc.next_challenge();
        // option pair_resident = 0: every pass is an ordinary launch (k_pair_pass_sp) -- the same pass bodies, visible to
        // ncu, which cannot run the resident kernel (it serialises kernel and host)
        const bool resident = opt(OPT_pair_resident) != 0;
        while (pc.msgs < p->num_vars && rc == SCB_OK) {
            base = pc.used.size() - 2;
            const uint64_t* pair = &pc.used[base];
            uint32_t live = 0, max_passes = 0;
            RC_TRY(maybe_consolidate(p, !resident));
            RC_TRY(scb_poly_num_vars(p->g, &live));
            // The pass over the caller's 8-byte tables runs as its own launch when the tables are large (>= 2^26
            // entries): the stand-alone kernel fits 3 CTAs per SM (80 registers), the resident one 2; without the
            // 8-byte path the resident kernel has the registers to pipeline its packed loads across tables; and the
            // extra launch + wait costs less than the two gains (option pair_first_alone = 0: everything in the
            // resident kernel; n: threshold 2^n).
            const uint32_t first_alone = (uint32_t)opt(OPT_pair_first_alone);
            const bool alone_now = resident && ((first_alone != 0 && live >= first_alone) || poly_has_w21(p->g)) && !poly_is_packed(p->g) && live >= 4 &&
                                   (!p->sharded || (live > p->consolidate_at && live >= 4));  // sharded: two local variables stay
            if (!resident || alone_now) {
                if (live < 4) {  // too small for a grid pass: the per-round path below finishes the proof
                    rc = SCB_ETAIL;
                    break;
                }
                uint64_t w[32], next_pair[2];
                scb_poly* next = nullptr;
                {
                    PeersScope scope(p->sharded ? p->peers : nullptr);  // sharded: exchange inside the kernel's finish
                    RC_TRY(scb_poly_pair_pass(p->g, pair, pair + 1, &next, w));
                }
                scb_poly_free(p->g);
                p->g = next;
                RC_TRY(pair_pass_cb(&pc, 0, p->np * p->np, w, next_pair));
                continue;
            }
            if (p->sharded) {
                // sharded passes stop at the consolidation point and always leave two local variables for the grid
                for (uint32_t mm = live; mm > p->consolidate_at && mm >= 4; mm -= 2) ++max_passes;
                if (max_passes == 0) {
                    RC_TRY(maybe_consolidate(p, true));
                    RC_TRY(scb_poly_num_vars(p->g, &live));
                }
            }
            const bool was_sharded = p->sharded;
            uint32_t done = 0;
            scb_poly* folded = nullptr;
            {
                PeersScope scope(was_sharded ? p->peers : nullptr);
                rc = scb_poly_resident_pairs(p->g, pair, pair + 1, max_passes, pair_pass_cb, &pc, &done, was_sharded ? &folded : nullptr);
            }
            if (rc == SCB_OK && was_sharded) {  // carry on from the slab the kernel left behind: consolidation comes next
                scb_poly_free(p->g);
                p->g = folded;
            }
        }
        if (rc != SCB_OK && rc != SCB_ETAIL) return rc;
        if (rc == SCB_OK && resident) resident_succeeded();
        j = pc.msgs;
        p->have_last = false;  // the messages above did not go through prover_round
        if (rc == SCB_ETAIL) {
            // lock-step lost (e.g. a profiler serialises kernel and host): keep the messages that are out, fold the
            // tables by the challenges they were derived with and carry on with one launch per round
            if (resident) resident_failed(p->peers != nullptr);
            scb_poly* refolded = nullptr;
            RC_TRY(scb_poly_fix_variables(p->g, pc.used.data() + base, (uint32_t)(j - 1 - base), &refolded));
            scb_poly_free(p->g);
            p->g = refolded;
        }
    }
    while (j < p->num_vars) {
        Fe r_j = chain.challenge();  // == hash_to_field(hash_input)
        RC_TRY(maybe_consolidate(p));
        uint32_t live = 0;  // variables of the table about to be folded
        RC_TRY(scb_poly_num_vars(p->g, &live));
        if (product && resident_ok && resident_allowed(false) && live >= 2 && resident_rounds_ok(p->g, p->sharded)) {
            // all remaining rounds (sharded: all rounds up to the consolidation point, with the per-round exchange
            // inside the kernel) in one resident kernel, challenges through a mailbox
            TailCtx tc{&F, p->kind, &hash_input, &chain, offsets, j, p->np, {}};
            uint64_t rw[kHostMaxLimbs];
            F.store(r_j, rw);
            tc.used.insert(tc.used.end(), rw, rw + F.n);
            const bool was_sharded = p->sharded;
            const uint32_t max_rounds = was_sharded ? live - p->consolidate_at : 0;
            uint32_t done = 0;
            scb_poly* folded = nullptr;
            int rc;
            {
                PeersScope scope(was_sharded ? p->peers : nullptr);
                rc = scb_poly_resident_rounds(p->g, rw, p->np, max_rounds, tail_round_cb, &tc, &done, was_sharded ? &folded : nullptr);
            }
            if (rc == SCB_OK) {
                resident_succeeded();
                if (!was_sharded) break;
                scb_poly_free(p->g);  // carry on from the slab the kernel left behind: consolidation comes next
                p->g = folded;
                j += done;
                p->have_last = false;
                continue;
            }
            if (rc != SCB_ETAIL) return rc;
            // lock-step lost (e.g. a profiler serialises kernel and host): keep what was done, fold the tables by the
            // challenges already consumed and carry on with one launch per round
            resident_failed(p->peers != nullptr);
            if (done > 0) {
                scb_poly* refolded = nullptr;
                RC_TRY(scb_poly_fix_variables(p->g, tc.used.data(), done, &refolded));
                scb_poly_free(p->g);
                p->g = refolded;
                j += done;
                p->have_last = false;
            }
            continue;
        }
        RC_TRY(prover_round(p, &r_j, j, &sp));
        const size_t before = hash_input.size();
        sp.serialize(F, hash_input);
        chain.absorb(hash_input.data() + before, hash_input.size() - before);
        offsets[j + 1] = hash_input.size();
        ++j;
    }
    ARG_TRY(hash_input.size() <= cap, "transcript buffer too small");
    std::memcpy(out, hash_input.data(), hash_input.size());
    *out_len = hash_input.size();
    return SCB_OK;
}

extern "C" int scb_fs_verify_transcript(scb_verifier* v, const uint8_t* transcript, size_t transcript_len, const uint64_t* offsets,
                                        uint32_t n_msgs, int* accepted) {
    // fiat-shamir/src/lib.rs:123-143 with InteractiveVerifier for Verifier (:151-171).  The transcript comes from an
    // untrusted prover: offsets must start at 0, be non-decreasing and stay inside transcript_len.
    ARG_TRY(v && transcript && offsets && accepted, "null argument");
    const HostField& F = v->f->h;
    *accepted = 0;
    ARG_TRY(n_msgs == 0 || offsets[0] == 0, "offsets[0] must be 0");
    for (uint32_t j = 0; j < n_msgs; ++j) ARG_TRY(offsets[j] <= offsets[j + 1] && offsets[j + 1] <= transcript_len, "message offsets out of range");
    const bool strict = opt(OPT_strict_verifier) != 0;
    bool reached_final = false;
    for (uint32_t j = 0; j < n_msgs; ++j) {
        const uint8_t* msg = transcript + offsets[j];
        const size_t len = offsets[j + 1] - offsets[j];
        Fe r_j = hash_to_field(F, transcript, offsets[j + 1]);  // hash of g_1 || ... || g_j
        size_t off = 0;
        if (j == 0) {
            if (len < F.ser_bytes()) {
                set_error("Codec error");
                return SCB_EINVAL;
            }
            Fe c_1;
            if (!F.deserialize(msg, c_1)) {
                set_error("Codec error");
                return SCB_EINVAL;
            }
            v->c_1 = c_1;
            off = F.ser_bytes();
        }
        SparsePoly g_j;
        size_t used = SparsePoly::deserialize(F, msg + off, len - off, g_j);
        if (used == 0 || (strict && used != len - off)) {
            set_error("Codec error");
            return SCB_EINVAL;
        }
        int fin = 0, acc = 0;
        RC_TRY(verifier_round(v, g_j, r_j, &fin, &acc));
        if (fin) reached_final = true;
        if (fin && !acc) return SCB_OK;  // *accepted stays 0
        if (j == 0 && !fin) continue;    // :155-161 returns Ok(true) for the first message
    }
    // the reference accepts whatever prefix it was given (:131-141 loops over transcript.g.len()); a truncated
    // transcript never reaches the oracle check, so strict mode rejects it
    *accepted = (!strict || reached_final) ? 1 : 0;
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------ transcript object
// The Fiat-Shamir chain of fiat-shamir/src/lib.rs:75-98 as an explicit state machine, for provers whose
// round sums arrive in pieces (the sharded multi-GPU prover: one row of partial sums per rank).
struct scb_transcript {
    std::shared_ptr<FieldImpl> f;
    uint32_t kind = 0;
    std::vector<uint8_t> bytes;      // g_1 || g_2 || ...  (== hash_input)
    std::vector<uint64_t> offsets{0};
    Fe c_1;
    std::unique_ptr<FsChain> chain;
};
extern "C" int scb_transcript_new(const scb_field* f, uint32_t kind, scb_transcript** out) {
    ARG_TRY(f && out, "null argument");
    auto t = std::make_unique<scb_transcript>();
    t->f = f->impl;
    t->kind = kind;
    t->chain = std::make_unique<FsChain>(t->f->h);
    *out = t.release();
    return SCB_OK;
}
extern "C" void scb_transcript_free(scb_transcript* t) { delete t; }
extern "C" int scb_transcript_absorb_round(scb_transcript* t, const uint64_t* parts, uint32_t n_parts, uint32_t n_points, uint64_t* out_r) {
    ARG_TRY(t && parts && out_r, "null argument");
    ARG_TRY(n_parts >= 1 && n_points >= 2 && n_points <= kMaxTerms, "bad shape");
    const HostField& F = t->f->h;
    std::vector<Fe> ev(n_points, F.zero());
    for (uint32_t g = 0; g < n_parts; ++g)
        for (uint32_t x = 0; x < n_points; ++x) {
            Fe e;
            F.load(parts + ((size_t)g * n_points + x) * F.n, e);
            ARG_TRY(F.is_canonical(e), "partial sum is not a canonical field element");
            ev[x] = F.add(ev[x], e);
        }
    const size_t before = t->bytes.size();
    if (t->offsets.size() == 1) {  // g_1 = (c_1, poly): c_1 = sum over the hypercube = g_1(0) + g_1(1)
        t->c_1 = F.add(ev[0], ev[1]);
        F.serialize(t->c_1, t->bytes);
    }
    evals_to_poly(F, t->kind, ev).serialize(F, t->bytes);
    t->offsets.push_back(t->bytes.size());
    t->chain->absorb(t->bytes.data() + before, t->bytes.size() - before);
    F.store(t->chain->challenge(), out_r);
    return SCB_OK;
}
extern "C" int scb_transcript_c_1(const scb_transcript* t, uint64_t* out_elem) {
    ARG_TRY(t && out_elem, "null argument");
    t->f->h.store(t->c_1, out_elem);
    return SCB_OK;
}
extern "C" int scb_transcript_bytes(const scb_transcript* t, uint8_t* out, size_t cap, size_t* out_len, uint64_t* offsets, uint32_t cap_msgs,
                                    uint32_t* n_msgs) {
    ARG_TRY(t && out_len && n_msgs, "null argument");
    *out_len = t->bytes.size();
    *n_msgs = (uint32_t)t->offsets.size() - 1;
    if (!out) return SCB_OK;  // size query
    ARG_TRY(offsets && cap >= t->bytes.size() && cap_msgs + 1 >= t->offsets.size(), "output buffer too small");
    std::memcpy(out, t->bytes.data(), t->bytes.size());
    std::memcpy(offsets, t->offsets.data(), 8 * t->offsets.size());
    return SCB_OK;
}
