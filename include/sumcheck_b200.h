/*
 * sumcheck_b200.h -- C ABI of libsumcheck_b200.so, the B200-native drop-in for the sum-check
 * prover hot path of montekki/thaler-study.
 *
 * The reference has no FFI; its one extension point is the Rust trait
 *   sum_check_protocol::SumCheckPolynomial<F>      (sum-check-protocol/src/lib.rs:121-156)
 * which Prover/Verifier (:73-117, :227-331), fiat-shamir (fiat-shamir/src/lib.rs:44-98) and the GKR
 * prover (gkr-protocol/src/lib.rs:335,426,446,453) are generic over.  The entry points below are
 * what a Rust `impl SumCheckPolynomial<F> for GpuPoly<F>` binds (see INTEGRATION.md and
 * rust/sumcheck-b200/src/lib.rs); each one cites the reference interface it replaces.
 *
 * Data format across the boundary: ark-ff's in-memory Fp -- n_limbs little-endian uint64_t limbs
 * per element, MONTGOMERY form with R = 2^(64 n_limbs), value < p -- so a Rust `&[F]` crosses as
 * `(const uint64_t*, len)` with no conversion.  Host buffers unless a name says `_device`.
 *
 * Conventions: every function returns SCB_OK (0) or a negative scb_status; the message is kept in a
 * thread-local string (scb_last_error).  Nothing throws or aborts across the boundary.  A handle is
 * not thread-safe; distinct handles may be used from distinct threads.  Calls are synchronous with
 * respect to their host outputs; device work is ordered on the library's current stream
 * (scb_set_stream; default = the legacy default stream, which is also torch's default stream).
 * There is no CPU fallback: without a CUDA device every compute call returns SCB_ECUDA.
 */
#ifndef SUMCHECK_B200_H
#define SUMCHECK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum scb_status {
    SCB_OK = 0,
    SCB_EINVAL = -1,   /* bad argument / dimension mismatch (Rust shim maps it to `None`, :126) */
    SCB_ENOMEM = -2,
    SCB_ECUDA = -3,    /* CUDA runtime error or no device */
    SCB_ENCCL = -4,    /* reserved for the multi-GPU exchange */
    SCB_EVERIFY = -5,  /* sum_check_protocol::Error::ProverClaimMismatch (:26-27) */
    SCB_ENOPOLY = -6,  /* sum_check_protocol::Error::NoPolySet (:29-30) */
    SCB_ETAIL = -7     /* the resident tail kernel lost lock-step with the host (e.g. under a serialising profiler);
                          *rounds_done rounds were completed, the caller continues with one launch per round */
} scb_status;

typedef struct scb_field scb_field;       /* a prime field (MontConfig of the reference, e.g. :349-354) */
typedef struct scb_mle scb_mle;           /* [ARK] DenseMultilinearExtension<F>, table resident in HBM */
typedef struct scb_poly scb_poly;         /* an implementor of SumCheckPolynomial<F> */
typedef struct scb_prover scb_prover;     /* sum_check_protocol::Prover<F,P>   (:73-117) */
typedef struct scb_verifier scb_verifier; /* sum_check_protocol::Verifier<F,P> (:227-331) */

/* ------------------------------------------------------------------ library / device */
const char* scb_last_error(void);
const char* scb_version(void);
int scb_device_count(int* out);
int scb_set_stream(void* cuda_stream);    /* cudaStream_t; NULL = legacy default stream */
int scb_synchronize(void);
/* number of kernels this library launched since load / last reset (bench.py's gpu_launches) */
int scb_launch_count(uint64_t* out, int reset);

/* Tuning / diagnostic switches (thaler_study_b200/csrc/options.hpp lists them with their defaults).  This is the ONLY
 * way to change them: the library never reads the process environment.  Every switch selects between code paths that
 * return the same field elements and transcript bytes.  Process-wide; set them before the calls they affect.
 * SCB_EINVAL for an unknown name.  scb_option_name(i) enumerates the names (NULL past the last). */
int scb_set_option(const char* name, int64_t value);
int scb_get_option(const char* name, int64_t* out);
const char* scb_option_name(uint32_t index);
void scb_reset_options(void);

/* ------------------------------------------------------------------ field (a11) */
/* replaces #[derive(MontConfig)] #[modulus = ..] + Fp64<MontBackend<_,1>> (sum-check-protocol/src/lib.rs:349-354);
 * n_limbs in {1, 4}; derives -p^-1 mod 2^64, R, R^2 */
int scb_field_create(uint32_t n_limbs, const uint64_t* modulus_le, scb_field** out);
void scb_field_free(scb_field* f);
int scb_field_n_limbs(const scb_field* f, uint32_t* out);
int scb_field_modulus_bits(const scb_field* f, uint32_t* out);
/* 0 = small-prime 32-bit path (p < 2^28), 1 = generic 1 limb, 4 = 4 limbs */
int scb_field_policy(const scb_field* f, uint32_t* out);
/* host helpers: canonical integer limbs <-> Montgomery limbs (F::from_bigint / into_bigint) */
int scb_field_to_mont(const scb_field* f, const uint64_t* canonical, uint64_t* mont, size_t count);
int scb_field_from_mont(const scb_field* f, const uint64_t* mont, uint64_t* canonical, size_t count);

/* ------------------------------------------------------------------ dense MLE (a4, a10) */
/* DenseMultilinearExtension::from_evaluations_vec / _slice (matrix-multiplication/src/lib.rs:81,85) */
int scb_mle_from_host(const scb_field* f, uint32_t num_vars, const uint64_t* evals, scb_mle** out);
/* table already in HBM: copy != 0 copies it, copy == 0 borrows the pointer (caller keeps it alive,
 * 32-byte aligned) */
int scb_mle_from_device(const scb_field* f, uint32_t num_vars, const uint64_t* d_evals, int copy, scb_mle** out);
/* synthetic uniform table generated on the device: entry i = stream(seed, start + i) */
int scb_mle_synthetic(const scb_field* f, uint32_t num_vars, uint64_t seed, uint64_t start, scb_mle** out);
int scb_mle_clone(const scb_mle* m, scb_mle** out);          /* #[derive(Clone)]; O(1), shares the table */
void scb_mle_free(scb_mle* m);
int scb_mle_num_vars(const scb_mle* m, uint32_t* out);
int scb_mle_device_ptr(const scb_mle* m, const uint64_t** out);
/* [ARK] fix_variables(partial_point): t[b] = t[2b] + r (t[2b+1] - t[2b]) per coordinate, variable 0 =
 * index LSB (called at matrix-multiplication/src/lib.rs:83,86,104-105) */
int scb_mle_fix_variables(const scb_mle* m, const uint64_t* partial_point, uint32_t n_point, scb_mle** out);
/* [ARK] Polynomial::evaluate(point) = fix_variables(point)[0], LSB-first (matrix-multiplication/src/lib.rs:97-98) */
int scb_mle_evaluate(const scb_mle* m, const uint64_t* point, uint32_t n_point, uint64_t* out_elem);
/* same value with the point in the big-endian order of multilinear-extensions (r[0] <-> index MSB) */
int scb_mle_evaluate_be(const scb_mle* m, const uint64_t* r, uint32_t n_r, uint64_t* out_elem);
/* n_points evaluations of ONE table, LSB-first points stored back to back (points[t * n_point + j]): what the GKR
 * prover's restrict_poly needs -- W~ along the line through b* and c*, k + 1 points (gkr-protocol/src/lib.rs:291-321) --
 * as one pass over the table per 8 points instead of one per point.  out_elems: n_points elements. */
int scb_mle_evaluate_many(const scb_mle* m, const uint64_t* points, uint32_t n_point, uint32_t n_points, uint64_t* out_elems);
/* [ARK] relabel(a, b, k) (matrix-multiplication/src/lib.rs:82) */
int scb_mle_relabel(const scb_mle* m, uint32_t a, uint32_t b, uint32_t k, scb_mle** out);
/* [ARK] to_evaluations(): device -> host copy of the 2^num_vars entries */
int scb_mle_to_evaluations(const scb_mle* m, uint64_t* out, size_t cap_elems);
/* device -> device copy of the table (stream-ordered; used to hand slabs to the multi-GPU exchange) */
int scb_mle_copy_to_device(const scb_mle* m, uint64_t* d_out);

/* ------------------------------------------------------------------ multilinear-extensions (a8, a9) */
/* vsbw_multilinear_from_evaluations(evals, r)  multilinear-extensions/src/lib.rs:6-24 */
int scb_vsbw_multilinear_from_evaluations(const scb_field* f, const uint64_t* evals, size_t n_evals, const uint64_t* r,
                                          uint32_t n_r, uint64_t* out_elem);
/* cti_multilinear_from_evaluations(evals, r)   multilinear-extensions/src/lib.rs:29-48 (same element) */
int scb_cti_multilinear_from_evaluations(const scb_field* f, const uint64_t* evals, size_t n_evals, const uint64_t* r,
                                         uint32_t n_r, uint64_t* out_elem);

/* ------------------------------------------------------------------ SumCheckPolynomial implementors */
typedef enum scb_poly_kind {
    SCB_POLY_PRODUCT = 0,    /* ProductMLE<K>: K dense MLEs over the same variables (new impl) */
    SCB_POLY_MATMUL_G = 1,   /* matrix_multiplication::G {f_a, f_b}   matrix-multiplication/src/lib.rs:12-15 */
    SCB_POLY_TRIANGLE_G = 2, /* triangle_counting::G                  triangle-counting/src/lib.rs:22-27 */
    SCB_POLY_GKR_W = 3       /* gkr_protocol::round_polynomial::W     gkr-protocol/src/round_polynomial.rs:23-28 */
} scb_poly_kind;

int scb_poly_product(const scb_mle* const* tables, uint32_t k, scb_poly** out);
int scb_poly_matmul_g(const scb_mle* f_a, const scb_mle* f_b, scb_poly** out);
/* ProductMLE<K> straight from K caller-owned host tables of 2^num_vars elements each, in ark's in-memory format
 * (SURVEY 8b scb_poly_create; what a Rust shim calls with `evaluations.as_ptr()` of each DenseMultilinearExtension,
 * sum-check-protocol/src/lib.rs:88-97 being the consumer).  The host tables may be reused on return.
 * Small-prime fields (one limb, p < 2^28) and num_vars >= 22: the tables cross PCIe narrowed where the host cores keep
 * up (several threads, pinned staging; three 21-bit entries per 64-bit word when p < 2^21, uint32 otherwise) and as
 * they are where not, narrowed on the device -- the handle then holds the packed uint32 layout the prover's own folded
 * tables use (scb_poly_allow_packed).  SCB_EINVAL if an entry is not a canonical field element (>= p).  Options
 * (scb_set_option): host_pack = 0 (plain copies), host_pack_threads, host_pack_min_vars, host_pack_chunk_log2,
 * host_pack_raw = 0 (no device-side lane; it is also skipped when a table is not in pinned memory), host_pack_wire = 32
 * (uint32 on the wire), host_pack_nt = 1 (streaming stores into the staging buffers), host_pack_prefetch = bytes the pack
 * threads prefetch ahead of their loads (default 4096; 0 = none). */
int scb_poly_product_from_host(const scb_field* f, uint32_t k, uint32_t num_vars, const uint64_t* const* host_tables, scb_poly** out);
/* scb_mle_from_host and the two *_multilinear_from_evaluations calls take the same narrowing upload for such tables
 * when the process is the only rank of its box (option local_ranks = 1, the default; the sharded driver sets it) and
 * widen to 8-byte entries on the device. */
/* the last packed upload: chunks narrowed by the host lane and by the device lane, bytes of all its H2D copies */
int scb_host_pack_stats(uint64_t* packed_chunks, uint64_t* raw_chunks, uint64_t* h2d_bytes);
/* the upload's host-side scheduler run against a memcpy back end on k random tables (no device needed): SCB_OK if
 * every entry arrives, narrowed, where it belongs */
int scb_host_pack_selftest(uint32_t k, uint32_t num_vars, uint32_t chunk_log2, uint32_t workers, uint32_t raw_lane, uint64_t seed);
/* G::new(n, a, b, point)  matrix-multiplication/src/lib.rs:77-92: a, b = row-major n x n matrices
 * (2^(2n) elements each), point = 2n elements */
int scb_poly_matmul_g_new(const scb_field* f, uint32_t n, const uint64_t* a, const uint64_t* b, const uint64_t* point,
                          scb_poly** out);
/* G::new_adj_matrix(num_vars, matrix)  triangle-counting/src/lib.rs:32-51: adj = 2^num_vars bytes (0/1) */
int scb_poly_triangle_g_new(const scb_field* f, uint32_t num_vars, const uint8_t* adj, scb_poly** out);
/* W::new(add_i, mul_i, w_b, w_c)  gkr-protocol/src/round_polynomial.rs:32-44 */
int scb_poly_gkr_w(const scb_mle* add_i, const scb_mle* mul_i, const scb_mle* w_b, const scb_mle* w_c, scb_poly** out);
int scb_poly_clone(const scb_poly* p, scb_poly** out);
void scb_poly_free(scb_poly* p);
int scb_poly_kind_of(const scb_poly* p, uint32_t* out);
int scb_poly_n_tables(const scb_poly* p, uint32_t* out);
int scb_poly_table(const scb_poly* p, uint32_t idx, scb_mle** out); /* clone of table idx */
/* number of points of a round message (degree bound + 1) */
int scb_poly_n_points(const scb_poly* p, uint32_t* out);

/* trait methods, sum-check-protocol/src/lib.rs:121-156 */
int scb_poly_evaluate(const scb_poly* p, const uint64_t* point, uint32_t n_point, uint64_t* out_elem);  /* :126 */
int scb_poly_fix_variables(const scb_poly* p, const uint64_t* partial_point, uint32_t n_point, scb_poly** out); /* :130 */
int scb_poly_num_vars(const scb_poly* p, uint32_t* out);                                                /* :151 */
int scb_poly_to_evaluations(const scb_poly* p, uint64_t* out, size_t cap_elems);                        /* :155 */
/* to_univariate (:148) as (degree, coefficient) terms of the univariate::SparsePolynomial */
int scb_poly_to_univariate(const scb_poly* p, uint64_t* degrees, uint64_t* coeffs, uint32_t cap_terms, uint32_t* n_terms);
/* the device half of to_univariate: sums over the hypercube of the remaining variables at
 * X = 0 .. n_points-1 (matrix-multiplication/src/lib.rs:110-122) */
int scb_poly_round_evals(const scb_poly* p, uint32_t n_points, uint64_t* out_elems);
/* c_1 = to_evaluations().into_iter().sum() without the 2^v-entry host Vec (Prover::new, :89) */
int scb_poly_sum(const scb_poly* p, uint64_t* out_elem);
/* fused `g = g.fix_variables(&[r]); g.to_univariate()` of Prover::round (:105-112): one pass */
int scb_poly_fix_and_round_evals(const scb_poly* p, const uint64_t* r, uint32_t n_points, scb_poly** out,
                                 uint64_t* out_elems);
/* The same with the claim the message must satisfy, g(0) + g(1) = claim (= g_{j-1}(r_{j-1}), which Prover::round's caller
 * holds: sum-check-protocol/src/lib.rs:286-291,316-323 is the verifier's side of it).  4-limb fields then run the leaner
 * kernel of csrc/g4.cuh, which skips the X = 1 point and the highest point (it sums the leading coefficient instead) and
 * rebuilds the K+1 values with exact field arithmetic: identical out_elems, 12 instead of 14 Montgomery products per 4
 * table entries for K = 3.  Other fields: same as scb_poly_fix_and_round_evals.  A wrong claim gives wrong out_elems[1..]. */
int scb_poly_fix_and_round_evals_claim(const scb_poly* p, const uint64_t* r, const uint64_t* claim, uint32_t n_points, scb_poly** out,
                                       uint64_t* out_elems);
/* sharded / device-resident variant: the partial sums stay on the device (n_points elements at
 * d_out) so that the per-round exchange of the multi-GPU prover never touches the host */
int scb_poly_round_evals_device(const scb_poly* p, uint32_t n_points, uint64_t* d_out);
int scb_poly_fix_and_round_evals_device(const scb_poly* p, const uint64_t* r, uint32_t n_points, scb_poly** out,
                                        uint64_t* d_out);

/* Opt-in layout optimisation for one-limb fields with p < 2^28 (all of the reference's moduli): polynomials
 * DERIVED from `p` by the fix_and_round calls may keep their (internal, never exposed) folded tables as packed
 * uint32 -- a quarter less HBM traffic over a proof, same field elements.  Every other entry point converts such a
 * handle back to ark's 8-byte layout on entry, so behaviour is unchanged.  scb_prover_new enables it on its clone. */
int scb_poly_allow_packed(scb_poly* p, int enable);

/* All remaining rounds of a product polynomial (num_vars = m >= 2 -> m-1 rounds of Prover::round, :105-112) in ONE
 * resident kernel: per round the callback gets the n_points sums and returns the next challenge through a mailbox
 * in mapped pinned memory -- no launches or stream synchronisation between rounds (latency-bound tail).
 * The callback returns SCB_OK or an error, which aborts the kernel.  `p` itself is not modified. */
typedef int (*scb_round_cb)(void* user, uint32_t round, const uint64_t* evals, uint64_t* next_challenge_out);
int scb_poly_tail_rounds(const scb_poly* p, const uint64_t* r_first, uint32_t n_points, scb_round_cb cb, void* user,
                         uint32_t* rounds_done);
/* Same, for at most `max_rounds` rounds (0 = all m-1): small tables run in a single-CTA kernel, large ones in a
 * grid-wide cooperative kernel whose CTAs meet at a ticket/flag barrier between rounds.  With a current peer group
 * (scb_peers_set_current) the grid-wide kernel also exchanges the round sums with the peer GPUs every round.
 * `*out_folded` (optional) receives the polynomial after the rounds that ran -- what max_rounds calls of
 * scb_poly_fix_and_round_evals would have returned last. */
int scb_poly_resident_rounds(const scb_poly* p, const uint64_t* r_first, uint32_t n_points, uint32_t max_rounds, scb_round_cb cb,
                             void* user, uint32_t* rounds_done, scb_poly** out_folded);

/* Two rounds per pass over the tables (small-prime fields, product polynomials; csrc/pairs.cuh).
 * H[a][b] = sum over x'' of prod_k f_k(a, b, x'') for a, b in {0..K} is a bivariate polynomial whose slices are two
 * consecutive round messages: Prover::round j sends g_j(X) = H(X,0) + H(X,1) and round j+1 sends g_{j+1}(Y) = H(r_j, Y)
 * (sum-check-protocol/src/lib.rs:105-112 applied twice), so one pass yields both and the next pass folds two
 * variables at once.  grid_evals: the (K+1)^2 sums, a-major, of the polynomial as it is.  pair_pass: fold the two
 * lowest variables by (ra, rb), then the grid of the folded polynomial (needs >= 4 variables).  resident_pairs: every
 * pair pass of a proof (at most max_passes, 0 = all) in one cooperative kernel; the callback gets the grid ((K+1)^2
 * values) or, when a single variable is left, the line ((K+1) values) and returns the next challenge pair;
 * *out_folded (optional) receives the polynomial the last pass left behind.  With a current peer group the grid
 * kernels add the peer GPUs' sums (sharded prover). */
typedef int (*scb_pair_cb)(void* user, uint32_t pass, uint32_t n_vals, const uint64_t* vals, uint64_t* next_pair_out);
int scb_poly_grid_evals(const scb_poly* p, uint64_t* out_elems);
int scb_poly_pair_pass(const scb_poly* p, const uint64_t* ra, const uint64_t* rb, scb_poly** out, uint64_t* out_elems);
int scb_poly_resident_pairs(const scb_poly* p, const uint64_t* ra, const uint64_t* rb, uint32_t max_passes, scb_pair_cb cb,
                            void* user, uint32_t* passes_done, scb_poly** out_folded);

/* Measurement hooks for the resident kernels (bench.py's roofline): launches and summed CUDA-event kernel time since
 * the last reset, and for the last grid-wide launch the per-round device time up to "sums posted" and the host
 * turn-around (posted -> next challenge seen by the kernel), microseconds, from %globaltimer stamps. */
int scb_resident_stats(uint64_t* launches, double* total_kernel_ms, uint32_t* last_rounds, double* last_work_us,
                       double* last_turn_us, uint32_t cap);
/* launches and summed CUDA-event kernel time of the stand-alone pair passes (scb_poly_pair_pass) since the last reset */
int scb_pair_pass_stats(uint64_t* launches, double* total_kernel_ms);
/* the same for the grid passes (scb_poly_grid_evals, Prover::new), and how many grid / pair passes wrote / read the
 * 21-bit triples of option pair_w21 (K = 3 tables over a field of at most 21 bits; csrc/pairs.cuh) */
int scb_grid_pass_stats(uint64_t* launches, double* total_kernel_ms, uint64_t* w21_grid_launches, uint64_t* w21_pair_launches);
void scb_resident_stats_reset(void);

/* ------------------------------------------------------------------ round-message algebra (host) */
/* (d+1) sums at X = 0..d  ->  the SparsePolynomial the reference would send, per implementor:
 * MATMUL_G: interpolate_quadratic_poly (matrix-multiplication/src/lib.rs:17-60, explicit zero terms kept);
 * others: unique interpolant, Dense -> Sparse (triangle-counting/src/lib.rs:128-131) */
int scb_evals_to_univariate(const scb_field* f, uint32_t kind, const uint64_t* evals, uint32_t n_points, uint64_t* degrees,
                            uint64_t* coeffs, uint32_t cap_terms, uint32_t* n_terms);
/* SparsePolynomial::serialize_uncompressed (fiat-shamir/src/lib.rs:58) */
int scb_unipoly_serialize(const scb_field* f, const uint64_t* degrees, const uint64_t* coeffs, uint32_t n_terms, uint8_t* out,
                          size_t cap, size_t* out_len);
int scb_unipoly_evaluate(const scb_field* f, const uint64_t* degrees, const uint64_t* coeffs, uint32_t n_terms,
                         const uint64_t* x, uint64_t* out_elem);
/* DefaultFieldHasher<Sha256>::new(&[]).hash_to_field::<1>(msg)[0] (fiat-shamir/src/lib.rs:78,88) */
int scb_hash_to_field(const scb_field* f, const uint8_t* msg, size_t len, uint64_t* out_elem);

/* ------------------------------------------------------------------ Prover / Verifier */
int scb_prover_new(const scb_poly* g, scb_prover** out);                   /* Prover::new  :88-97 */
void scb_prover_free(scb_prover* p);
int scb_prover_c_1(const scb_prover* p, uint64_t* out_elem);               /* :100-102 */
int scb_prover_num_vars(const scb_prover* p, uint32_t* out);               /* :114-116 */
/* Prover::round(r_prev, j) :105-112 (r_prev ignored when j == 0) */
int scb_prover_round(scb_prover* p, const uint64_t* r_prev, uint32_t j, uint64_t* degrees, uint64_t* coeffs,
                     uint32_t cap_terms, uint32_t* n_terms);

/* Verifier::new(n, g) :261-269; g may be NULL (no oracle access) */
int scb_verifier_new(const scb_field* f, uint32_t n, const scb_poly* g, scb_verifier** out);
void scb_verifier_free(scb_verifier* v);
int scb_verifier_set_c_1(scb_verifier* v, const uint64_t* c_1);            /* :271-273 */
/* Verifier::round(g_j, rng) :278-330 with rng.draw() == r_j.  *final_round = 1 and *accepted set on the
 * last round (VerifierRoundResult::FinalRound), else *final_round = 0 (JthRound(r_j)).
 * SCB_EVERIFY = ProverClaimMismatch, SCB_ENOPOLY = NoPolySet.
 * Option strict_verifier (default 1): the final round ALSO checks g_n(0) + g_n(1) == g_{n-1}(r_{n-1}), and a
 * one-variable verifier evaluates its oracle in its only round (FinalRound) -- the reference's :298-310 and
 * :284-297 skip both, which lets a false c_1 through.  Honest transcripts are unaffected;
 * strict_verifier = 0 reproduces the reference literally. */
int scb_verifier_round(scb_verifier* v, const uint64_t* degrees, const uint64_t* coeffs, uint32_t n_terms, const uint64_t* r_j,
                       int* final_round, int* accepted);

/* ------------------------------------------------------------------ fiat-shamir */
/* generate_transcript::<F, Prover<F,P>, DefaultFieldHasher<Sha256>>(prover)  fiat-shamir/src/lib.rs:75-98.
 * Consumes the prover's rounds.  out = g_1 || g_2 || ...; offsets[i]..offsets[i+1] delimits message i
 * (offsets has num_vars + 1 entries). */
int scb_fs_generate_transcript(scb_prover* p, uint8_t* out, size_t cap, size_t* out_len, uint64_t* offsets);
/* verify_transcript(transcript, verifier)  fiat-shamir/src/lib.rs:123-143.  `transcript` is untrusted input:
 * offsets[0] == 0 <= offsets[1] <= ... <= offsets[n_msgs] <= transcript_len is enforced (SCB_EINVAL otherwise).
 * With option strict_verifier (default 1) *accepted = 1 only if the verifier reached and passed its final round
 * (n_msgs == n): the reference loops over whatever prefix it is given and never reaches the oracle check for a
 * truncated transcript. */
int scb_fs_verify_transcript(scb_verifier* v, const uint8_t* transcript, size_t transcript_len, const uint64_t* offsets,
                             uint32_t n_msgs, int* accepted);

/* The same hash chain as an explicit state machine, for provers whose round sums arrive in pieces (the sharded
 * multi-GPU prover: one row of (d+1) partial sums per rank).  absorb_round adds the n_parts rows mod p, turns the
 * sums into the implementor's SparsePolynomial, appends its serialization (the first message is prefixed with
 * c_1 = g_1(0) + g_1(1), fiat-shamir/src/lib.rs:48-50) and returns r = hash_to_field(all bytes so far) (:88). */
typedef struct scb_transcript scb_transcript;
int scb_transcript_new(const scb_field* f, uint32_t kind, scb_transcript** out);
void scb_transcript_free(scb_transcript* t);
int scb_transcript_absorb_round(scb_transcript* t, const uint64_t* parts, uint32_t n_parts, uint32_t n_points, uint64_t* out_r);
int scb_transcript_c_1(const scb_transcript* t, uint64_t* out_elem);
/* out == NULL: size query (out_len, n_msgs) */
int scb_transcript_bytes(const scb_transcript* t, uint8_t* out, size_t cap, size_t* out_len, uint64_t* offsets, uint32_t cap_msgs,
                         uint32_t* n_msgs);

/* ------------------------------------------------------------------ GKR with gate-list wiring
 * gkr_protocol::Circuit (gkr-protocol/src/circuit.rs:70-124) and the layer prover of gkr_protocol::Prover
 * (gkr-protocol/src/lib.rs:346-456) without the reference's dense 2^(k_i + 2 k_{i+1})-entry wiring tables: each layer is
 * two k-round sum-checks of the form P*Q + S over 2^k-entry tables built from the gate list (thaler_study_b200/csrc/gkr.cuh).
 * Messages are the same polynomials as the reference's W::to_univariate (round_polynomial.rs:78-90). */
typedef struct scb_circuit scb_circuit;
typedef struct scb_gkr_prover scb_gkr_prover;
/* Circuit::new(layers, num_inputs): layer 0 = outputs; gates concatenated layer after layer; type 0 = Add, 1 = Mul;
 * in0/in1 index the next layer (the inputs for the last layer).  All sizes must be powers of two. */
int scb_circuit_create(const scb_field* f, uint32_t n_layers, const uint32_t* layer_sizes, const uint8_t* types, const uint32_t* in0,
                       const uint32_t* in1, uint32_t num_inputs, scb_circuit** out);
void scb_circuit_free(scb_circuit* c);
int scb_circuit_num_vars_at(const scb_circuit* c, uint32_t layer, uint32_t* out);              /* circuit.rs:86-96 */
/* add~_i(r_i, b, c) and mul~_i(r_i, b, c) (circuit.rs:152-212 evaluated at (b, c)) in O(#gates) */
int scb_circuit_wiring_eval(const scb_circuit* c, uint32_t layer, const uint64_t* r_i, const uint64_t* b, const uint64_t* cpt,
                            uint64_t* add_out, uint64_t* mul_out);
int scb_gkr_prover_new(const scb_circuit* c, const uint64_t* input, size_t n_input, scb_gkr_prover** out); /* lib.rs:346-357 */
void scb_gkr_prover_free(scb_gkr_prover* p);
/* CircuitEvaluation::layers[layer] as a dense MLE handle (layer 0 = outputs: start_protocol, lib.rs:363-367) */
int scb_gkr_prover_layer(const scb_gkr_prover* p, uint32_t layer, scb_mle** out);
/* start_round(i, r_i) lib.rs:373-436 -> StartSumCheck { c_1, num_vars = 2 k_{i+1} } */
int scb_gkr_prover_start_round(scb_gkr_prover* p, uint32_t i, const uint64_t* r_i, uint64_t* c_1_out, uint32_t* num_vars_out);
/* round_msg(j) lib.rs:439-456 as the sums at X = 0,1,2 (-> scb_evals_to_univariate, kind SCB_POLY_GKR_W); r_prev = r_{j-1} */
int scb_gkr_prover_round_evals(scb_gkr_prover* p, uint32_t j, const uint64_t* r_prev, uint64_t* out_evals);
/* q = restrict_poly(b*, c*, W~_{i+1}) lib.rs:291-321 as its k+1 values at t = 0..k; r_final = the final random point */
int scb_gkr_prover_restrict_evals(scb_gkr_prover* p, const uint64_t* r_final, uint64_t* out_evals, uint32_t cap, uint32_t* n_out);
/* One whole layer proof with the verifier's 2k public-coin challenges rs[0..2k) handed over up front: equivalent to
 * start_round + round_evals(j, rs[j-1]) for j = 0..2k-1 + restrict_evals(rs[2k-1]); each phase is one cooperative
 * launch with the challenges in device memory (one host wait per layer instead of one per round).  evals_out: 2k x 3 elements (sums at X = 0,1,2 per round); q_out: k+1. */
int scb_gkr_prover_prove_layer(scb_gkr_prover* p, uint32_t i, const uint64_t* r_i, const uint64_t* rs, uint32_t n_rs, uint64_t* c_1_out,
                               uint64_t* evals_out, uint64_t* q_out, uint32_t cap_q, uint32_t* num_vars_out);

/* ------------------------------------------------------------------ multi-GPU: peer windows over NVLink
 * One process per GPU.  Each rank owns a small window in its HBM (cudaMalloc) that every peer maps through CUDA IPC;
 * the round kernels' finishing thread posts the rank's (d+1) partial sums into all peers' windows with P2P stores,
 * waits for the peers' posts and adds the rows mod p, so "per-round all-gather + modular sum" is part of the round
 * kernel (no collective launch).  At consolidation the slabs are published the same way.  The handles (64 bytes per
 * rank) are exchanged by the caller with any transport (torch.distributed in thaler_study_b200/distributed.py). */
typedef struct scb_peers scb_peers;
int scb_peers_create(uint32_t rank, uint32_t world, size_t gather_bytes, scb_peers** out, uint8_t* handle_out /* 64 B */);
/* capacity in bytes of one gather area of the window (what scb_peers_create was given, rounded up): the sharded prover
 * picks its consolidation point so that all ranks' slabs of all tables fit */
int scb_peers_gather_capacity(const scb_peers* p, size_t* out_bytes);
int scb_peers_connect(scb_peers* p, const uint8_t* all_handles /* world * 64 B, rank order */);
void scb_peers_free(scb_peers* p);
/* make `p` (or none) the exchange group used by this thread's subsequent scb_poly_round_evals /
 * scb_poly_fix_and_round_evals calls: their results are then the sums over ALL ranks' slabs */
int scb_peers_set_current(scb_peers* p);
/* all-gather the slabs of `slab` (rank-order concatenation of every table) into a replicated polynomial */
int scb_peers_gather_poly(scb_peers* p, const scb_poly* slab, scb_poly** out);
/* Prover::new for a polynomial sharded by its top log2(world) variables; `slab` is this rank's part.  The prover
 * consolidates (scb_peers_gather_poly) once a slab has at most consolidate_at variables (0: the library's default,
 * option consolidate_auto); scb_prover_round and
 * scb_fs_generate_transcript then work as for a single GPU and return identical bytes on every rank. */
int scb_prover_new_sharded(const scb_poly* slab, scb_peers* peers, uint32_t world, uint32_t consolidate_at, scb_prover** out);
/* vsbw_multilinear_from_evaluations / [ARK] evaluate (multilinear-extensions/src/lib.rs:6-24) of a table sharded by its top
 * log2(world) variables: `slab` = this rank's entries [rank 2^lv, (rank+1) 2^lv), point = lv + log2(world) coordinates
 * (big_endian != 0: r[0] <-> index MSB as in multilinear-extensions; 0: LSB-first as in ark).  One launch per rank; the
 * E-byte exchange and the modular sum run in the kernel's finishing thread over the peer windows.  Same element on
 * every rank.  The slab needs at least 8 variables (10 for 4-limb fields). */
int scb_mle_evaluate_sharded(const scb_mle* slab, scb_peers* peers, const uint64_t* point, uint32_t n_point, int big_endian,
                             uint64_t* out_elem);

#ifdef __cplusplus
}
#endif
#endif /* SUMCHECK_B200_H */
