/*
 * oracle.c -- plain-C CPU ORACLE for the sum-check hot path of montekki/thaler-study.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under thaler_study_b200/ links, loads or
 * calls this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, and only as the checker / the CPU
 * baseline.  The product path has no CPU fallback.
 *
 * It restates, in the reference's own naive structure (separate fold pass and
 * message pass per round, table copies included), the functions of SURVEY.md
 * section 8(a).  Elements are ark-ff's in-memory representation: N little-endian
 * u64 limbs in Montgomery form with R = 2^(64N), value < p ([ARK]
 * Fp<MontBackend<_,N>,N>; the ark-ff 0.6 source is not vendored in the
 * reference, Cargo.toml:20-25, so its published behaviour is restated).
 * Paths below are relative to /root/reference.
 *
 * Pinning: cross-checked against oracle/pyoracle.py (Python big-int restatement,
 * itself pinned to every KAT the reference's tests hold) in
 * tests/test_oracle_c.py.  PARITY UNPINNED for fields wider than one limb and
 * tables above 2^10 entries (the reference has no vector there).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_LIMBS 4
typedef unsigned __int128 u128;

typedef struct {
    uint32_t n;                    /* limbs */
    uint64_t p[ORC_MAX_LIMBS];     /* modulus */
    uint64_t inv;                  /* -p^{-1} mod 2^64 */
    uint64_t one[ORC_MAX_LIMBS];   /* R mod p */
    uint64_t r2[ORC_MAX_LIMBS];    /* R^2 mod p */
} orc_field;

/* ------------------------------------------------------------------ limbs */
static int ge(const uint64_t *a, const uint64_t *b, uint32_t n) {
    for (int i = (int)n - 1; i >= 0; --i) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static uint64_t add_n(uint64_t *o, const uint64_t *a, const uint64_t *b, uint32_t n) {
    u128 c = 0;
    for (uint32_t i = 0; i < n; ++i) { c += (u128)a[i] + b[i]; o[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static uint64_t sub_n(uint64_t *o, const uint64_t *a, const uint64_t *b, uint32_t n) {
    uint64_t br = 0;
    for (uint32_t i = 0; i < n; ++i) {
        u128 d = (u128)a[i] - b[i] - br;
        o[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1;
    }
    return br;
}

/* [ARK] Fp add: a + b, subtract p if >= p */
void orc_add(const orc_field *F, uint64_t *o, const uint64_t *a, const uint64_t *b) {
    uint64_t t[ORC_MAX_LIMBS];
    uint64_t c = add_n(t, a, b, F->n);
    if (c || ge(t, F->p, F->n)) sub_n(t, t, F->p, F->n);
    memcpy(o, t, 8 * F->n);
}
/* [ARK] Fp sub: if b > a add p first */
void orc_sub(const orc_field *F, uint64_t *o, const uint64_t *a, const uint64_t *b) {
    uint64_t t[ORC_MAX_LIMBS];
    if (sub_n(t, a, b, F->n)) add_n(t, t, F->p, F->n);
    memcpy(o, t, 8 * F->n);
}
/* [ARK] MontBackend::mul_assign: CIOS Montgomery product a*b*R^{-1} mod p */
void orc_mul(const orc_field *F, uint64_t *o, const uint64_t *a, const uint64_t *b) {
    const uint32_t n = F->n;
    uint64_t t[ORC_MAX_LIMBS + 2];
    memset(t, 0, sizeof t);
    for (uint32_t i = 0; i < n; ++i) {
        u128 c = 0;
        for (uint32_t j = 0; j < n; ++j) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c; c >>= 64;
        }
        c += t[n]; t[n] = (uint64_t)c; t[n + 1] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = (u128)m * F->p[0] + t[0]; c >>= 64;
        for (uint32_t j = 1; j < n; ++j) {
            c += (u128)m * F->p[j] + t[j];
            t[j - 1] = (uint64_t)c; c >>= 64;
        }
        c += t[n]; t[n - 1] = (uint64_t)c; c >>= 64;
        t[n] = t[n + 1] + (uint64_t)c;
    }
    if (t[n] || ge(t, F->p, n)) sub_n(t, t, F->p, n);
    memcpy(o, t, 8 * n);
}

int orc_field_init(orc_field *F, uint32_t n, const uint64_t *modulus) {
    if (n < 1 || n > ORC_MAX_LIMBS || !(modulus[0] & 1)) return -1;
    memset(F, 0, sizeof *F);
    F->n = n;
    memcpy(F->p, modulus, 8 * n);
    uint64_t inv = 1;                         /* Newton: inv = p^{-1} mod 2^64 */
    for (int i = 0; i < 6; ++i) inv *= 2 - modulus[0] * inv;
    F->inv = (uint64_t)0 - inv;
    /* R mod p and R^2 mod p by repeated doubling of 1 (64n and 128n times) */
    uint64_t x[ORC_MAX_LIMBS] = {1, 0, 0, 0};
    if (n == 1 && modulus[0] == 1) return -1;
    if (ge(x, F->p, n)) sub_n(x, x, F->p, n);
    for (uint32_t i = 0; i < 128 * n; ++i) {
        uint64_t c = add_n(x, x, x, n);
        if (c || ge(x, F->p, n)) sub_n(x, x, F->p, n);
        if (i + 1 == 64 * n) memcpy(F->one, x, 8 * n);
    }
    memcpy(F->r2, x, 8 * n);
    return 0;
}

/* canonical <-> Montgomery (host helpers for the tests) */
void orc_to_mont(const orc_field *F, uint64_t *o, const uint64_t *a, size_t cnt) {
    for (size_t i = 0; i < cnt; ++i) orc_mul(F, o + i * F->n, a + i * F->n, F->r2);
}
void orc_from_mont(const orc_field *F, uint64_t *o, const uint64_t *a, size_t cnt) {
    uint64_t one[ORC_MAX_LIMBS] = {1, 0, 0, 0};
    for (size_t i = 0; i < cnt; ++i) orc_mul(F, o + i * F->n, a + i * F->n, one);
}

/* ------------------------------------------------- synthetic tables (bench) */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* same stream as pyoracle.synth_element and the CUDA generator: limbs ARE the
 * Montgomery-form representation */
void orc_synth_fill(const orc_field *F, uint64_t seed, uint64_t start, size_t cnt, uint64_t *out) {
    const uint32_t n = F->n;
    int bits = 0;
    for (int i = (int)n - 1; i >= 0 && !bits; --i)
        if (F->p[i]) bits = 64 * i + 64 - __builtin_clzll(F->p[i]);
    const int topbits = bits - 64 * ((int)n - 1);
    const uint64_t topmask = topbits >= 64 ? ~0ull : ((1ull << topbits) - 1);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < cnt; ++i) {
        uint64_t v[ORC_MAX_LIMBS];
        uint64_t base = ((seed << 40) + start + i) * n;
        for (uint32_t l = 0; l < n; ++l) v[l] = splitmix64(base + l);
        v[n - 1] &= topmask;
        if (ge(v, F->p, n)) sub_n(v, v, F->p, n);
        memcpy(out + i * n, v, 8 * n);
    }
}

/* ------------------------------------------------------- a4: fix_variables */
/* [ARK] DenseMultilinearExtension::fix_variables(&[r]) -- copy the table, then
 * t[b] = t[2b] + r*(t[2b+1]-t[2b]) for b < N/2, then copy out the low half
 * (from_evaluations_slice).  The copies are part of the reference's cost and
 * are kept.  `threads` > 1 splits the b-range with OpenMP (pairs are
 * independent once a separate output buffer is used). */
void orc_fix_variable(const orc_field *F, const uint64_t *tab, size_t len, const uint64_t *r,
                      uint64_t *out, int threads) {
    const uint32_t n = F->n;
    const size_t half = len / 2;
    if (threads <= 1) {
        uint64_t *poly = (uint64_t *)malloc(len * n * 8);   /* self.evaluations.to_vec() */
        memcpy(poly, tab, len * n * 8);
        for (size_t b = 0; b < half; ++b) {
            uint64_t d[ORC_MAX_LIMBS], m[ORC_MAX_LIMBS];
            const uint64_t *left = poly + 2 * b * n, *right = poly + (2 * b + 1) * n;
            orc_sub(F, d, right, left);
            orc_mul(F, m, r, d);
            orc_add(F, poly + b * n, left, m);
        }
        memcpy(out, poly, half * n * 8);                     /* from_evaluations_slice */
        free(poly);
    } else {
#pragma omp parallel for schedule(static) num_threads(threads)
        for (size_t b = 0; b < half; ++b) {
            uint64_t d[ORC_MAX_LIMBS], m[ORC_MAX_LIMBS];
            const uint64_t *left = tab + 2 * b * n, *right = tab + (2 * b + 1) * n;
            orc_sub(F, d, right, left);
            orc_mul(F, m, r, d);
            orc_add(F, out + b * n, left, m);
        }
    }
}

/* a10: [ARK] evaluate(point) = fix_variables(point)[0], LSB-first */
void orc_mle_evaluate_le(const orc_field *F, const uint64_t *tab, uint32_t v, const uint64_t *point,
                         uint64_t *out) {
    const uint32_t n = F->n;
    size_t len = (size_t)1 << v;
    uint64_t *cur = (uint64_t *)malloc(len * n * 8), *nxt = (uint64_t *)malloc((len / 2 + 1) * n * 8);
    memcpy(cur, tab, len * n * 8);
    for (uint32_t i = 0; i < v; ++i) {
        orc_fix_variable(F, cur, len, point + i * n, nxt, 1);
        len /= 2;
        memcpy(cur, nxt, len * n * 8);
    }
    memcpy(out, cur, n * 8);
    free(cur); free(nxt);
}

/* ------------------------------------------- a2: Prover::new, c_1 = sum(evals) */
/* sum-check-protocol/src/lib.rs:89 with ProductMLE / G::to_evaluations
 * (matrix-multiplication/src/lib.rs:137-146): elementwise product then sum */
void orc_product_sum(const orc_field *F, uint32_t K, const uint64_t *const *tabs, size_t len,
                     uint64_t *out, int threads) {
    const uint32_t n = F->n;
    int T = threads > 1 ? threads : 1;
    uint64_t *part = (uint64_t *)calloc((size_t)T * n, 8);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        uint64_t acc[ORC_MAX_LIMBS] = {0};
#pragma omp for schedule(static)
        for (size_t i = 0; i < len; ++i) {
            uint64_t v[ORC_MAX_LIMBS];
            memcpy(v, tabs[0] + i * n, 8 * n);
            for (uint32_t k = 1; k < K; ++k) orc_mul(F, v, v, tabs[k] + i * n);
            orc_add(F, acc, acc, v);
        }
        memcpy(part + (size_t)tid * n, acc, 8 * n);
    }
    uint64_t acc[ORC_MAX_LIMBS] = {0};
    for (int t = 0; t < T; ++t) orc_add(F, acc, acc, part + (size_t)t * n);
    memcpy(out, acc, 8 * n);
    free(part);
}

/* ------------------------------------------------- a5: to_univariate sums */
/* Generalisation of matrix-multiplication/src/lib.rs:110-122 to K tables and
 * X = 0..npts-1: one pass over adjacent pairs; the value at X is
 * lo + X*(hi-lo), obtained for X >= 2 by repeated addition of (hi-lo)
 * (for K=2, X=2 this equals (two*a[i]-a[i-1])*(two*b[i]-b[i-1]) exactly). */
void orc_product_round_evals(const orc_field *F, uint32_t K, const uint64_t *const *tabs,
                             size_t len, uint32_t npts, uint64_t *out, int threads) {
    const uint32_t n = F->n;
    int T = threads > 1 ? threads : 1;
    if (npts > 8) npts = 8;
    uint64_t *part = (uint64_t *)calloc((size_t)T * npts * n, 8);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        uint64_t acc[8][ORC_MAX_LIMBS];
        memset(acc, 0, sizeof acc);
#pragma omp for schedule(static)
        for (size_t b = 0; b < len / 2; ++b) {
            uint64_t prod[8][ORC_MAX_LIMBS];
            for (uint32_t k = 0; k < K; ++k) {
                const uint64_t *lo = tabs[k] + 2 * b * n, *hi = tabs[k] + (2 * b + 1) * n;
                uint64_t d[ORC_MAX_LIMBS], v[ORC_MAX_LIMBS];
                orc_sub(F, d, hi, lo);
                memcpy(v, lo, 8 * n);
                for (uint32_t x = 0; x < npts; ++x) {
                    if (k == 0) memcpy(prod[x], v, 8 * n);
                    else orc_mul(F, prod[x], prod[x], v);
                    orc_add(F, v, v, d);
                }
            }
            for (uint32_t x = 0; x < npts; ++x) orc_add(F, acc[x], acc[x], prod[x]);
        }
        for (uint32_t x = 0; x < npts; ++x) memcpy(part + ((size_t)tid * npts + x) * n, acc[x], 8 * n);
    }
    for (uint32_t x = 0; x < npts; ++x) {
        uint64_t acc[ORC_MAX_LIMBS] = {0};
        for (int t = 0; t < T; ++t) orc_add(F, acc, acc, part + ((size_t)t * npts + x) * n);
        memcpy(out + x * n, acc, 8 * n);
    }
    free(part);
}

/* ------------------------------- a3: the whole prover loop (CPU baseline) */
/* Prover::new + v x Prover::round (sum-check-protocol/src/lib.rs:88-112) for a
 * ProductMLE<K>, challenges supplied by the caller (as the reference's bench does
 * with Fp5::rand, matrix-multiplication/benches/mm_benchmark.rs:88-96).
 * Tables are consumed (folded in place into fresh buffers each round, like
 * `self.g = self.g.fix_variables(..)`).  Writes c_1 and npts sums per round. */
void orc_product_prove(const orc_field *F, uint32_t K, uint32_t v, uint64_t **tabs,
                       const uint64_t *challenges /* v-1 elements */, uint32_t npts,
                       uint64_t *c1_out, uint64_t *round_evals_out /* v*npts */, int threads) {
    const uint32_t n = F->n;
    size_t len = (size_t)1 << v;
    uint64_t *cur[8];
    for (uint32_t k = 0; k < K; ++k) cur[k] = tabs[k];
    if (c1_out) orc_product_sum(F, K, (const uint64_t *const *)cur, len, c1_out, threads);
    for (uint32_t j = 0; j < v; ++j) {
        if (j != 0) {
            for (uint32_t k = 0; k < K; ++k) {
                uint64_t *nxt = (uint64_t *)malloc((len / 2) * n * 8);
                orc_fix_variable(F, cur[k], len, challenges + (size_t)(j - 1) * n, nxt, threads);
                if (cur[k] != tabs[k]) free(cur[k]);
                cur[k] = nxt;
            }
            len /= 2;
        }
        orc_product_round_evals(F, K, (const uint64_t *const *)cur, len, npts,
                                round_evals_out + (size_t)j * npts * n, threads);
    }
    for (uint32_t k = 0; k < K; ++k)
        if (cur[k] != tabs[k]) free(cur[k]);
}

/* --------------------------------------------------------- a8 / a9: MLE eval */
/* multilinear-extensions/src/lib.rs:6-24: chi table by doubling (r[0] -> MSB),
 * then dot with evals */
void orc_mle_vsbw(const orc_field *F, const uint64_t *evals, uint32_t v, const uint64_t *r,
                  uint64_t *out) {
    const uint32_t n = F->n;
    size_t len = 1;
    uint64_t *table = (uint64_t *)malloc(n * 8);
    memcpy(table, F->one, n * 8);
    for (uint32_t j = 0; j < v; ++j) {
        uint64_t *nt = (uint64_t *)malloc(2 * len * n * 8);
        uint64_t omr[ORC_MAX_LIMBS];
        orc_sub(F, omr, F->one, r + j * n);
        for (size_t i = 0; i < len; ++i) {
            orc_mul(F, nt + (2 * i) * n, table + i * n, omr);
            orc_mul(F, nt + (2 * i + 1) * n, table + i * n, r + j * n);
        }
        free(table); table = nt; len *= 2;
    }
    uint64_t acc[ORC_MAX_LIMBS] = {0};
    for (size_t i = 0; i < len; ++i) {
        uint64_t m[ORC_MAX_LIMBS];
        orc_mul(F, m, table + i * n, evals + i * n);
        orc_add(F, acc, acc, m);
    }
    memcpy(out, acc, n * 8);
    free(table);
}
/* multilinear-extensions/src/lib.rs:29-60: streaming Lagrange basis, big-endian bits */
void orc_mle_cti(const orc_field *F, const uint64_t *evals, uint32_t v, const uint64_t *r,
                 uint64_t *out) {
    const uint32_t n = F->n;
    uint64_t acc[ORC_MAX_LIMBS] = {0};
    uint64_t zero[ORC_MAX_LIMBS] = {0};
    for (size_t i = 0; i < ((size_t)1 << v); ++i) {
        uint64_t basis[ORC_MAX_LIMBS];
        memcpy(basis, F->one, 8 * n);
        for (uint32_t j = 0; j < v; ++j) {
            /* x_i*w_i + (1-x_i)*(1-w_i), w_i = bit (v-1-j) of i */
            const uint64_t *w = ((i >> (v - 1 - j)) & 1) ? F->one : zero;
            uint64_t a[ORC_MAX_LIMBS], b[ORC_MAX_LIMBS], c[ORC_MAX_LIMBS];
            orc_mul(F, a, r + j * n, w);
            orc_sub(F, b, F->one, r + j * n);
            orc_sub(F, c, F->one, w);
            orc_mul(F, b, b, c);
            orc_add(F, a, a, b);
            orc_mul(F, basis, basis, a);
        }
        orc_mul(F, basis, evals + i * n, basis);
        orc_add(F, acc, acc, basis);
    }
    memcpy(out, acc, n * 8);
}

/* ------------------------------------------------- a6: triangle-counting G */
/* State of triangle-counting/src/lib.rs:22-27 after some variables were fixed:
 * f1 over (x:xn, y:yn), f2 over (y:yn, z:zn), f3 over (x:xn, z:zn), index
 * (hi << lo_bits) | lo (:170-172). */
/* to_evaluations().sum() (:138-165 + sum-check-protocol/src/lib.rs:89) */
void orc_triangle_sum(const orc_field *F, const uint64_t *f1, const uint64_t *f2, const uint64_t *f3,
                      uint32_t xn, uint32_t yn, uint32_t zn, uint64_t *out) {
    const uint32_t n = F->n;
    uint64_t acc[ORC_MAX_LIMBS] = {0};
    for (size_t x = 0; x < ((size_t)1 << xn); ++x)
        for (size_t y = 0; y < ((size_t)1 << yn); ++y)
            for (size_t z = 0; z < ((size_t)1 << zn); ++z) {
                uint64_t v[ORC_MAX_LIMBS];
                orc_mul(F, v, f1 + ((y << xn) | x) * n, f2 + ((z << yn) | y) * n);
                orc_mul(F, v, v, f3 + ((z << xn) | x) * n);
                orc_add(F, acc, acc, v);
            }
    memcpy(out, acc, 8 * n);
}
/* one evaluation of to_univariate (:120-126): fix_variables(&[e]) then
 * to_evaluations().sum().  The fold schedule follows :89-118: variable 0 is an
 * x variable while xn > 0 (folds f1, f3), then a y variable (f1, f2), then z (f2, f3). */
void orc_triangle_round_eval_at(const orc_field *F, const uint64_t *f1, const uint64_t *f2,
                                const uint64_t *f3, uint32_t xn, uint32_t yn, uint32_t zn,
                                const uint64_t *e, uint64_t *out) {
    const uint32_t n = F->n;
    size_t l1 = (size_t)1 << (xn + yn), l2 = (size_t)1 << (yn + zn), l3 = (size_t)1 << (xn + zn);
    uint64_t *g1 = (uint64_t *)malloc(l1 * n * 8), *g2 = (uint64_t *)malloc(l2 * n * 8),
             *g3 = (uint64_t *)malloc(l3 * n * 8);
    memcpy(g1, f1, l1 * n * 8); memcpy(g2, f2, l2 * n * 8); memcpy(g3, f3, l3 * n * 8);
    if (xn > 0) {
        orc_fix_variable(F, f1, l1, e, g1, 1); orc_fix_variable(F, f3, l3, e, g3, 1); xn--;
    } else if (yn > 0) {
        orc_fix_variable(F, f1, l1, e, g1, 1); orc_fix_variable(F, f2, l2, e, g2, 1); yn--;
    } else {
        orc_fix_variable(F, f2, l2, e, g2, 1); orc_fix_variable(F, f3, l3, e, g3, 1); zn--;
    }
    orc_triangle_sum(F, g1, g2, g3, xn, yn, zn, out);
    free(g1); free(g2); free(g3);
}

/* ------------------------------------------------------------ a7: GKR W */
/* gkr-protocol/src/round_polynomial.rs:96-118 + sum: add/mul over (b:bn, c:cn)
 * with index (c << bn) | b (:108,123-125), w_b over b, w_c over c */
void orc_gkrw_sum(const orc_field *F, const uint64_t *add, const uint64_t *mul, const uint64_t *wb,
                  const uint64_t *wc, uint32_t bn, uint32_t cn, uint64_t *out) {
    const uint32_t n = F->n;
    uint64_t acc[ORC_MAX_LIMBS] = {0};
    for (size_t b = 0; b < ((size_t)1 << bn); ++b)
        for (size_t c = 0; c < ((size_t)1 << cn); ++c) {
            size_t bc = (c << bn) | b;
            uint64_t s[ORC_MAX_LIMBS], m[ORC_MAX_LIMBS];
            orc_add(F, s, wb + b * n, wc + c * n);
            orc_mul(F, s, add + bc * n, s);
            orc_mul(F, m, wb + b * n, wc + c * n);
            orc_mul(F, m, mul + bc * n, m);
            orc_add(F, s, s, m);
            orc_add(F, acc, acc, s);
        }
    memcpy(out, acc, 8 * n);
}
/* :78-84 for one point e; fold schedule :59-76 (add, mul always; w_b while bn > 0, else w_c) */
void orc_gkrw_round_eval_at(const orc_field *F, const uint64_t *add, const uint64_t *mul,
                            const uint64_t *wb, const uint64_t *wc, uint32_t bn, uint32_t cn,
                            const uint64_t *e, uint64_t *out) {
    const uint32_t n = F->n;
    size_t la = (size_t)1 << (bn + cn), lb = (size_t)1 << bn, lc = (size_t)1 << cn;
    uint64_t *a2 = (uint64_t *)malloc(la * n * 8), *m2 = (uint64_t *)malloc(la * n * 8),
             *b2 = (uint64_t *)malloc(lb * n * 8), *c2 = (uint64_t *)malloc(lc * n * 8);
    memcpy(b2, wb, lb * n * 8); memcpy(c2, wc, lc * n * 8);
    orc_fix_variable(F, add, la, e, a2, 1);
    orc_fix_variable(F, mul, la, e, m2, 1);
    if (bn > 0) { orc_fix_variable(F, wb, lb, e, b2, 1); bn--; }
    else        { orc_fix_variable(F, wc, lc, e, c2, 1); cn--; }
    orc_gkrw_sum(F, a2, m2, b2, c2, bn, cn, out);
    free(a2); free(m2); free(b2); free(c2);
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
