// eqfix.cuh -- eq/chi tables and the kernels that contract a table against one.
//
// * k_eq_tables: builds up to two eq tables (low / high index bits) by parallel doubling, in shared memory, with the
//   point passed as a kernel argument (no allocation, no H2D copy): multilinear-extensions/src/lib.rs:9-18.
// * k_fix_low_eq / k_fix_high_eq: [ARK] fix_variables of SEVERAL variables in ONE pass -- out[j] = sum_i eq(r; i) *
//   t[...] -- instead of one fold pass per variable.  Exact field arithmetic, so the result equals the reference's
//   sequence of folds element for element.  Used by G::new (matrix-multiplication/src/lib.rs:77-92), where the
//   set-up folds are ~1000x the work of the sum-check itself: f_A is obtained by fixing the HIGH
//   variable block of the row-major matrix directly, which equals relabel(0,n,n) followed by fixing the low block.
#pragma once
#include <cstdint>

#include "kernels.cuh"

namespace scb {

constexpr int kMaxPointCoords = 40;
struct PointArg {  // coordinate bound to index bit j at w[j * n_limbs ...]
    uint64_t w[kMaxPointCoords * kMaxLimbs];
};

// Block b builds table b: bits [first_b, first_b + nb_b) of the index, first_0 = 0 / nb_0 = lb, first_1 = lb / nb_1 = v - lb.
// Level l appends index bit l:  t[i + 2^l] = t[i]*c ; t[i] -= t[i + 2^l]  (= t[i]*(1-c)).  Levels that fit are done in
// shared memory (cap_bits), the rest in global memory.
template <class A>
__global__ void __launch_bounds__(1024) k_eq_tables(FieldDesc f, PointArg pt, uint32_t lb, uint32_t v, uint64_t* lo_tab, uint64_t* hi_tab,
                                                    uint32_t cap_bits) {
    constexpr int N = A::N;
    extern __shared__ uint64_t eq_sm[];
    const A ar(f);
    const uint32_t first = blockIdx.x == 0 ? 0 : lb;
    const uint32_t nb = blockIdx.x == 0 ? lb : v - lb;
    uint64_t* out = blockIdx.x == 0 ? lo_tab : hi_tab;
    const uint32_t nb_sm = nb < cap_bits ? nb : cap_bits;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) eq_sm[i] = f.one[i];
    }
    __syncthreads();
    for (uint32_t l = 0; l < nb_sm; ++l) {
        const typename A::El c = ar.from_words(pt.w + (size_t)(first + l) * N);
        const uint32_t half = 1u << l;
        for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
            typename A::El cur = ar.from_words(eq_sm + (size_t)i * N);
            typename A::El hi = ar.mul(cur, c);
            ar.to_words(hi, eq_sm + (size_t)(i + half) * N);
            ar.to_words(ar.sub(cur, hi), eq_sm + (size_t)i * N);
        }
        __syncthreads();
    }
    const uint64_t words = ((uint64_t)N) << nb_sm;
    for (uint64_t i = threadIdx.x; i < words; i += blockDim.x) out[i] = eq_sm[i];
    __syncthreads();
    for (uint32_t l = nb_sm; l < nb; ++l) {  // large tables: remaining levels in global memory
        const typename A::El c = ar.from_words(pt.w + (size_t)(first + l) * N);
        const uint64_t half = 1ull << l;
        for (uint64_t i = threadIdx.x; i < half; i += blockDim.x) {
            typename A::El cur = ar.from_words(out + i * N);
            typename A::El hi = ar.mul(cur, c);
            ar.to_words(hi, out + (i + half) * N);
            ar.to_words(ar.sub(cur, hi), out + i * N);
        }
        __syncthreads();
    }
}

// The same two tables with every level but the last done on sub-tables: a table over nb index bits is the outer product
// of one over its low nb/2 bits and one over the rest (exact arithmetic: the same field elements as nb doublings), so
// each block doubles two sub-tables of at most 2^9 entries side by side in shared memory and then writes its slice of
// the products -- the last, widest doubling step spread over the whole grid instead of one block's loop in global
// memory.  Blocks [0, blocks_lo) write the low table, the rest the high table.
template <class A>
__global__ void __launch_bounds__(1024) k_eq_tables_split(FieldDesc f, PointArg pt, uint32_t lb, uint32_t v, uint64_t* lo_tab, uint64_t* hi_tab,
                                                          uint32_t blocks_lo) {
    constexpr int N = A::N;
    extern __shared__ uint64_t eq_sm[];
    const A ar(f);
    const bool is_lo = blockIdx.x < blocks_lo;
    const uint32_t first = is_lo ? 0 : lb;
    const uint32_t nb = is_lo ? lb : v - lb;
    uint64_t* out = is_lo ? lo_tab : hi_tab;
    const uint32_t blk = is_lo ? blockIdx.x : blockIdx.x - blocks_lo;
    const uint32_t nblk = is_lo ? blocks_lo : gridDim.x - blocks_lo;
    const uint32_t h0 = nb / 2, h1 = nb - h0;  // h1 >= h0
    uint64_t* ta = eq_sm;                      // index bits [0, h0)
    uint64_t* tb = eq_sm + ((size_t)N << h0);  // index bits [h0, nb)
    const uint32_t half_threads = blockDim.x / 2;
    const bool in_a = threadIdx.x < half_threads;
    const uint32_t tid = in_a ? threadIdx.x : threadIdx.x - half_threads;
    uint64_t* tab = in_a ? ta : tb;
    const uint32_t bits = in_a ? h0 : h1, cfirst = first + (in_a ? 0 : h0);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) tab[i] = f.one[i];
    }
    __syncthreads();
    for (uint32_t l = 0; l < h1; ++l) {
        if (l < bits) {
            const typename A::El c = ar.from_words(pt.w + (size_t)(cfirst + l) * N);
            const uint32_t half = 1u << l;
            for (uint32_t i = tid; i < half; i += half_threads) {
                typename A::El cur = ar.from_words(tab + (size_t)i * N);
                typename A::El hi = ar.mul(cur, c);
                ar.to_words(hi, tab + (size_t)(i + half) * N);
                ar.to_words(ar.sub(cur, hi), tab + (size_t)i * N);
            }
        }
        __syncthreads();
    }
    const uint64_t n_out = 1ull << nb, mask = (1ull << h0) - 1;
    for (uint64_t i = (uint64_t)blk * blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)nblk * blockDim.x) {
        const typename A::El a = ar.from_words(ta + (size_t)(i & mask) * N), b = ar.from_words(tb + (size_t)(i >> h0) * N);
        ar.to_words(ar.mul(a, b), out + i * N);
    }
}

// Same for T points at once (coordinates in device memory, pts[t][j] bound to index bit j): block 2t builds point t's
// low table, block 2t+1 its high table; both must fit shared memory (lb, v - lb <= cap_bits).  Tables are stored
// point-major: lo_all[t << lb ...], hi_all[t << (v - lb) ...].
template <class A>
__global__ void __launch_bounds__(1024) k_eq_tables_multi(FieldDesc f, const uint64_t* __restrict__ pts, uint32_t lb, uint32_t v, uint64_t* lo_all,
                                                          uint64_t* hi_all) {
    constexpr int N = A::N;
    extern __shared__ uint64_t eq_sm[];
    const A ar(f);
    const uint32_t t = blockIdx.x >> 1, half_id = blockIdx.x & 1;
    const uint32_t first = half_id == 0 ? 0 : lb;
    const uint32_t nb = half_id == 0 ? lb : v - lb;
    uint64_t* out = half_id == 0 ? lo_all + ((size_t)t << lb) * N : hi_all + ((size_t)t << (v - lb)) * N;
    const uint64_t* coords = pts + (size_t)t * v * N;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) eq_sm[i] = f.one[i];
    }
    __syncthreads();
    for (uint32_t l = 0; l < nb; ++l) {
        uint64_t cw[N];
#pragma unroll
        for (int i = 0; i < N; ++i) cw[i] = __ldg(coords + (size_t)(first + l) * N + i);
        const typename A::El c = ar.from_words(cw);
        const uint32_t half = 1u << l;
        for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
            typename A::El cur = ar.from_words(eq_sm + (size_t)i * N);
            typename A::El hi = ar.mul(cur, c);
            ar.to_words(hi, eq_sm + (size_t)(i + half) * N);
            ar.to_words(ar.sub(cur, hi), eq_sm + (size_t)i * N);
        }
        __syncthreads();
    }
    const uint64_t words = ((uint64_t)N) << nb;
    for (uint64_t i = threadIdx.x; i < words; i += blockDim.x) out[i] = eq_sm[i];
}

// TC evaluations of ONE table in one pass: out[t] = sum_i evals[i] * lo_t[i & (2^lb - 1)] * hi_t[i >> lb].  The TC low
// tables are staged in shared memory; every entry of the table is loaded once and used TC times (restrict_poly needs
// k + 1 evaluations of the same W along a line, gkr-protocol/src/lib.rs:291-321).
template <class A, int TC>
__global__ void __launch_bounds__(kThreads) k_mle_dot_multi(FieldDesc f, const uint64_t* __restrict__ evals, const uint64_t* __restrict__ lo_all,
                                                            const uint64_t* __restrict__ hi_all, uint32_t lb, uint32_t v, uint32_t n_pts, uint64_t n,
                                                            uint64_t* partials, unsigned int* ticket, uint64_t* out) {
    constexpr int N = A::N;
    constexpr int VEC = N == 1 ? 4 : 1;  // one-limb fields: four consecutive entries per 256-bit load (same row: lb >= 2)
    extern __shared__ uint64_t lo_sm[];
    const A ar(f);
    const uint64_t lo_words = ((uint64_t)n_pts << lb) * N;
    for (uint64_t i = threadIdx.x; i < lo_words; i += blockDim.x) lo_sm[i] = lo_all[i];
    __syncthreads();
    typename A::Acc acc[TC];
#pragma unroll
    for (int t = 0; t < TC; ++t) ar.acc_zero(acc[t]);
    const uint64_t mask = (1ull << lb) - 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_groups = n / VEC;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        uint64_t w[VEC * N];
        ld_words<VEC * N>(evals + g * VEC * N, w);
        const uint64_t i = g * VEC;
        const uint64_t il = i & mask, ih = i >> lb;
#pragma unroll
        for (int t = 0; t < TC; ++t) {
            if (t < (int)n_pts) {
                uint64_t hw[N];
#pragma unroll
                for (int q = 0; q < N; ++q) hw[q] = __ldg(hi_all + (((size_t)t << (v - lb)) + ih) * N + q);
                typename A::Lz m;
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const typename A::Lz me = ar.lz_mul(ar.lz(ar.from_words(w + e * N)), ar.lz(ar.from_words(lo_sm + (((size_t)t << lb) + il + e) * N)));
                    m = e == 0 ? me : ar.lz_add(m, me);
                }
                ar.acc_add(acc[t], ar.lz_mul(m, ar.lz(ar.from_words(hw))));
            }
        }
    }
    grid_reduce_finish<A, TC>(ar, acc, partials, ticket, out);
}

// The same for one-limb fields and tables of at most 2^20 entries (the GKR layer width of BASELINE configs[4]), row-wise
// like k_mle_eval_fused: lb is fixed at 8, so the TC low tables are 16 KB of shared memory (many CTAs per SM, staged
// with one 256-bit load per thread), a WARP takes a row of 256 entries -- two 256-bit loads per lane, read once and used
// for all TC points -- and multiplies each point's row sum by that point's high-table entry.  k_mle_dot_multi staged
// 64 KB per CTA through 32 dependent 8-byte loads per thread and ran one CTA per SM: 66-99 us per launch for an 8 MB
// table (profiles/r02_launches_gkr.csv); this form takes a few microseconds.
constexpr int kRowsMultiLB = 8;
template <class A, int TC>
__global__ void __launch_bounds__(kThreads) k_mle_rows_multi(FieldDesc f, const uint64_t* __restrict__ evals, const uint64_t* __restrict__ lo_all,
                                                             const uint64_t* __restrict__ hi_all, uint32_t v, uint32_t n_pts, uint64_t n,
                                                             uint64_t* partials, unsigned int* ticket, uint64_t* out) {
    static_assert(A::N == 1, "one-limb fields only");
    constexpr int LB = kRowsMultiLB, JL = (1 << LB) / 128;
    extern __shared__ __align__(32) uint64_t lo_sm[];  // n_pts << LB words (TC = 24: up to 48 KB, one launch for the 21 points of a 2^20-wide layer)
    const A ar(f);
    const int lane = threadIdx.x & 31;
    const uint64_t n_rows = n >> LB;
    const uint64_t n_warps = (uint64_t)gridDim.x * (kThreads / 32);
    uint64_t row = (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    uint64_t w[JL][4];
    if (row < n_rows) {  // the row's loads are in flight while the low tables are staged
#pragma unroll
        for (int j = 0; j < JL; ++j) ld_words<4>(evals + ((row << LB) + (size_t)(j * 32 + lane) * 4), w[j]);
    }
    for (uint32_t i = threadIdx.x * 4; i < (n_pts << LB); i += kThreads * 4) {
        uint64_t q[4];
        ld_words<4>(lo_all + i, q);
#pragma unroll
        for (int e = 0; e < 4; ++e) lo_sm[i + e] = q[e];
    }
    __syncthreads();
    typename A::Acc acc[TC];
#pragma unroll
    for (int t = 0; t < TC; ++t) ar.acc_zero(acc[t]);
    while (row < n_rows) {
#pragma unroll
        for (int t = 0; t < TC; ++t) {
            if (t < (int)n_pts) {
                typename A::Lz s;
#pragma unroll
                for (int j = 0; j < JL; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const typename A::Lz m = ar.lz_mul(ar.lz(ar.from_words(&w[j][e])), ar.lz(ar.from_words(lo_sm + ((size_t)t << LB) + (size_t)(j * 32 + lane) * 4 + e)));
                        s = (j == 0 && e == 0) ? m : ar.lz_add(s, m);
                    }
                const uint64_t hw = __ldg(hi_all + (((size_t)t << (v - LB)) + row));
                ar.acc_add(acc[t], ar.lz_mul(s, ar.lz(ar.from_words(&hw))));
            }
        }
        row += n_warps;
        if (row < n_rows) {
#pragma unroll
            for (int j = 0; j < JL; ++j) ld_words<4>(evals + ((row << LB) + (size_t)(j * 32 + lane) * 4), w[j]);
        }
    }
    grid_reduce_finish<A, TC>(ar, acc, partials, ticket, out);
}

// ------------------------------------------------------------------------------------------
// MLE evaluation in ONE launch (multilinear-extensions/src/lib.rs:6-24; [ARK] evaluate for the LSB-first order).
// eq(i) over v index bits is the outer product of a table over the low LB bits and up to three sub-tables of at most
// 9 bits each over the rest (exact arithmetic: the same field elements as v doublings).  Every CTA builds all of them
// in its own shared memory by parallel doubling (four thread groups side by side, at most 10 levels), so nothing has
// to be written to global memory or waited for; then a WARP takes a row (2^LB consecutive entries, one value of the
// high bits): each lane adds up its entries' products with the low table and multiplies the sum by the row's high
// factor -- 1 + 3/EPL multiplications per entry -- and the last CTA to finish adds the per-CTA partial sums
// (grid_reduce_finish; sharded: the finishing thread also adds the peer GPUs' sums over NVLink, `row0` being this
// rank's first row of the whole table).
// ------------------------------------------------------------------------------------------
template <class A>
struct MleFusedCfg {
    static constexpr int LB = A::N == 1 ? 8 : 10;      // index bits of a row
    static constexpr int SUB = 9;                       // bits per high sub-table
    static constexpr int MAX_HB = 3 * SUB;              // high bits covered
    static constexpr size_t smem_bytes = ((size_t)8 * A::N << LB) + 3 * ((size_t)8 * A::N << SUB);
};
template <class A>
__global__ void __launch_bounds__(kThreads) k_mle_eval_fused(FieldDesc f, PointArg pt, const uint64_t* __restrict__ evals, uint32_t v_local,
                                                            uint32_t v_total, uint64_t row0, uint64_t* partials, unsigned int* ticket,
                                                            uint64_t* out, PeerArg peer) {
    constexpr int N = A::N, LB = MleFusedCfg<A>::LB, SUB = MleFusedCfg<A>::SUB;
    extern __shared__ uint64_t eq_sm[];
    const A ar(f);
    const uint32_t hb = v_total - LB;
    const uint32_t nsub = hb == 0 ? 0 : (hb + SUB - 1) / SUB;  // <= 3
    const uint32_t q = nsub ? hb / nsub : 0, rem = nsub ? hb % nsub : 0;
    const uint32_t b1 = nsub > 0 ? q + (0 < rem ? 1 : 0) : 0, b2 = nsub > 1 ? q + (1 < rem ? 1 : 0) : 0, b3 = nsub > 2 ? q + (2 < rem ? 1 : 0) : 0;
    uint64_t* const t1 = eq_sm + ((size_t)N << LB);
    uint64_t* const t2 = t1 + ((size_t)N << SUB);
    uint64_t* const t3 = t2 + ((size_t)N << SUB);
    const int lane = threadIdx.x & 31;
    const uint64_t n_rows = 1ull << (v_local - LB);
    const uint64_t n_warps = (uint64_t)gridDim.x * (kThreads / 32);
    // one-limb fields: the loads of a warp's first row are issued BEFORE the eq tables are built, and every later row is
    // loaded while the previous one is multiplied (register double buffer) -- at 2^24 entries a warp only sees ~7 rows, so
    // a load-wait-compute chain per row left the kernel latency-bound (38 us for a 21 us read)
    constexpr int JL1 = A::N == 1 ? (1 << LB) / 128 : 1;  // 256-bit loads per lane per row
    uint64_t wn[JL1][4];
    uint64_t row_n = (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if constexpr (N == 1) {
        if (row_n < n_rows) {
#pragma unroll
            for (int j = 0; j < JL1; ++j) ld_words<4>(evals + ((row_n << LB) + (size_t)(j * 32 + lane) * 4), wn[j]);
        }
    }
    {  // thread group g (64 threads) doubles table g: g = 0 the low table, 1..3 the high sub-tables
        const uint32_t g = threadIdx.x >> 6, tid = threadIdx.x & 63;
        uint64_t* const tg = g == 0 ? eq_sm : (g == 1 ? t1 : (g == 2 ? t2 : t3));
        const uint32_t gbits = g == 0 ? (uint32_t)LB : (g == 1 ? b1 : (g == 2 ? b2 : b3));
        const uint32_t gfirst = g == 0 ? 0u : (g == 1 ? (uint32_t)LB : (g == 2 ? LB + b1 : LB + b1 + b2));
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) tg[i] = f.one[i];
        }
        __syncthreads();
        for (uint32_t l = 0; l < (uint32_t)(LB > SUB ? LB : SUB); ++l) {
            if (l < gbits) {
                const typename A::El c = ar.from_words(pt.w + (size_t)(gfirst + l) * N);
                const uint32_t half = 1u << l;
                for (uint32_t i = tid; i < half; i += 64) {
                    typename A::El cur = ar.from_words(tg + (size_t)i * N);
                    typename A::El hi = ar.mul(cur, c);
                    ar.to_words(hi, tg + (size_t)(i + half) * N);
                    ar.to_words(ar.sub(cur, hi), tg + (size_t)i * N);
                }
            }
            __syncthreads();
        }
    }
    typename A::Acc acc[1];
    ar.acc_zero(acc[0]);
    const uint64_t m1 = (1ull << b1) - 1, m2 = (1ull << b2) - 1;
    for (uint64_t row = (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); row < n_rows; row += n_warps) {
        const uint64_t* src = evals + ((row << LB) * N);
        typename A::Lz s;
        if constexpr (N == 1) {
            constexpr int JL = JL1;
            uint64_t w[JL][4];
#pragma unroll
            for (int j = 0; j < JL; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) w[j][e] = wn[j][e];
            row_n = row + n_warps;
            if (row_n < n_rows) {
#pragma unroll
                for (int j = 0; j < JL; ++j) ld_words<4>(evals + ((row_n << LB) + (size_t)(j * 32 + lane) * 4), wn[j]);
            }
#pragma unroll
            for (int j = 0; j < JL; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const typename A::Lz m = ar.lz_mul(ar.lz(ar.from_words(&w[j][e])), ar.lz(ar.from_words(eq_sm + (size_t)(j * 32 + lane) * 4 + e)));
                    s = (j == 0 && e == 0) ? m : ar.lz_add(s, m);
                }
        } else {
            constexpr int JL = (1 << LB) / 32;  // one element (256-bit load) per lane per step
#pragma unroll 1
            for (int j0 = 0; j0 < JL; j0 += 4) {
                uint64_t w[4][N];
#pragma unroll
                for (int j = 0; j < 4; ++j) ld_words<N>(src + (size_t)((j0 + j) * 32 + lane) * N, w[j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const typename A::Lz m = ar.lz_mul(ar.lz(ar.from_words(w[j])), ar.lz(ar.from_words(eq_sm + (size_t)((j0 + j) * 32 + lane) * N)));
                    s = (j0 == 0 && j == 0) ? m : ar.lz_add(s, m);
                }
            }
        }
        if (nsub > 0) {
            const uint64_t ih = row0 + row;
            typename A::Lz h = ar.lz(ar.from_words(t1 + (size_t)(ih & m1) * N));
            if (nsub > 1) h = ar.lz_mul(h, ar.lz(ar.from_words(t2 + (size_t)((ih >> b1) & m2) * N)));
            if (nsub > 2) h = ar.lz_mul(h, ar.lz(ar.from_words(t3 + (size_t)(ih >> (b1 + b2)) * N)));
            s = ar.lz_mul(s, h);
        }
        ar.acc_add(acc[0], s);
    }
    grid_reduce_finish<A, 1>(ar, acc, partials, ticket, out, 0, &peer);
}

// out[j] = sum_{i < 2^m} eq[i] * t[j * 2^m + i]      (the LOW m variables fixed); one warp per output
template <class A>
__global__ void __launch_bounds__(kThreads) k_fix_low_eq(FieldDesc f, const uint64_t* __restrict__ tab, const uint64_t* __restrict__ eq,
                                                         uint32_t m, uint64_t* __restrict__ outp, uint64_t n_out) {
    constexpr int N = A::N, AW = A::AW;
    const A ar(f);
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t len = 1ull << m;
    for (uint64_t j = warp0; j < n_out; j += n_warps) {
        typename A::Acc acc;
        ar.acc_zero(acc);
        for (uint64_t i = lane; i < len; i += 32) {
            uint64_t e[N];
#pragma unroll
            for (int q = 0; q < N; ++q) e[q] = __ldg(eq + i * N + q);
            ar.acc_add(acc, ar.lz_mul(ar.lz(ld_el(ar, tab, j * len + i)), ar.lz(ar.from_words(e))));
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            uint64_t w[AW];
            ar.acc_to_words(acc, w);
#pragma unroll
            for (int q = 0; q < AW; ++q) w[q] = __shfl_xor_sync(0xffffffffu, w[q], off);
            typename A::Acc o;
            ar.acc_from_words(o, w);
            ar.acc_merge(acc, o);
        }
        if (lane == 0) {
            uint64_t o[N];
            ar.to_words(ar.acc_final(acc), o);
            st_words<N>(outp + j * N, o);
        }
    }
}

// out[j] = sum_{i < 2^m} eq[i] * t[i * n_out + j]    (the HIGH m variables fixed): a vector-matrix product.
// Thread (j, slice s) accumulates rows i = s, s + S, s + 2S, ... (rows are read coalesced across j, several
// independent loads in flight); the S partial rows go to scratch and k_fix_high_finish adds them.
template <class A>
__global__ void __launch_bounds__(kThreads) k_fix_high_eq(FieldDesc f, const uint64_t* __restrict__ tab, const uint64_t* __restrict__ eq,
                                                          uint32_t m, uint64_t* __restrict__ scratch, uint64_t n_out, uint32_t n_slices) {
    constexpr int N = A::N;
    const A ar(f);
    const uint64_t len = 1ull << m;
    const uint64_t total = n_out * n_slices;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint64_t j = t % n_out, s = t / n_out;
        typename A::Acc acc;
        ar.acc_zero(acc);
#pragma unroll 4
        for (uint64_t i = s; i < len; i += n_slices) {
            uint64_t e[N];
#pragma unroll
            for (int q = 0; q < N; ++q) e[q] = __ldg(eq + i * N + q);
            ar.acc_add(acc, ar.lz_mul(ar.lz(ld_el(ar, tab, i * n_out + j)), ar.lz(ar.from_words(e))));
        }
        uint64_t o[N];
        ar.to_words(ar.acc_final(acc), o);
        st_words<N>(scratch + t * N, o);
    }
}
template <class A>
__global__ void __launch_bounds__(kThreads) k_fix_high_finish(FieldDesc f, const uint64_t* __restrict__ scratch, uint64_t* __restrict__ outp,
                                                              uint64_t n_out, uint32_t n_slices) {
    constexpr int N = A::N;
    const A ar(f);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += stride) {
        typename A::El acc = ld_el(ar, scratch, j);
        for (uint32_t s = 1; s < n_slices; ++s) acc = ar.add(acc, ld_el(ar, scratch, (uint64_t)s * n_out + j));
        uint64_t o[N];
        ar.to_words(acc, o);
        st_words<N>(outp + j * N, o);
    }
}

}  // namespace scb
