// pairs.cuh -- TWO sum-check rounds per pass over the tables (small-prime policy, product polynomials).
//
// Round j's message is g_j(X) = sum_{x'} prod_k f_k(X, x'), round j+1's is g_{j+1}(Y) = sum_{x''} prod_k f_k(r_j, Y, x'').
// Both are slices of ONE bivariate polynomial of degree K in each variable,
//     H(a, b) = sum_{x''} prod_k f_k(a, b, x''),        g_j(X) = H(X, 0) + H(X, 1),     g_{j+1}(Y) = H(r_j, Y),
// so a single pass that accumulates H on the (K+1) x (K+1) grid a, b in {0..K} yields both messages: the host takes
// g_j from the grid, hashes it into r_j, interpolates the grid rows at r_j (exact field arithmetic: the same field
// elements the reference's second pass would produce), hashes g_{j+1} into r_{j+1} -- and only then does the device
// touch the tables again, folding BOTH variables in one pass while accumulating the next grid.  The tables are
// streamed once per two rounds: with packed uint32 intermediates a proof moves 8 + 9 + 1.25 (1 + 1/4 + ...) = 18.7
// bytes per table entry instead of 24 (one round per pass) or 32, and half the rounds need no device
// pass -- hence no barrier, launch or host<->device turn-around -- at all.
//
// Multiplication count per entry is unchanged (the grid has (K+1)^2 products per 4 folded entries, the two line
// messages it replaces have 2 (K+1) per 2 + 2 (K+1) per ... entries); for p < 2^28 the passes stay HBM-bound.
#pragma once
#include <cstdint>

#include "kernels.cuh"
#include "persist.cuh"

namespace scb {

constexpr int kMaxGridPts = 25;  // (K+1)^2 for K <= 4

// Accumulates prod_k v_k(a, b) into acc[a * NP + b] for one 2x2 block of every table: c[k][y1 + 2 y2] canonical.
template <int K>
__device__ __forceinline__ void grid_accumulate(const PolSP& ar, const uint32_t (&c)[K][4], uint64_t (&acc)[(K + 1) * (K + 1)]) {
    constexpr int NP = K + 1;
    uint32_t P[NP * NP];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        // along y1, exact: va[y2][a] = c[y2][0] + a (c[y2][1] - c[y2][0])
        uint32_t va0[NP], va1[NP];
        const uint32_t d0 = ar.sub(c[k][1], c[k][0]), d1 = ar.sub(c[k][3], c[k][2]);
        va0[0] = c[k][0];
        va1[0] = c[k][2];
        va0[1] = c[k][1];
        va1[1] = c[k][3];
#pragma unroll
        for (int a = 2; a < NP; ++a) {
            va0[a] = ar.add(va0[a - 1], d0);
            va1[a] = ar.add(va1[a - 1], d1);
        }
        // along y2, lazy: v(a, b) = va0 + b (va1 - va0) < (2 NP - 1) p < 2^31
#pragma unroll
        for (int a = 0; a < NP; ++a) {
            const uint32_t e = ar.lz_diff(va1[a], va0[a]);
            uint32_t v = va0[a];
#pragma unroll
            for (int b = 0; b < NP; ++b) {
                if (b == 1) v = va1[a];
                else if (b > 1) v = ar.lz_add(v, e);
                P[a * NP + b] = k == 0 ? v : ar.msg_mul(P[a * NP + b], v);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NP * NP; ++i) ar.acc_add(acc[i], P[i]);
}

// Pass without a fold (the first pass of a proof, or the first after a consolidation): grid of the table as it is.
// One 2x2 block = 4 adjacent entries per table per thread-iteration.
template <int K, bool IN32, bool NC>
__device__ __forceinline__ void grid_pass_sp(const PolSP& ar, const uint64_t* const (&src)[K], uint64_t n_groups, uint64_t start, uint64_t stride,
                                             uint64_t (&acc)[(K + 1) * (K + 1)]) {
    for (uint64_t g = start; g < n_groups; g += stride) {
        uint32_t c[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if constexpr (IN32) {
                uint64_t w[2];
                ld_words_sel<2, NC>(src[k] + g * 2, w);
                c[k][0] = (uint32_t)w[0];
                c[k][1] = (uint32_t)(w[0] >> 32);
                c[k][2] = (uint32_t)w[1];
                c[k][3] = (uint32_t)(w[1] >> 32);
            } else {
                uint64_t w[4];
                ld_words_sel<4, NC>(src[k] + g * 4, w);
#pragma unroll
                for (int q = 0; q < 4; ++q) c[k][q] = (uint32_t)w[q];
            }
        }
        grid_accumulate<K>(ar, c, acc);
    }
}

// Fold two variables (challenges ra, rb as fold constants) and accumulate the grid of the folded table: 16 adjacent
// entries per table per thread-iteration -> one 2x2 block of the folded table, stored as 4 packed uint32.
template <int K, bool IN32, bool NC, bool GRID>
__device__ __forceinline__ void pair_pass_sp(const PolSP& ar, const PolSP::FoldC ra, const PolSP::FoldC rb, const uint64_t* const (&src)[K],
                                             uint64_t* const (&dst)[K], uint64_t n_groups, uint64_t start, uint64_t stride,
                                             uint64_t (&acc)[(K + 1) * (K + 1)]) {
    for (uint64_t g = start; g < n_groups; g += stride) {
        uint32_t c[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint32_t t[16];
            if constexpr (IN32) {
                uint64_t w[8];
                ld_words_sel<8, NC>(src[k] + g * 8, w);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    t[2 * q] = (uint32_t)w[q];
                    t[2 * q + 1] = (uint32_t)(w[q] >> 32);
                }
            } else {
                uint64_t w[16];
                ld_words_sel<16, NC>(src[k] + g * 16, w);
#pragma unroll
                for (int q = 0; q < 16; ++q) t[q] = (uint32_t)w[q];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t lo = ar.fold_c(t[4 * q], t[4 * q + 1], ra), hi = ar.fold_c(t[4 * q + 2], t[4 * q + 3], ra);
                c[k][q] = ar.fold_c(lo, hi, rb);
            }
            uint64_t o[2] = {(uint64_t)c[k][0] | ((uint64_t)c[k][1] << 32), (uint64_t)c[k][2] | ((uint64_t)c[k][3] << 32)};
            st_words<2>(dst[k] + g * 2, o);
        }
        if constexpr (GRID) grid_accumulate<K>(ar, c, acc);
    }
}

// ------------------------------------------------------------------------------------------ stand-alone kernels
// K6a  grid of a table as it is (Prover::new for the small-prime policy: c_1, g_1 and g_2 from one pass).
template <int K, bool IN32>
__global__ void __launch_bounds__(kThreads, 3)
    k_grid_sp(FieldDesc f, TabsIn<K> in, uint64_t n_groups, uint64_t* partials, unsigned int* ticket, uint64_t* out) {
    constexpr int NG = (K + 1) * (K + 1);
    const PolSP ar(f);
    uint64_t acc[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = 0;
    const uint64_t* src[K];
#pragma unroll
    for (int k = 0; k < K; ++k) src[k] = in.p[k];
    grid_pass_sp<K, IN32, true>(ar, src, n_groups, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, (uint64_t)gridDim.x * blockDim.x, acc);
    grid_reduce_finish<PolSP, NG>(ar, acc, partials, ticket, out, K, nullptr);
}
// K6b  one pair pass as its own launch (what the resident kernel runs per pass; used for profiling and as the
// non-resident path).
template <int K, bool IN32>
__global__ void __launch_bounds__(kThreads, 3)
    k_pair_pass_sp(FieldDesc f, TabsIn<K> in, TabsOut<K> outp, ElemArg ra_arg, ElemArg rb_arg, uint64_t n_groups, uint64_t* partials,
                   unsigned int* ticket, uint64_t* out) {
    constexpr int NG = (K + 1) * (K + 1);
    const PolSP ar(f);
    const PolSP::FoldC ra = ar.fold_const(ar.from_words(ra_arg.w)), rb = ar.fold_const(ar.from_words(rb_arg.w));
    uint64_t acc[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = 0;
    const uint64_t* src[K];
    uint64_t* dst[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        src[k] = in.p[k];
        dst[k] = outp.p[k];
    }
    pair_pass_sp<K, IN32, true, true>(ar, ra, rb, src, dst, n_groups, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x,
                                      (uint64_t)gridDim.x * blockDim.x, acc);
    grid_reduce_finish<PolSP, NG>(ar, acc, partials, ticket, out, K, nullptr);
}

// ------------------------------------------------------------------------------------------ resident kernel
// All pair passes of a proof in one cooperative launch.  Pass t folds the two lowest variables of a table of m
// entries-bits (m >= 3) by the challenge pair number t and posts: the (K+1)^2 grid of the folded table when it still
// has >= 2 variables, its (K+1) line sums when it has exactly one.  n_passes = ceil((m - 2) / 2).  Barrier, mailbox
// and solo-CTA endgame as in k_persist_rounds.
template <int K>
__global__ void __launch_bounds__(kThreads, (K <= 2 ? 3 : 2))
    k_persist_pairs_sp(FieldDesc f, TabsIn<K> in0, TabsOut<K> buf_a, TabsOut<K> buf_b, ElemArg ra0, ElemArg rb0, uint32_t m, uint32_t n_passes,
                       int in0_w32, TailMailbox* mb, PersistCtl* ctl, uint64_t* partials, uint64_t timeout_ns) {
    using A = PolSP;
    constexpr int NP = K + 1, NG = NP * NP;
    bool src_w32 = in0_w32 != 0;
    const A ar(f);
    __shared__ uint64_t sm[32 * NG];
    __shared__ uint64_t r_sm[2];
    __shared__ int flag_sm;  // 1: this CTA took the last ticket, 2: abort
    if (threadIdx.x == 0) {
        r_sm[0] = ra0.w[0];
        r_sm[1] = rb0.w[0];
        if (blockIdx.x == 0) st_sys(&mb->stamp[2 * kTailMaxRounds + 1], globaltimer_ns());  // kernel start
    }
    __syncthreads();
    bool have_r = true;
    for (uint32_t t = 0; t < n_passes; ++t) {
        // m >= 4: 2^(m-4) thread-iterations of 16 entries; m == 3: one thread folds 8 entries to a line
        const uint64_t n_groups = m >= 4 ? 1ull << (m - 4) : 1;
        uint64_t active = (n_groups + blockDim.x - 1) / blockDim.x;
        if (active > gridDim.x) active = gridDim.x;
        const bool solo = active == 1;
        if (solo && blockIdx.x != 0) return;
        if (!have_r) {  // barrier + challenge pair: released by the CTA that finished pass t-1
            if (threadIdx.x == 0) {
                uint64_t c[2];
                const uint64_t t0 = globaltimer_ns();
                int bad = 0;
                for (;;) {
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(c[0]), "=l"(c[1]) : "l"(ctl->challenge) : "memory");
                    if ((uint32_t)(c[0] >> 32) == t && (uint32_t)(c[1] >> 32) == t) break;
                    if ((uint32_t)(c[0] >> 32) == kMbAbortTag || globaltimer_ns() - t0 > 4 * timeout_ns) {
                        bad = 1;
                        break;
                    }
                }
                asm volatile("fence.acq_rel.gpu;" ::: "memory");  // pass t-1's folded entries before this pass's loads
                r_sm[0] = (uint32_t)c[0];
                r_sm[1] = (uint32_t)c[1];
                flag_sm = bad ? 2 : 0;
            }
            __syncthreads();
            if (flag_sm == 2) return;
        }
        const uint64_t* src[K];
        uint64_t* dst[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            src[k] = t == 0 ? in0.p[k] : ((t & 1) ? buf_a.p[k] : buf_b.p[k]);
            dst[k] = (t & 1) ? buf_b.p[k] : buf_a.p[k];
        }
        const A::FoldC ra = ar.fold_const((uint32_t)r_sm[0]), rb = ar.fold_const((uint32_t)r_sm[1]);
        uint64_t acc[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) acc[i] = 0;
        const bool grid_out = m >= 4;  // else (m == 3) a line of NP sums in acc[0..NP)
        if (blockIdx.x < active) {
            const uint64_t start = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = active * blockDim.x;
            if (m >= 4) {
                if (!src_w32) pair_pass_sp<K, false, true, true>(ar, ra, rb, src, dst, n_groups, start, stride, acc);  // only the caller's tables
                else if (t == 0) pair_pass_sp<K, true, true, true>(ar, ra, rb, src, dst, n_groups, start, stride, acc);
                else pair_pass_sp<K, true, false, true>(ar, ra, rb, src, dst, n_groups, start, stride, acc);
            } else if (start == 0) {  // 8 entries per table -> 2 -> line sums; nothing reads the folded pair again
                uint32_t prod[NP];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    uint32_t e[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (src_w32) {
                            const uint64_t w = __ldcg(src[k] + q / 2);
                            e[q] = (q & 1) ? (uint32_t)(w >> 32) : (uint32_t)w;
                        } else {
                            e[q] = (uint32_t)__ldcg(src[k] + q);
                        }
                    }
                    const uint32_t u0 = ar.fold_c(ar.fold_c(e[0], e[1], ra), ar.fold_c(e[2], e[3], ra), rb);
                    const uint32_t u1 = ar.fold_c(ar.fold_c(e[4], e[5], ra), ar.fold_c(e[6], e[7], ra), rb);
                    pair_into_prod<A, NP>(ar, k == 0, u0, u1, prod);
                }
#pragma unroll
                for (int x = 0; x < NP; ++x) acc[x] += prod[x];
            }
            __threadfence();  // folded entries are visible device-wide before this CTA's ticket / next pass's loads
            block_reduce<A, NG>(ar, acc, sm);
            bool finisher = solo;
            if (!solo) {
                if (threadIdx.x == 0) {
#pragma unroll
                    for (int i = 0; i < NG; ++i) __stcg(&partials[(size_t)blockIdx.x * NG + i], acc[i]);
                    __threadfence();
                    const unsigned int tk = atomicAdd(&ctl->ticket[t], 1u);
                    flag_sm = (tk == (unsigned int)active - 1) ? 1 : 0;
                }
                __syncthreads();
                finisher = flag_sm == 1;
                if (finisher) {
                    __threadfence();
#pragma unroll
                    for (int i = 0; i < NG; ++i) acc[i] = 0;
                    for (unsigned int b = threadIdx.x; b < (unsigned int)active; b += blockDim.x) {
#pragma unroll
                        for (int i = 0; i < NG; ++i) acc[i] += __ldcg(&partials[(size_t)b * NG + i]);
                    }
                    block_reduce<A, NG>(ar, acc, sm);
                }
            }
            if (finisher && threadIdx.x == 0) {
                const uint64_t hi = (uint64_t)(t + 1) << 32;
                const int n_out = grid_out ? NG : NP;
#pragma unroll
                for (int i = 0; i < NG; ++i)
                    if (i < n_out) st_sys(&mb->evals[i], hi | ar.msg_final(acc[i], K));
                st_sys(&mb->stamp[2 * t], globaltimer_ns());
                int bad = 0;
                if (t + 1 < n_passes) {
                    uint64_t c[2];
                    const uint64_t t0 = globaltimer_ns();
                    for (;;) {
                        asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(c[0]), "=l"(c[1]) : "l"(mb->challenge) : "memory");
                        if ((uint32_t)(c[0] >> 32) == t + 1 && (uint32_t)(c[1] >> 32) == t + 1) break;
                        if ((uint32_t)(c[0] >> 32) == kMbAbortTag || globaltimer_ns() - t0 > timeout_ns) {
                            bad = 1;
                            break;
                        }
                    }
                    if (bad) {
                        st_gpu(&ctl->challenge[0], (uint64_t)kMbAbortTag << 32);
                        st_sys(&mb->dev_status, 2);
                    } else {
                        r_sm[0] = (uint32_t)c[0];
                        r_sm[1] = (uint32_t)c[1];
                        if (!solo) {
                            st_gpu(&ctl->challenge[0], hi | (uint32_t)c[0]);
                            st_gpu(&ctl->challenge[1], hi | (uint32_t)c[1]);
                        }
                        st_sys(&mb->stamp[2 * t + 1], globaltimer_ns());
                    }
                } else {
                    st_sys(&mb->dev_status, 1);
                }
                if (solo) flag_sm = bad ? 2 : 0;
            }
            if (solo) {
                __syncthreads();
                if (flag_sm == 2) return;
            }
        }
        have_r = solo;
        m -= 2;
        src_w32 = true;
    }
}

}  // namespace scb
