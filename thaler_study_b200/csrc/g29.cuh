// g29.cuh -- third-generation fused fold + round message for 4-limb fields: the arithmetic of lazy29.hpp (radix 2^29,
// lazy carries, full-rate IMAD.WIDE with no carry flags) inside the pass structure of g4.cuh (one point fewer thanks to
// the claim, leading coefficient instead of the highest point).
//
// Replaces, for round j >= 1 of a product polynomial,
//     self.g = self.g.fix_variables(&[r_prev]); self.g.to_univariate()     sum-check-protocol/src/lib.rs:105-112
// Per 4 adjacent entries of each of the K tables: 2 folds (t0 + r (t1 - t0), [ARK] fix_variables) whose results are made
// canonical in the 32-bit word domain and stored, then the factors lo, hi - lo (+ k p), hi + (hi - lo), ... of the folded
// pair are multiplied into the K running products; the last product of each point is added limb by limb into a 10-limb
// lazy sum.  What the kernel returns per point is the sum of x_1 .. x_K 2^(-261 (K-1)), i.e. the Montgomery-256 message
// sum times 2^(-5 (K-1)); the host multiplies 32^(K-1) back (engine.cu) -- exact field arithmetic either way.
#pragma once
#include <cstdint>

#include "g4.cuh"
#include "lazy29.hpp"

namespace scb {
namespace g29 {

using l29::Desc29;
using l29::L9;

struct Desc29x {      // Desc29 + the constants of the final per-thread reduction
    Desc29 d;
    uint32_t c1[9];   // 2^261 mod p  (mont(V, c1) = V mod p)
    uint32_t c2[9];   // 2^522 mod p  (mont(h, c2) = h 2^261 mod p)
};

struct A10 {  // lazy sum of normalised products: nine limbs + one limb of overflow (weight 2^261)
    uint32_t l[10];
};
__device__ __forceinline__ void acc_zero(A10& a) {
#pragma unroll
    for (int j = 0; j < 10; ++j) a.l[j] = 0;
}
__device__ __forceinline__ void acc_add(A10& a, const L9& x) {
#pragma unroll
    for (int j = 0; j < 9; ++j) a.l[j] += x.l[j];
}
__device__ __forceinline__ void acc_carry(A10& a) {
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const uint64_t v = (uint64_t)a.l[j] + c;
        a.l[j] = (uint32_t)v & l29::M29;
        c = (uint32_t)(v >> 29);
    }
    a.l[9] += c;
}

__device__ __forceinline__ L9 limbs_of(const uint64_t* w64) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w[2 * i] = (uint32_t)w64[i];
        w[2 * i + 1] = (uint32_t)(w64[i] >> 32);
    }
    return l29::from_words(w);
}
// lazy limbs (value below 2^256 + p) -> canonical words
__device__ __forceinline__ g4::W8 canonical(const g4::Arith& ar, const L9& x) {
    uint32_t w[8], top;
    l29::to_words(x, w, top);
    g4::W8 c = ar.reduce_once(w, top);
    return ar.reduce_once(c.w, 0);
}

template <int K, int MINB = 2>
__global__ void __launch_bounds__(kThreads, MINB)
    k_fold_round_g29(FieldDesc f, Desc29x dx, TabsIn<K> in, TabsOut<K> outp, ElemArg r5arg, uint64_t n_quads, uint64_t* partials, unsigned int* ticket,
                     uint64_t* out, PeerArg peer) {
    constexpr int NS = g4::n_sums(K);
    const g4::Arith ar(f);
    const Desc29& d = dx.d;
    const L9 r5 = limbs_of(r5arg.w);  // r * 2^5 (Montgomery-256 form), canonical
    A10 acc[NS];
#pragma unroll
    for (int x = 0; x < NS; ++x) acc_zero(acc[x]);
    uint32_t since_carry = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += stride) {
        L9 prod[NS];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint64_t w[16], o[8];
            ld_words<16>(in.p[k] + i * 16, w);
            L9 u[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const L9 t0 = limbs_of(w + 8 * h), t1 = limbs_of(w + 8 * h + 4);
                // t0 + r (t1 - t0): the difference as t1 - t0 + k p (limbs below 2^31), product below ((k+1) p / 2^261 + 1) p
                const L9 m = l29::mont(d, l29::sub_kp(d, t1, t0), r5);
                const g4::W8 c = canonical(ar, l29::add(t0, m));
                g4::store8(c, o + 4 * h);
                u[h] = l29::from_words(c.w);
            }
            st_words<8>(outp.p[k] + i * 8, o);
            // factors of this table at the points 0, inf, 2, 3: lo, hi - lo (+ k p), hi + (hi - lo), ...
            L9 fac[NS];
            fac[0] = u[0];
            if constexpr (K >= 2) {
                fac[1] = l29::sub_kp(d, u[1], u[0]);                                        // limbs < 1.5 2^30
#pragma unroll
                for (int x = 2; x < NS; ++x) {
                    fac[x] = l29::add(x == 2 ? u[1] : fac[x - 1], fac[1]);                  // x = 2: limbs <= 2^31
                    if (x >= 3) fac[x] = l29::normalise(fac[x]);                            // beyond the lazy limb bound
                }
            }
#pragma unroll
            for (int x = 0; x < NS; ++x) {
                if (k == 0 && K > 1) {
                    prod[x] = (x == 0 || x >= 3) ? fac[x] : l29::normalise(fac[x]);         // first operand of a product: normalised
                } else if (k < K - 1) {
                    prod[x] = l29::mont(d, prod[x], fac[x]);
                } else {
                    acc_add(acc[x], K == 1 ? fac[x] : l29::mont(d, prod[x], fac[x]));
                }
            }
        }
        if (++since_carry == 4) {  // limbs below 2^29: four sums stay below 2^32
            since_carry = 0;
#pragma unroll
            for (int x = 0; x < NS; ++x) acc_carry(acc[x]);
        }
    }
    // lazy sums -> canonical elements:  V = lo9 + h 2^261  =>  V mod p = mont(lo9, 2^261) + mont(h, 2^522)
    const PolGN<4> A(f);
    typename PolGN<4>::Acc fin[NS];
    L9 c1, c2;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        c1.l[j] = dx.c1[j];
        c2.l[j] = dx.c2[j];
    }
#pragma unroll
    for (int x = 0; x < NS; ++x) {
        acc_carry(acc[x]);
        L9 lo, hi;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            lo.l[j] = acc[x].l[j];
            hi.l[j] = 0;
        }
        hi.l[0] = acc[x].l[9] & l29::M29;
        hi.l[1] = acc[x].l[9] >> 29;
        const g4::W8 a = canonical(ar, l29::mont(d, lo, c1));
        const g4::W8 b = canonical(ar, l29::mont(d, hi, c2));
        const g4::W8 s = ar.add(a, b);
        uint64_t l[4];
        g4::store8(s, l);
        fin[x] = A.from_words(l);
    }
    grid_reduce_finish<PolGN<4>, NS>(A, fin, partials, ticket, out, 0, &peer);
}

// Round-0 message (Prover::new's pass; sum-check-protocol/src/lib.rs:88-97 with G::to_univariate,
// matrix-multiplication/src/lib.rs:110-122, generalised to K tables): no claim is known yet, so X = 1 is summed as well.
// One hypercube pair of every table per thread-iteration; sums at the points 0, inf, 2 .. K-1 and, last, 1 -- K + 1
// values, each short of 2^(5 (K-1)) like k_fold_round_g29's.
template <int K>
__global__ void __launch_bounds__(kThreads, 2)
    k_round_evals_g29(FieldDesc f, Desc29x dx, TabsIn<K> in, uint64_t n_pairs, uint64_t* partials, unsigned int* ticket, uint64_t* out, PeerArg peer) {
    constexpr int NS = g4::n_sums(K) + 1;  // + the point X = 1
    const g4::Arith ar(f);
    const Desc29& d = dx.d;
    A10 acc[NS];
#pragma unroll
    for (int x = 0; x < NS; ++x) acc_zero(acc[x]);
    uint32_t since_carry = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
        L9 prod[NS];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint64_t w[8];
            ld_words<8>(in.p[k] + i * 8, w);
            const L9 lo = limbs_of(w), hi = limbs_of(w + 4);
            L9 fac[NS];
            fac[0] = lo;
            fac[NS - 1] = hi;
            if constexpr (K >= 2) {
                fac[1] = l29::sub_kp(d, hi, lo);
#pragma unroll
                for (int x = 2; x < NS - 1; ++x) {
                    fac[x] = l29::add(x == 2 ? hi : fac[x - 1], fac[1]);
                    if (x >= 3) fac[x] = l29::normalise(fac[x]);
                }
            }
#pragma unroll
            for (int x = 0; x < NS; ++x) {
                const bool plain = x == 0 || x == NS - 1 || x >= 3;  // already normalised
                if (k == 0 && K > 1) {
                    prod[x] = plain ? fac[x] : l29::normalise(fac[x]);
                } else if (k < K - 1) {
                    prod[x] = l29::mont(d, prod[x], fac[x]);
                } else {
                    acc_add(acc[x], K == 1 ? fac[x] : l29::mont(d, prod[x], fac[x]));
                }
            }
        }
        if (++since_carry == 4) {
            since_carry = 0;
#pragma unroll
            for (int x = 0; x < NS; ++x) acc_carry(acc[x]);
        }
    }
    const PolGN<4> A(f);
    typename PolGN<4>::Acc fin[NS];
    L9 c1, c2;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        c1.l[j] = dx.c1[j];
        c2.l[j] = dx.c2[j];
    }
#pragma unroll
    for (int x = 0; x < NS; ++x) {
        acc_carry(acc[x]);
        L9 lo, hi;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            lo.l[j] = acc[x].l[j];
            hi.l[j] = 0;
        }
        hi.l[0] = acc[x].l[9] & l29::M29;
        hi.l[1] = acc[x].l[9] >> 29;
        const g4::W8 a = canonical(ar, l29::mont(d, lo, c1));
        const g4::W8 b = canonical(ar, l29::mont(d, hi, c2));
        const g4::W8 s = ar.add(a, b);
        uint64_t l[4];
        g4::store8(s, l);
        fin[x] = A.from_words(l);
    }
    grid_reduce_finish<PolGN<4>, NS>(A, fin, partials, ticket, out, 0, &peer);
}

}  // namespace g29
}  // namespace scb
