// g4.cuh -- second-generation fused fold + round message for 4-limb fields (BLS12-381 Fr =
// ark_ed_on_bls12_381::Fq, /root/reference/Cargo.toml:20; [ARK] Fp<MontBackend<_,4>,4>).
//
// Replaces, for round j >= 1 of a product polynomial,
//     self.g = self.g.fix_variables(&[r_prev]); self.g.to_univariate()     sum-check-protocol/src/lib.rs:105-112
// like k_fold_round (kernels.cuh), with three changes that cut the instruction count per 4 table entries from ~4100 to
// ~3000 (profiles/r01_ncu_bls_fold_round_evenodd.md found the kernel issue-bound at IPC 0.38, not HBM-bound):
//
//  1. One point fewer.  The prover knows the claim g_j(0) + g_j(1) = g_{j-1}(r_{j-1}) before the pass starts, so g_j(1)
//     is NOT accumulated; and the leading coefficient (the "point at infinity", prod_k (hi_k - lo_k)) replaces the
//     highest finite point, which saves the repeated additions that build lo + X (hi - lo).  Per folded pair the pass
//     accumulates  S_0 = prod lo_k,  S_inf = prod (hi_k - lo_k),  S_x = prod (lo_k + x (hi_k - lo_k)) for x = 2..K-1:
//     K products instead of K + 1, i.e. 12 Montgomery products per 4 entries instead of 14 for K = 3.  The host rebuilds
//     g_j(0..K) from them and the claim with exact field arithmetic (engine.cu: g4_rebuild_evals) -- the same field
//     elements the reference computes, so messages and transcript bytes are unchanged.
//  2. Leaner carry chains.  In the even/odd CIOS of mont32.cuh every 4-product chain ended with two carry adds into
//     words 8 and 9 of its accumulator window; an accumulator never exceeds 2^259 relative to its window base (it
//     receives less than 2^257 per row and loses 64 bits every two rows), so word 9 never changes and its add is
//     dropped: 32 instructions fewer per product.
//  3. Lazy sums.  The last factor of every message product is multiplied WITHOUT the final conditional subtraction and
//     added into a 288-bit integer accumulator (9 instructions instead of 17 + 25); the accumulators are reduced once
//     per thread (hi * 2^256 = hi * R, one Montgomery product by R^2).  The fold's difference t1 - t0 is formed as
//     t1 - t0 + p in (0, 2p) without a select: with the canonical challenge as the other operand the product stays
//     below 2p and the usual single subtraction makes it canonical.
//
// A fourth generation (second half of the file: k_fold_round_g4w, k_round_evals_g4w; the default) goes further once the
// multiplier pipe was measured to be the bound: unreduced last products in 544-bit accumulators, folds through a table of the
// challenge, a variant for moduli that are 1 modulo 2^32 -- 1018 instead of 1537 wide multiply-adds per 4 entries (K = 3).
//
// Everything stored to HBM is canonical, as everywhere else.
#pragma once
#include <cstdint>

#include "g4_types.hpp"
#include "kernels.cuh"

namespace scb {
namespace g4 {

// (t0,t1) += x0*b, (t2,t3) += x1*b, (t4,t5) += x2*b, (t6,t7) += x3*b, carry into t8.  ptxas fuses each
// mad.lo.cc / madc.hi.cc pair into one IMAD.WIDE.U32.X on the aligned register pair.
__device__ __forceinline__ void chain(uint32_t* t, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    asm("mad.lo.cc.u32   %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32  %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, %8, 0;\n\t"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
}
// the same chain started by the carry of `e0 += orphan` (the word that fell out of the other accumulator when the
// window moved on by one word)
__device__ __forceinline__ void chain_fix(uint32_t& e0, uint32_t orphan, uint32_t* t, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    asm("add.cc.u32      %9, %9, %15;\n\t"
        "madc.lo.cc.u32  %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32  %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32  %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32  %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32  %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32  %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32  %6, %13, %14, %6;\n\t"
        "madc.hi.cc.u32  %7, %13, %14, %7;\n\t"
        "addc.u32        %8, %8, 0;\n\t"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(e0)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b), "r"(orphan));
}

struct W8 {
    uint32_t w[8];
};
struct W9 {  // 288-bit integer: lazy sum of unreduced products
    uint32_t w[9];
};

__device__ __forceinline__ W8 load8(const uint64_t* l) {
    W8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.w[2 * i] = (uint32_t)l[i];
        r.w[2 * i + 1] = (uint32_t)(l[i] >> 32);
    }
    return r;
}
__device__ __forceinline__ void store8(const W8& a, uint64_t* l) {
#pragma unroll
    for (int i = 0; i < 4; ++i) l[i] = (uint64_t)a.w[2 * i] | ((uint64_t)a.w[2 * i + 1] << 32);
}

// the second reduction chain of a row when p = 1 (mod 2^32): word 0 receives m * 1 as a plain add (E0 + m = 0 mod 2^32,
// only its carry matters), the other three products as in chain().  Two asm blocks, and the carry re-enters the second
// through `add.cc c, 0xffffffff` -- the shape of chain_fix -- because ptxas only fuses mad.lo.cc / madc.hi.cc pairs into
// IMAD.WIDE.U32.X when the chain starts that way (with add.cc, addc.cc in front it emits IMAD.X + IMAD.HI.U32.X pairs).
__device__ __forceinline__ void chain_p0one(uint32_t* t, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t m) {
    uint32_t c;
    asm("add.cc.u32      %0, %0, %3;\n\t"
        "addc.cc.u32     %1, %1, 0;\n\t"
        "addc.u32        %2, 0, 0;\n\t"
        : "+r"(t[0]), "+r"(t[1]), "=r"(c)
        : "r"(m));
    asm("add.cc.u32      %7, %7, 0xffffffff;\n\t"
        "madc.lo.cc.u32  %0, %8, %11, %0;\n\t"
        "madc.hi.cc.u32  %1, %8, %11, %1;\n\t"
        "madc.lo.cc.u32  %2, %9, %11, %2;\n\t"
        "madc.hi.cc.u32  %3, %9, %11, %3;\n\t"
        "madc.lo.cc.u32  %4, %10, %11, %4;\n\t"
        "madc.hi.cc.u32  %5, %10, %11, %5;\n\t"
        "addc.u32        %6, %6, 0;\n\t"
        : "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(c)
        : "r"(x1), "r"(x2), "r"(x3), "r"(m));
}

// P0ONE: the modulus is 1 modulo 2^32 (every field whose two-adicity is at least 32 -- BLS12-381 Fr is one): then
// n0 = -p^-1 = -1 (mod 2^32), so m = -E[0] needs no multiplication and m * p[0] = m is an addition: 120 instead of 128
// wide multiply-adds per product.  The host picks the variant from the modulus (g4.cu).
template <bool P0ONE = false>
struct ArithT {
    uint32_t p[8];
    uint32_t n0;

    __device__ __forceinline__ explicit ArithT(const FieldDesc& f) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            p[2 * i] = (uint32_t)f.p[i];
            p[2 * i + 1] = (uint32_t)(f.p[i] >> 32);
        }
        n0 = (uint32_t)f.inv;
    }
    // a * b * 2^-256 as an UNREDUCED 257-bit value (lo, top) < a*b/2^256 + p: below 2p whenever one factor is below p
    // and the other below 2p.  Even/odd two-accumulator CIOS of mont32.cuh without the dead word-9 carry adds.
    __device__ __forceinline__ void mul_raw(uint32_t (&lo)[8], uint32_t& top, const W8& a, const W8& b) const {
        uint32_t A0[20], A1[20], sink = 0;
#pragma unroll
        for (int i = 0; i < 20; ++i) A0[i] = A1[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t* E = (i & 1) ? A1 + (i - 1) : A0 + i;
            uint32_t* O = (i & 1) ? A0 + (i + 1) : A1 + i;
            if (i == 0) {
                chain(O, a.w[1], a.w[3], a.w[5], a.w[7], b.w[i]);
            } else {
                const uint32_t orphan = (i & 1) ? A0[i] : A1[i - 1];  // previous E[1]
                chain_fix(E[0], orphan, O, a.w[1], a.w[3], a.w[5], a.w[7], b.w[i]);
            }
            chain(E, a.w[0], a.w[2], a.w[4], a.w[6], b.w[i]);
            if constexpr (P0ONE) {
                const uint32_t m = E[0] * n0;  // n0 = 0xffffffff; written as the product because ptxas stops fusing the
                                               // mad.lo / mad.hi pairs below into IMAD.WIDE when m is a plain negation
                chain(O, p[1], p[3], p[5], p[7], m);
                chain_p0one(E, p[2], p[4], p[6], m);  // E[0] becomes 0
            } else {
                const uint32_t m = E[0] * n0;
                chain(O, p[1], p[3], p[5], p[7], m);
                chain(E, p[0], p[2], p[4], p[6], m);  // E[0] becomes 0
                // E[0] is dead from here on; OR-ing it into a sink keeps the low half of its product alive, so ptxas emits
                // ONE IMAD.WIDE for the pair instead of IMAD + IMAD.HI (the carry is all that is needed)
                sink |= E[0];
            }
        }
        A0[16] |= sink;  // always zero
        // window after row 7: E = A0 + 8, O = A1 + 8, orphan = A1[7];  t = orphan + E + (O << 32)
        asm("add.cc.u32  %0, %9, %18;\n\t"
            "addc.cc.u32 %1, %10, %19;\n\t"
            "addc.cc.u32 %2, %11, %20;\n\t"
            "addc.cc.u32 %3, %12, %21;\n\t"
            "addc.cc.u32 %4, %13, %22;\n\t"
            "addc.cc.u32 %5, %14, %23;\n\t"
            "addc.cc.u32 %6, %15, %24;\n\t"
            "addc.cc.u32 %7, %16, %25;\n\t"
            "addc.u32    %8, %17, %26;\n\t"
            : "=r"(lo[0]), "=r"(lo[1]), "=r"(lo[2]), "=r"(lo[3]), "=r"(lo[4]), "=r"(lo[5]), "=r"(lo[6]), "=r"(lo[7]), "=r"(top)
            : "r"(A0[8]), "r"(A0[9]), "r"(A0[10]), "r"(A0[11]), "r"(A0[12]), "r"(A0[13]), "r"(A0[14]), "r"(A0[15]), "r"(A0[16]),
              "r"(A1[7]), "r"(A1[8]), "r"(A1[9]), "r"(A1[10]), "r"(A1[11]), "r"(A1[12]), "r"(A1[13]), "r"(A1[14]), "r"(A1[15]));
    }
    // a * b as a plain 512-bit integer (no reduction): 64 wide multiply-adds in the same even/odd two-accumulator layout
    // (A0[k] holds word k, A1[k] word k + 1; a row's chain ends in a word no earlier row has touched, so its carry-out
    // word cannot overflow), then one 16-word add joins the two.
    __device__ __forceinline__ void mul_wide(uint32_t (&t)[16], const W8& a, const W8& b) const {
        uint32_t A0[18], A1[18];
#pragma unroll
        for (int i = 0; i < 18; ++i) A0[i] = A1[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t* E = (i & 1) ? A1 + (i - 1) : A0 + i;
            uint32_t* O = (i & 1) ? A0 + (i + 1) : A1 + i;
            chain(O, a.w[1], a.w[3], a.w[5], a.w[7], b.w[i]);
            chain(E, a.w[0], a.w[2], a.w[4], a.w[6], b.w[i]);
        }
        t[0] = A0[0];
        asm("add.cc.u32  %0, %15, %30;\n\t"
            "addc.cc.u32 %1, %16, %31;\n\t"
            "addc.cc.u32 %2, %17, %32;\n\t"
            "addc.cc.u32 %3, %18, %33;\n\t"
            "addc.cc.u32 %4, %19, %34;\n\t"
            "addc.cc.u32 %5, %20, %35;\n\t"
            "addc.cc.u32 %6, %21, %36;\n\t"
            "addc.cc.u32 %7, %22, %37;\n\t"
            "addc.cc.u32 %8, %23, %38;\n\t"
            "addc.cc.u32 %9, %24, %39;\n\t"
            "addc.cc.u32 %10, %25, %40;\n\t"
            "addc.cc.u32 %11, %26, %41;\n\t"
            "addc.cc.u32 %12, %27, %42;\n\t"
            "addc.cc.u32 %13, %28, %43;\n\t"
            "addc.u32    %14, %29, %44;\n\t"
            : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]),
              "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
            : "r"(A0[1]), "r"(A0[2]), "r"(A0[3]), "r"(A0[4]), "r"(A0[5]), "r"(A0[6]), "r"(A0[7]), "r"(A0[8]), "r"(A0[9]), "r"(A0[10]), "r"(A0[11]),
              "r"(A0[12]), "r"(A0[13]), "r"(A0[14]), "r"(A0[15]),
              "r"(A1[0]), "r"(A1[1]), "r"(A1[2]), "r"(A1[3]), "r"(A1[4]), "r"(A1[5]), "r"(A1[6]), "r"(A1[7]), "r"(A1[8]), "r"(A1[9]), "r"(A1[10]),
              "r"(A1[11]), "r"(A1[12]), "r"(A1[13]), "r"(A1[14]));
    }
    // r * d * 2^-0 for the FIXED r of a fold pass, d any 256-bit integer congruent to the Montgomery-form difference:
    // d = sum_i d_i 2^(32 i), so r d = sum_i d_i (r 2^(32 i)); with the table t[i] = r 2^(32 i + 64) mod p the eight
    // rows d_i * t[i] all land on the SAME words (64 multiply-adds, a value below 2^35 p), and two Montgomery word
    // steps (16, or 14 when p = 1 mod 2^32) divide the 2^64 out again: 80 / 78 multiply-adds instead of 128 / 120, result
    // below p (1 + 2^-29), unreduced (lo, top).  Same even/odd layout as mul_raw: A0[k] holds word k, A1[k] word k + 1.
    __device__ __forceinline__ void mul_fixed_raw(uint32_t (&lo)[8], uint32_t& top, const FoldTab& ft, const W8& d) const {
        uint32_t A0[12], A1[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) A0[i] = A1[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            chain(A1, ft.t[i][1], ft.t[i][3], ft.t[i][5], ft.t[i][7], d.w[i]);
            chain(A0, ft.t[i][0], ft.t[i][2], ft.t[i][4], ft.t[i][6], d.w[i]);
        }
        {  // word 0
            const uint32_t m = A0[0] * n0;
            chain(A1, p[1], p[3], p[5], p[7], m);
            if constexpr (P0ONE) chain_p0one(A0, p[2], p[4], p[6], m);
            else chain(A0, p[0], p[2], p[4], p[6], m);
        }
        {  // word 1 = A0[1] + A1[0]
            const uint32_t m = (A0[1] + A1[0]) * n0;
            chain(A0 + 2, p[1], p[3], p[5], p[7], m);
            if constexpr (P0ONE) chain_p0one(A1, p[2], p[4], p[6], m);
            else chain(A1, p[0], p[2], p[4], p[6], m);
        }
        uint32_t sink;  // word 1 is now 0 (mod 2^32); its carry starts the final add
        asm("add.cc.u32  %9, %10, %11;\n\t"
            "addc.cc.u32 %0, %12, %21;\n\t"
            "addc.cc.u32 %1, %13, %22;\n\t"
            "addc.cc.u32 %2, %14, %23;\n\t"
            "addc.cc.u32 %3, %15, %24;\n\t"
            "addc.cc.u32 %4, %16, %25;\n\t"
            "addc.cc.u32 %5, %17, %26;\n\t"
            "addc.cc.u32 %6, %18, %27;\n\t"
            "addc.cc.u32 %7, %19, %28;\n\t"
            "addc.u32    %8, %20, 0;\n\t"
            : "=r"(lo[0]), "=r"(lo[1]), "=r"(lo[2]), "=r"(lo[3]), "=r"(lo[4]), "=r"(lo[5]), "=r"(lo[6]), "=r"(lo[7]), "=r"(top), "=r"(sink)
            : "r"(A0[1]), "r"(A1[0]),
              "r"(A0[2]), "r"(A0[3]), "r"(A0[4]), "r"(A0[5]), "r"(A0[6]), "r"(A0[7]), "r"(A0[8]), "r"(A0[9]), "r"(A0[10]),
              "r"(A1[1]), "r"(A1[2]), "r"(A1[3]), "r"(A1[4]), "r"(A1[5]), "r"(A1[6]), "r"(A1[7]), "r"(A1[8]));
    }
    // t0 + r * (t1 - t0) with the pass's table, canonical
    __device__ __forceinline__ W8 fold_fixed(const W8& t0, const W8& t1, const FoldTab& ft) const {
        uint32_t lo[8], top;
        mul_fixed_raw(lo, top, ft, diff_lazy(t1, t0));
        return add(t0, reduce_once(lo, top));
    }
    // (lo, top) < 2p  ->  canonical
    __device__ __forceinline__ W8 reduce_once(const uint32_t (&lo)[8], uint32_t top) const {
        uint32_t d[8];
        const uint32_t borrow = sub8(d, lo, p);
        const bool use_d = top != 0 || borrow == 0;
        W8 r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.w[i] = use_d ? d[i] : lo[i];
        return r;
    }
    __device__ __forceinline__ W8 mul(const W8& a, const W8& b) const {
        uint32_t lo[8], top;
        mul_raw(lo, top, a, b);
        return reduce_once(lo, top);
    }
    __device__ __forceinline__ W8 add(const W8& a, const W8& b) const {
        uint32_t s[8], d[8];
        const uint32_t carry = add8(s, a.w, b.w);
        const uint32_t borrow = sub8(d, s, p);
        const bool use_d = carry != 0 || borrow == 0;
        W8 r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.w[i] = use_d ? d[i] : s[i];
        return r;
    }
    __device__ __forceinline__ W8 sub(const W8& a, const W8& b) const {
        uint32_t d[8], e[8];
        const uint32_t borrow = sub8(d, a.w, b.w);
        add8(e, d, p);
        W8 r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.w[i] = borrow ? e[i] : d[i];
        return r;
    }
    // a - b + p in (0, 2p) for canonical a, b: no select.  2p < 2^256 needs bits(p) <= 255 (host checks).
    __device__ __forceinline__ W8 diff_lazy(const W8& a, const W8& b) const {
        uint32_t s[8];
        W8 r;
        add8(s, a.w, p);
        sub8(r.w, s, b.w);
        return r;
    }
    // t0 + r * (t1 - t0), canonical
    __device__ __forceinline__ W8 fold(const W8& t0, const W8& t1, const W8& r) const { return add(t0, mul(diff_lazy(t1, t0), r)); }
};
using Arith = ArithT<false>;

__device__ __forceinline__ void acc_zero(W9& a) {
#pragma unroll
    for (int i = 0; i < 9; ++i) a.w[i] = 0;
}
__device__ __forceinline__ void acc_add(W9& a, const uint32_t (&lo)[8], uint32_t top) {
    asm("add.cc.u32  %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32    %8, %8, %17;\n\t"
        : "+r"(a.w[0]), "+r"(a.w[1]), "+r"(a.w[2]), "+r"(a.w[3]), "+r"(a.w[4]), "+r"(a.w[5]), "+r"(a.w[6]), "+r"(a.w[7]), "+r"(a.w[8])
        : "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(top));
}

// Number of sums the pass accumulates for K tables: S_0, S_inf (K >= 2), S_2 .. S_{K-1}.
__host__ __device__ constexpr int n_sums(int K) { return K; }

// One thread-iteration handles 4 adjacent entries of every table (a "quad"): two folded entries per table = one
// hypercube pair of the next round.  Sums are written as n_sums(K) canonical elements (order: S_0, S_inf, S_2, ...).
template <int K, int MINB = 2>
__global__ void __launch_bounds__(kThreads, MINB)
    k_fold_round_g4(FieldDesc f, TabsIn<K> in, TabsOut<K> outp, ElemArg rarg, uint64_t n_quads, uint64_t* partials, unsigned int* ticket,
                    uint64_t* out, PeerArg peer) {
    constexpr int NS = n_sums(K);
    const Arith ar(f);
    const W8 r = load8(rarg.w);
    W9 acc[NS];
#pragma unroll
    for (int x = 0; x < NS; ++x) acc_zero(acc[x]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += stride) {
        W8 prod[NS];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint64_t w[16];
            ld_words<16>(in.p[k] + i * 16, w);
            const W8 u0 = ar.fold(load8(w), load8(w + 4), r);
            const W8 u1 = ar.fold(load8(w + 8), load8(w + 12), r);
            uint64_t o[8];
            store8(u0, o);
            store8(u1, o + 4);
            st_words<8>(outp.p[k] + i * 8, o);
            // factors of this table at the points 0, inf, 2, 3, ...: lo, hi - lo, hi + (hi - lo), ...
            W8 fac[NS];
            fac[0] = u0;
            if constexpr (K >= 2) {
                fac[1] = ar.sub(u1, u0);
#pragma unroll
                for (int x = 2; x < NS; ++x) fac[x] = ar.add(x == 2 ? u1 : fac[x - 1], fac[1]);
            }
#pragma unroll
            for (int x = 0; x < NS; ++x) {
                if (k == 0 && K > 1) {
                    prod[x] = fac[x];
                } else if (k < K - 1) {
                    prod[x] = ar.mul(prod[x], fac[x]);
                } else {  // last factor: unreduced product straight into the 288-bit sum
                    uint32_t lo[8], top;
                    if constexpr (K == 1) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) lo[q] = fac[x].w[q];
                        top = 0;
                    } else {
                        ar.mul_raw(lo, top, prod[x], fac[x]);
                    }
                    acc_add(acc[x], lo, top);
                }
            }
        }
    }
    // 288-bit sums -> canonical elements: value = hi * 2^256 + lo256, 2^256 = R (mod p), hi * R = montmul(hi, R^2)
    const PolGN<4> A(f);
    typename PolGN<4>::Acc fin[NS];
#pragma unroll
    for (int x = 0; x < NS; ++x) {
        // lo256 is an arbitrary 256-bit integer (many multiples of p for small moduli):
        // lo256 mod p = montmul(montmul(lo256, R^2), 1); mul_raw only needs its operands below 2^256
        W8 lo, hi, one_int;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            lo.w[q] = acc[x].w[q];
            hi.w[q] = 0;
            one_int.w[q] = 0;
        }
        hi.w[0] = acc[x].w[8];
        one_int.w[0] = 1;
        const W8 r2 = load8(f.r2);
        const W8 hr = ar.mul(hi, r2);
        const W8 s = ar.add(ar.mul(ar.mul(lo, r2), one_int), hr);
        uint64_t l[4];
        store8(s, l);
        fin[x] = A.from_words(l);
    }
    grid_reduce_finish<PolGN<4>, NS>(A, fin, partials, ticket, out, 0, &peer);
}

// ------------------------------------------------------------------------------------------------ wide accumulators (r2b)
// Fourth generation.  Measured on B200 (scripts/mont29_bench.cu, profiles/r02_mont29.md): a 32 x 32 -> 64-bit multiply-add
// (IMAD.WIDE.U32, with or without carry) issues once per 4 cycles per SM sub-partition on the fmaheavy pipe, and a bare
// product loop keeps that pipe 91-95 % busy (these kernels: 70-76 %) -- so the way to be faster is to issue fewer of them:
//   * the LAST product of every message point is not reduced at all: its 512-bit integer value prod * fac goes into a
//     544-bit accumulator (64 wide multiply-adds instead of 128), and each thread does ONE Montgomery reduction per
//     accumulator at the end (REDC is linear: sum REDC(x_i) = REDC(sum x_i) mod p);
//   * moduli with p = 1 (mod 2^32) skip the n0 multiplication and the p[0] column of every reduction (ArithT<true>);
//   * the fold's multiplier r is the same for the whole pass: with a host-built table of r 2^(32 i + 64) mod p the
//     product r d is eight rows on the same words plus two reduction steps (mul_fixed_raw): 80 / 78 instead of 128 / 120.
// The accumulators (17 words each) would not fit the 128-register budget of two resident CTAs per SM, so they live in
// shared memory: 5 x 128-bit per thread and sum, 80-byte stride (bank-conflict-free for 128-bit accesses), read, added to
// and written back once per product.  K >= 2 (K = 1 has no product to defer; the host keeps k_fold_round_g4 for it).
constexpr int kWaccQuads = 5;  // 128-bit words per accumulator slot (17 of the 20 32-bit words used)
__host__ __device__ constexpr size_t wacc_smem_bytes(int n_sums) { return (size_t)n_sums * kWaccQuads * 16 * kThreads; }

__device__ __forceinline__ uint4* wacc_slot(uint4* base, int x) { return base + ((size_t)x * kThreads + threadIdx.x) * kWaccQuads; }
__device__ __forceinline__ void wacc_zero(uint4* slot) {
#pragma unroll
    for (int q = 0; q < kWaccQuads; ++q) slot[q] = make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void wacc_add(uint4* slot, const uint32_t (&t)[16]) {
    uint4 a0 = slot[0], a1 = slot[1], a2 = slot[2], a3 = slot[3];
    uint32_t a16 = slot[4].x;
    asm("add.cc.u32  %0, %0, %17;\n\t"
        "addc.cc.u32 %1, %1, %18;\n\t"
        "addc.cc.u32 %2, %2, %19;\n\t"
        "addc.cc.u32 %3, %3, %20;\n\t"
        "addc.cc.u32 %4, %4, %21;\n\t"
        "addc.cc.u32 %5, %5, %22;\n\t"
        "addc.cc.u32 %6, %6, %23;\n\t"
        "addc.cc.u32 %7, %7, %24;\n\t"
        "addc.cc.u32 %8, %8, %25;\n\t"
        "addc.cc.u32 %9, %9, %26;\n\t"
        "addc.cc.u32 %10, %10, %27;\n\t"
        "addc.cc.u32 %11, %11, %28;\n\t"
        "addc.cc.u32 %12, %12, %29;\n\t"
        "addc.cc.u32 %13, %13, %30;\n\t"
        "addc.cc.u32 %14, %14, %31;\n\t"
        "addc.cc.u32 %15, %15, %32;\n\t"
        "addc.u32    %16, %16, 0;\n\t"
        : "+r"(a0.x), "+r"(a0.y), "+r"(a0.z), "+r"(a0.w), "+r"(a1.x), "+r"(a1.y), "+r"(a1.z), "+r"(a1.w), "+r"(a2.x), "+r"(a2.y), "+r"(a2.z),
          "+r"(a2.w), "+r"(a3.x), "+r"(a3.y), "+r"(a3.z), "+r"(a3.w), "+r"(a16)
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]),
          "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]));
    slot[0] = a0;
    slot[1] = a1;
    slot[2] = a2;
    slot[3] = a3;
    slot[4].x = a16;
}
// T = C0 + C1 2^256 + C2 2^512 (sum of products of Montgomery-form factors)  ->  REDC(T) = T / R mod p, canonical:
// C0 / R = montmul(C0, 1),  C1 2^256 / R = C1 (an arbitrary 256-bit integer: montmul(montmul(C1, R^2), 1) brings it below p),
// C2 2^512 / R = C2 R = montmul(C2, R^2).  Once per thread and accumulator.
template <bool P0ONE>
__device__ __forceinline__ W8 wacc_reduce(const ArithT<P0ONE>& ar, const FieldDesc& f, const uint4* slot) {
    W8 c0, c1, c2, one_int;
    const uint4 a0 = slot[0], a1 = slot[1], a2 = slot[2], a3 = slot[3];
    c0.w[0] = a0.x, c0.w[1] = a0.y, c0.w[2] = a0.z, c0.w[3] = a0.w, c0.w[4] = a1.x, c0.w[5] = a1.y, c0.w[6] = a1.z, c0.w[7] = a1.w;
    c1.w[0] = a2.x, c1.w[1] = a2.y, c1.w[2] = a2.z, c1.w[3] = a2.w, c1.w[4] = a3.x, c1.w[5] = a3.y, c1.w[6] = a3.z, c1.w[7] = a3.w;
#pragma unroll
    for (int q = 0; q < 8; ++q) c2.w[q] = one_int.w[q] = 0;
    c2.w[0] = slot[4].x;
    one_int.w[0] = 1;
    const W8 r2 = load8(f.r2);
    const W8 x0 = ar.mul(c0, one_int);
    const W8 x1 = ar.mul(ar.mul(c1, r2), one_int);
    const W8 x2 = ar.mul(c2, r2);
    return ar.add(ar.add(x0, x1), x2);
}

// Fused fold + message with a claim (as k_fold_round_g4): sums S_0, S_inf, S_2 .. S_{K-1}, canonical.
template <int K, bool P0ONE, int MINB = 2>
__global__ void __launch_bounds__(kThreads, MINB)
    k_fold_round_g4w(FieldDesc f, TabsIn<K> in, TabsOut<K> outp, FoldTab ft, uint64_t n_quads, uint64_t* partials, unsigned int* ticket,
                     uint64_t* out, PeerArg peer) {
    static_assert(K >= 2, "K = 1 has no product to defer");
    constexpr int NS = n_sums(K);
    extern __shared__ uint4 wacc[];
    const ArithT<P0ONE> ar(f);
#pragma unroll
    for (int x = 0; x < NS; ++x) wacc_zero(wacc_slot(wacc, x));
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += stride) {
        W8 prod[NS];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint64_t w[16];
            ld_words<16>(in.p[k] + i * 16, w);
            const W8 u0 = ar.fold_fixed(load8(w), load8(w + 4), ft);  // 80 / 78 multiply-adds: the multiplier is the pass's r
            const W8 u1 = ar.fold_fixed(load8(w + 8), load8(w + 12), ft);
            uint64_t o[8];
            store8(u0, o);
            store8(u1, o + 4);
            st_words<8>(outp.p[k] + i * 8, o);
            W8 fac[NS];
            fac[0] = u0;
            fac[1] = ar.sub(u1, u0);
#pragma unroll
            for (int x = 2; x < NS; ++x) fac[x] = ar.add(x == 2 ? u1 : fac[x - 1], fac[1]);
#pragma unroll
            for (int x = 0; x < NS; ++x) {
                if (k == 0) {
                    prod[x] = fac[x];
                } else if (k < K - 1) {
                    prod[x] = ar.mul(prod[x], fac[x]);
                } else {  // last factor: the plain 512-bit product into the wide accumulator
                    uint32_t t[16];
                    ar.mul_wide(t, prod[x], fac[x]);
                    wacc_add(wacc_slot(wacc, x), t);
                }
            }
        }
    }
    const PolGN<4> A(f);
    typename PolGN<4>::Acc fin[NS];
#pragma unroll
    for (int x = 0; x < NS; ++x) {
        const W8 s = wacc_reduce<P0ONE>(ar, f, wacc_slot(wacc, x));
        uint64_t l[4];
        store8(s, l);
        fin[x] = A.from_words(l);
    }
    grid_reduce_finish<PolGN<4>, NS>(A, fin, partials, ticket, out, 0, &peer);
}

// Round-0 message (Prover::new's pass: sum-check-protocol/src/lib.rs:88-97 with G::to_univariate,
// matrix-multiplication/src/lib.rs:110-122, generalised to K tables) in the same arithmetic.  No claim exists yet, so
// X = 1 is summed too: K + 1 sums in the order S_0, S_inf, S_2 .. S_{K-1}, S_1 (the host rebuilds g(0..K), engine.cu).
// One hypercube pair of every table per thread-iteration; 64 K + 64 instead of 128 K wide multiply-adds per point, and for
// K = 3 the first-level product at X = 2 comes from the other three by additions (the product of two linear factors is
// quadratic): 3 reduced + 4 unreduced products per pair.
template <int K, bool P0ONE, int MINB = 2>
__global__ void __launch_bounds__(kThreads, MINB)
    k_round_evals_g4w(FieldDesc f, TabsIn<K> in, uint64_t n_pairs, uint64_t* partials, unsigned int* ticket, uint64_t* out, PeerArg peer) {
    static_assert(K >= 2, "K = 1 has no product to defer");
    constexpr int NS = n_sums(K) + 1;
    extern __shared__ uint4 wacc[];
    const ArithT<P0ONE> ar(f);
#pragma unroll
    for (int x = 0; x < NS; ++x) wacc_zero(wacc_slot(wacc, x));
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
        W8 prod[NS];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint64_t w[8];
            ld_words<8>(in.p[k] + i * 8, w);
            const W8 lo = load8(w), hi = load8(w + 4);
            W8 fac[NS];
            fac[0] = lo;
            fac[NS - 1] = hi;
            fac[1] = ar.sub(hi, lo);
#pragma unroll
            for (int x = 2; x < NS - 1; ++x) fac[x] = ar.add(x == 2 ? hi : fac[x - 1], fac[1]);
#pragma unroll
            for (int x = 0; x < NS; ++x) {
                if (k == 0) {
                    prod[x] = fac[x];
                } else if (k < K - 1) {
                    // K = 3: q(X) = a(X) b(X) is quadratic, so its value at X = 2 follows from the three products at
                    // 0, inf and 1 by additions (below) -- one reduced product fewer per pair
                    if (!(K == 3 && x == 2)) prod[x] = ar.mul(prod[x], fac[x]);
                } else {
                    uint32_t t[16];
                    ar.mul_wide(t, prod[x], fac[x]);
                    wacc_add(wacc_slot(wacc, x), t);
                }
            }
            if constexpr (K == 3) {
                if (k == 1) {  // q(2) = q(0) + 2 q_1 + 4 q(inf) with q_1 = q(1) - q(0) - q(inf):  2 q(1) - q(0) + 2 q(inf)
                    const W8 s1 = ar.add(prod[3], prod[1]);  // q(1) + q(inf)   (sums: 0 -> X = 0, 1 -> inf, 2 -> X = 2, 3 -> X = 1)
                    prod[2] = ar.sub(ar.add(s1, s1), prod[0]);
                }
            }
        }
    }
    const PolGN<4> A(f);
    typename PolGN<4>::Acc fin[NS];
#pragma unroll
    for (int x = 0; x < NS; ++x) {
        const W8 s = wacc_reduce<P0ONE>(ar, f, wacc_slot(wacc, x));
        uint64_t l[4];
        store8(s, l);
        fin[x] = A.from_words(l);
    }
    grid_reduce_finish<PolGN<4>, NS>(A, fin, partials, ticket, out, 0, &peer);
}

}  // namespace g4
}  // namespace scb
