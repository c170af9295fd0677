// mont32.cuh -- 256-bit Montgomery arithmetic as 32-bit IMAD carry chains in registers (4 ark limbs = 8 words).
//
// [ARK] Fp<MontBackend<C,4>,4> add / sub / mul for fields such as BLS12-381 Fr
// (= ark_ed_on_bls12_381::Fq, /root/reference/Cargo.toml:20).  The B200 integer pipe is 32 bits wide, so the
// product is word-serial CIOS over 32-bit words: per word b[i] one row `t += a*b[i]` and one row `t += m*p`,
// each row two carry chains (even and odd columns: the 64-bit partial products of the even / odd a[j] occupy
// disjoint word pairs, so `mad.lo.cc / madc.hi.cc` can run straight through them) -- 16 multiply-adds and 3
// carry-adds per row, 2 x 8 rows per product (~310 instructions, 264 of them IMAD), against ~700 for the
// 64-bit-limb formulation nvcc lowers to 32-bit multiplies anyway.
#pragma once
#include <cstdint>

namespace scb {

// t[0..9] += a(8 words) * b   (t points into a longer fully-unrolled register array)
__device__ __forceinline__ void mad_row8(uint32_t* t, const uint32_t (&a)[8], uint32_t b) {
    asm("mad.lo.cc.u32   %0, %10, %18, %0;\n\t"
        "madc.hi.cc.u32  %1, %10, %18, %1;\n\t"
        "madc.lo.cc.u32  %2, %12, %18, %2;\n\t"
        "madc.hi.cc.u32  %3, %12, %18, %3;\n\t"
        "madc.lo.cc.u32  %4, %14, %18, %4;\n\t"
        "madc.hi.cc.u32  %5, %14, %18, %5;\n\t"
        "madc.lo.cc.u32  %6, %16, %18, %6;\n\t"
        "madc.hi.cc.u32  %7, %16, %18, %7;\n\t"
        "addc.cc.u32     %8, %8, 0;\n\t"
        "addc.u32        %9, %9, 0;\n\t"
        "mad.lo.cc.u32   %1, %11, %18, %1;\n\t"
        "madc.hi.cc.u32  %2, %11, %18, %2;\n\t"
        "madc.lo.cc.u32  %3, %13, %18, %3;\n\t"
        "madc.hi.cc.u32  %4, %13, %18, %4;\n\t"
        "madc.lo.cc.u32  %5, %15, %18, %5;\n\t"
        "madc.hi.cc.u32  %6, %15, %18, %6;\n\t"
        "madc.lo.cc.u32  %7, %17, %18, %7;\n\t"
        "madc.hi.cc.u32  %8, %17, %18, %8;\n\t"
        "addc.u32        %9, %9, 0;\n\t"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b));
}

// One carry chain over five aligned word pairs: (t0,t1) += x0*b, (t2,t3) += x1*b, (t4,t5) += x2*b, (t6,t7) += x3*b,
// carry into t8, t9.  ptxas fuses each mad.lo.cc / madc.hi.cc pair into one IMAD.WIDE.U32.X on the register pair.
__device__ __forceinline__ void chain8(uint32_t* t, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    asm("mad.lo.cc.u32   %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32  %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32  %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32  %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32  %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32  %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32  %6, %13, %14, %6;\n\t"
        "madc.hi.cc.u32  %7, %13, %14, %7;\n\t"
        "addc.cc.u32     %8, %8, 0;\n\t"
        "addc.u32        %9, %9, 0;\n\t"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
}
// Same chain, started by the carry of `e0 += orphan` (the word that fell out of the other accumulator when the
// window moved on by one word, see Mont8x32::mul).
__device__ __forceinline__ void chain8_fix(uint32_t& e0, uint32_t orphan, uint32_t* t, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3,
                                           uint32_t b) {
    asm("add.cc.u32      %10, %10, %16;\n\t"
        "madc.lo.cc.u32  %0, %11, %15, %0;\n\t"
        "madc.hi.cc.u32  %1, %11, %15, %1;\n\t"
        "madc.lo.cc.u32  %2, %12, %15, %2;\n\t"
        "madc.hi.cc.u32  %3, %12, %15, %3;\n\t"
        "madc.lo.cc.u32  %4, %13, %15, %4;\n\t"
        "madc.hi.cc.u32  %5, %13, %15, %5;\n\t"
        "madc.lo.cc.u32  %6, %14, %15, %6;\n\t"
        "madc.hi.cc.u32  %7, %14, %15, %7;\n\t"
        "addc.cc.u32     %8, %8, 0;\n\t"
        "addc.u32        %9, %9, 0;\n\t"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]), "+r"(e0)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b), "r"(orphan));
}

// d = a - b over 8 words; returns the borrow (0 / 1)
__device__ __forceinline__ uint32_t sub8(uint32_t (&d)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t borrow;
    asm("sub.cc.u32  %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;\n\t"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(borrow)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]),
          "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return borrow & 1u;  // subc of 0-0-borrow is 0 or 0xffffffff
}
// s = a + b over 8 words; returns the carry (0 / 1)
__device__ __forceinline__ uint32_t add8(uint32_t (&s)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t carry;
    asm("add.cc.u32  %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;\n\t"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(carry)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]),
          "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return carry;
}

struct Mont8x32 {
    uint32_t p[8];
    uint32_t n0;  // -p^{-1} mod 2^32

    // (a + b) mod p, inputs canonical
    __device__ __forceinline__ void add(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) const {
        uint32_t s[8], d[8];
        const uint32_t carry = add8(s, a, b);
        const uint32_t borrow = sub8(d, s, p);
        const bool use_d = carry != 0 || borrow == 0;  // s >= p
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = use_d ? d[i] : s[i];
    }
    // (a - b) mod p, inputs canonical
    __device__ __forceinline__ void sub(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) const {
        uint32_t d[8], e[8];
        const uint32_t borrow = sub8(d, a, b);
        add8(e, d, p);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = borrow ? e[i] : d[i];
    }
    // a * b * 2^-256 mod p, inputs canonical, output canonical.
    // Word-serial CIOS with TWO accumulators so that every 64-bit partial product lands on an aligned register pair
    // (one IMAD.WIDE.U32.X, no re-pairing moves): E holds the products of the even words of the multiplicand
    // (pairs at words 0-1, 2-3, ...), O those of the odd words (pairs at words 1-2, 3-4, ...); t = E + (O << 32).
    // After a row t[0] = E[0] = 0 and the window moves on by one word: O becomes the even accumulator as it is,
    // E (minus its two lowest words) becomes the odd one, and the single word that falls out, E[1], is added into
    // the new E[0] at the head of the next row's first carry chain.  Offsets are compile-time (full unroll).
    __device__ __forceinline__ void mul(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) const {
        uint32_t A0[22], A1[22];
#pragma unroll
        for (int i = 0; i < 22; ++i) A0[i] = A1[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t* E = (i & 1) ? A1 + (i - 1) : A0 + i;
            uint32_t* O = (i & 1) ? A0 + (i + 1) : A1 + i;
            if (i == 0) {
                chain8(O, a[1], a[3], a[5], a[7], b[i]);
            } else {
                const uint32_t orphan = (i & 1) ? A0[i] : A1[i - 1];  // previous E[1]
                chain8_fix(E[0], orphan, O, a[1], a[3], a[5], a[7], b[i]);
            }
            chain8(E, a[0], a[2], a[4], a[6], b[i]);
            const uint32_t m = E[0] * n0;
            chain8(O, p[1], p[3], p[5], p[7], m);
            chain8(E, p[0], p[2], p[4], p[6], m);  // E[0] becomes 0
        }
        // window after row 7: E = A0 + 8, O = A1 + 8, orphan = A1[7];  t = orphan + E + (O << 32)
        uint32_t lo[8], d[8], top;
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
        const uint32_t borrow = sub8(d, lo, p);
        const bool use_d = top != 0 || borrow == 0;  // t >= p  (t < 2p always)
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = use_d ? d[i] : lo[i];
    }
};

}  // namespace scb
