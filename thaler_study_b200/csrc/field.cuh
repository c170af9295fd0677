// field.cuh -- device-side Montgomery arithmetic over the reference's prime fields.
//
// Semantics replaced: [ARK] ark_ff::Fp<MontBackend<C,N>,N> add/sub/mul (SURVEY.md 8a row
// a11; used by every hot loop of /root/reference: matrix-multiplication/src/lib.rs:110-122,
// DenseMultilinearExtension::fix_variables called at :104-105, multilinear-extensions/src/lib.rs:9-23).
// In-memory format is ark's: N little-endian u64 limbs, Montgomery form with R = 2^(64N),
// value < p.  Everything a kernel stores back to HBM is canonical (fully reduced).
//
// Three arithmetic policies share one interface so kernels are written once:
//   PolSP    1 limb, p < 2^28 (all three reference moduli 5 / 389 / 1572869): 32-bit word-wise
//            Montgomery (5 IMAD-class instructions per product), lazy reductions, u64 accumulators.
//   PolG1    1 limb, any odd p < 2^64.
//   PolGN<N> N limbs (2..4), CIOS over 64-bit limbs with PTX carry chains.
#pragma once
#include <cstdint>

#include "fielddesc.hpp"
#include "mont32.cuh"

namespace scb {

// FieldDesc (the by-value kernel argument) lives in fielddesc.hpp so host-only translation units can use it.

// ------------------------------------------------------------------------------------------
// PolSP: small prime (p < 2^28).  El/Lz are 32-bit values; an El is canonical (< p), an Lz is a
// "lazy" value < 2^32.  mul() accepts one operand < 2^32 and one < 2^31 and returns a value <= p.
// ------------------------------------------------------------------------------------------
struct PolSP {
    static constexpr bool kLight = true;  // few registers per element (launch-bounds class)
    static constexpr int N = 1;   // u64 words per element in memory
    static constexpr int AW = 1;  // u64 words per accumulator
    using El = uint32_t;
    using Lz = uint32_t;
    using Acc = uint64_t;
    uint32_t p, ninv;  // ninv = -p^{-1} mod 2^32

    __device__ __forceinline__ explicit PolSP(const FieldDesc& f) : p((uint32_t)f.p[0]), ninv((uint32_t)f.inv) {}

    __device__ __forceinline__ El from_words(const uint64_t* w) const { return (uint32_t)w[0]; }
    __device__ __forceinline__ void to_words(El a, uint64_t* w) const { w[0] = a; }
    __device__ __forceinline__ El zero() const { return 0; }

    __device__ __forceinline__ El reduce_once(uint32_t a) const { return a >= p ? a - p : a; }
    // T * 2^-64 mod p for T < 2^63, result <= p  (two 32-bit Montgomery steps)
    __device__ __forceinline__ uint32_t redc(uint64_t T) const {
        uint32_t m0 = (uint32_t)T * ninv;
        uint64_t t1 = (T + (uint64_t)m0 * p) >> 32;
        uint32_t m1 = (uint32_t)t1 * ninv;
        return (uint32_t)((t1 + (uint64_t)m1 * p) >> 32);
    }
    __device__ __forceinline__ Lz lz_mul(Lz a, Lz b) const { return redc((uint64_t)a * b); }
    __device__ __forceinline__ El mul(El a, El b) const { return reduce_once(lz_mul(a, b)); }
    // One 32-bit Montgomery step: T * 2^-32 mod p for T < 2^63, result < T/2^32 + p.  The round kernels use it for
    // every product (3 multiply-class instructions instead of 5) and repair the missing powers of 2^-32 where that
    // is free: in the fold the factor is folded into the per-kernel constant (fold_const), in the K-fold message
    // products it is a constant per output sum, applied once by the finishing thread (msg_final).
    // Written as mad.wide.u32 (IMAD.WIDE with the 64-bit addend T): left to itself the compiler splits the sum into
    // IMAD.HI + carry fix-ups, and IMAD.HI issues at half the rate of IMAD.WIDE on sm_100 (profiles/r01_imad_peak.md).
    __device__ __forceinline__ uint32_t redc32(uint64_t T) const {
        const uint32_t m = (uint32_t)T * ninv;
        uint64_t s;
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(s) : "r"(m), "r"(p), "l"(T));
        return (uint32_t)(s >> 32);
    }
    using FoldC = uint32_t;
    __device__ __forceinline__ FoldC fold_const(El r) const { return reduce_once(redc32(r)); }  // r * 2^-32
    // t0 + r*(t1 - t0) with rc = fold_const(r): redc32(d * rc) = d * r * 2^-64 = the Montgomery product.
    // d < 2p, rc < p  =>  redc32 < p + 2p^2/2^32 < 1.125 p  =>  sum < 2.125 p: two conditional subtractions.
    __device__ __forceinline__ El fold_c(El t0, El t1, FoldC rc) const {
        return reduce_once(reduce_once(t0 + redc32((uint64_t)(t1 - t0 + p) * rc)));
    }
    __device__ __forceinline__ Lz msg_mul(Lz a, Lz b) const { return redc32((uint64_t)a * b); }  // a*b*2^-32, < 2^31 + p
    // sum of products of k Montgomery values each computed with (k-1) msg_mul steps: multiply by 2^(-32(k-1))
    __device__ __forceinline__ El msg_final(const Acc& a, int k) const {
        uint32_t x = (uint32_t)(a % p);
        for (int i = 1; i < k; ++i) x = reduce_once(redc32(x));
        return x;
    }
    __device__ __forceinline__ El add(El a, El b) const { return reduce_once(a + b); }
    __device__ __forceinline__ El sub(El a, El b) const { return a >= b ? a - b : a - b + p; }
    // t0 + r*(t1 - t0), canonical
    __device__ __forceinline__ El fold(El t0, El t1, El r) const {
        return reduce_once(t0 + lz_mul(t1 - t0 + p, r));
    }
    __device__ __forceinline__ Lz lz(El a) const { return a; }
    __device__ __forceinline__ Lz lz_diff(El a, El b) const { return a - b + p; }  // a-b (mod p), in (0, 2p)
    __device__ __forceinline__ Lz lz_add(Lz a, Lz b) const { return a + b; }       // caller keeps < 2^32
    __device__ __forceinline__ Lz lz_sum(Lz a, Lz b) const { return reduce_once(reduce_once(a + b)); }  // a,b <= p
    __device__ __forceinline__ void acc_zero(Acc& a) const { a = 0; }
    __device__ __forceinline__ void acc_add(Acc& a, Lz x) const { a += x; }
    __device__ __forceinline__ void acc_merge(Acc& a, const Acc& b) const { a += b; }
    __device__ __forceinline__ void acc_to_words(const Acc& a, uint64_t* w) const { w[0] = a; }
    __device__ __forceinline__ void acc_from_words(Acc& a, const uint64_t* w) const { a = w[0]; }
    __device__ __forceinline__ El acc_final(const Acc& a) const { return (uint32_t)(a % p); }
};

// ------------------------------------------------------------------------------------------
// PolG1: one 64-bit limb, any odd modulus.
// ------------------------------------------------------------------------------------------
struct PolG1 {
    static constexpr bool kLight = false;
    static constexpr int N = 1;
    static constexpr int AW = 1;
    using El = uint64_t;
    using Lz = uint64_t;
    using Acc = uint64_t;
    uint64_t p, inv;

    __device__ __forceinline__ explicit PolG1(const FieldDesc& f) : p(f.p[0]), inv(f.inv) {}

    __device__ __forceinline__ El from_words(const uint64_t* w) const { return w[0]; }
    __device__ __forceinline__ void to_words(El a, uint64_t* w) const { w[0] = a; }
    __device__ __forceinline__ El zero() const { return 0; }

    __device__ __forceinline__ El add(El a, El b) const {
        uint64_t s = a + b;
        return (s < a || s >= p) ? s - p : s;
    }
    __device__ __forceinline__ El sub(El a, El b) const { return a >= b ? a - b : a - b + p; }
    __device__ __forceinline__ El mul(El a, El b) const {
        uint64_t lo = a * b, hi = __umul64hi(a, b);
        uint64_t m = lo * inv;
        uint64_t mh = __umul64hi(m, p);
        // lo + low64(m*p) == 0 mod 2^64, carry out iff lo != 0
        uint64_t t = hi + mh;
        bool c = t < hi;
        uint64_t t2 = t + (lo != 0);
        c |= t2 < t;
        return (c || t2 >= p) ? t2 - p : t2;
    }
    __device__ __forceinline__ El fold(El t0, El t1, El r) const { return add(t0, mul(sub(t1, t0), r)); }
    using FoldC = El;
    __device__ __forceinline__ FoldC fold_const(const El& r) const { return r; }
    __device__ __forceinline__ El fold_c(const El& t0, const El& t1, const FoldC& rc) const { return fold(t0, t1, rc); }
    __device__ __forceinline__ Lz msg_mul(const Lz& a, const Lz& b) const { return lz_mul(a, b); }
    __device__ __forceinline__ El msg_final(const Acc& a, int) const { return acc_final(a); }
    __device__ __forceinline__ Lz lz(El a) const { return a; }
    __device__ __forceinline__ Lz lz_diff(El a, El b) const { return sub(a, b); }
    __device__ __forceinline__ Lz lz_add(Lz a, Lz b) const { return add(a, b); }
    __device__ __forceinline__ Lz lz_sum(Lz a, Lz b) const { return add(a, b); }
    __device__ __forceinline__ Lz lz_mul(Lz a, Lz b) const { return mul(a, b); }
    __device__ __forceinline__ void acc_zero(Acc& a) const { a = 0; }
    __device__ __forceinline__ void acc_add(Acc& a, Lz x) const { a = add(a, x); }
    __device__ __forceinline__ void acc_merge(Acc& a, const Acc& b) const { a = add(a, b); }
    __device__ __forceinline__ void acc_to_words(const Acc& a, uint64_t* w) const { w[0] = a; }
    __device__ __forceinline__ void acc_from_words(Acc& a, const uint64_t* w) const { a = w[0]; }
    __device__ __forceinline__ El acc_final(const Acc& a) const { return a; }
};

// ------------------------------------------------------------------------------------------
// PolGN<N>: N 64-bit limbs, CIOS Montgomery product.
// ------------------------------------------------------------------------------------------
template <int NL>
struct ElN {
    uint64_t l[NL];
};

// (c, t) = t + a*b + c
__device__ __forceinline__ void mac64(uint64_t& t, uint64_t a, uint64_t b, uint64_t& c) {
    uint64_t lo, hi;
    asm("{\n\t"
        ".reg .u64 l, h;\n\t"
        "mul.lo.u64 l, %2, %3;\n\t"
        "mul.hi.u64 h, %2, %3;\n\t"
        "add.cc.u64 l, l, %4;\n\t"
        "addc.u64 h, h, 0;\n\t"
        "add.cc.u64 %0, l, %5;\n\t"
        "addc.u64 %1, h, 0;\n\t"
        "}"
        : "=l"(lo), "=l"(hi)
        : "l"(a), "l"(b), "l"(t), "l"(c));
    t = lo;
    c = hi;
}

template <int NL>
struct PolGN {
    static constexpr bool kLight = false;
    static constexpr int N = NL;
    static constexpr int AW = NL;
    using El = ElN<NL>;
    using Lz = ElN<NL>;
    using Acc = ElN<NL>;
    uint64_t p[NL];
    uint64_t inv;

    Mont8x32 m32;  // 8 x 32-bit carry-chain arithmetic (used when NL == 4)
    __device__ __forceinline__ explicit PolGN(const FieldDesc& f) : inv(f.inv) {
#pragma unroll
        for (int i = 0; i < NL; ++i) p[i] = f.p[i];
        if constexpr (NL == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                m32.p[2 * i] = (uint32_t)f.p[i];
                m32.p[2 * i + 1] = (uint32_t)(f.p[i] >> 32);
            }
            m32.n0 = (uint32_t)f.inv;
        }
    }
    __device__ __forceinline__ static void split8(const El& a, uint32_t (&w)[8]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            w[2 * i] = (uint32_t)a.l[i];
            w[2 * i + 1] = (uint32_t)(a.l[i] >> 32);
        }
    }
    __device__ __forceinline__ static El join8(const uint32_t (&w)[8]) {
        El e;
#pragma unroll
        for (int i = 0; i < 4; ++i) e.l[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
        return e;
    }

    __device__ __forceinline__ El from_words(const uint64_t* w) const {
        El e;
#pragma unroll
        for (int i = 0; i < NL; ++i) e.l[i] = w[i];
        return e;
    }
    __device__ __forceinline__ void to_words(const El& a, uint64_t* w) const {
#pragma unroll
        for (int i = 0; i < NL; ++i) w[i] = a.l[i];
    }
    __device__ __forceinline__ El zero() const {
        El e;
#pragma unroll
        for (int i = 0; i < NL; ++i) e.l[i] = 0;
        return e;
    }
    // a >= p ?
    __device__ __forceinline__ bool geq_p(const uint64_t* a) const {
        bool ge = true;  // equal so far => ge
#pragma unroll
        for (int i = 0; i < NL; ++i) {  // low to high: higher limbs override
            if (a[i] > p[i]) ge = true;
            else if (a[i] < p[i]) ge = false;
        }
        return ge;
    }
    __device__ __forceinline__ void sub_p(uint64_t* a) const {
        uint64_t borrow = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            uint64_t d = a[i] - p[i];
            uint64_t b1 = a[i] < p[i];
            uint64_t d2 = d - borrow;
            uint64_t b2 = d < borrow;
            a[i] = d2;
            borrow = b1 | b2;
        }
    }
    __device__ __forceinline__ El add(const El& a, const El& b) const {
        if constexpr (NL == 4) {
            uint32_t x[8], y[8], r[8];
            split8(a, x);
            split8(b, y);
            m32.add(r, x, y);
            return join8(r);
        }
        El s;
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            uint64_t x = a.l[i] + b.l[i];
            uint64_t c1 = x < a.l[i];
            uint64_t y = x + c;
            uint64_t c2 = y < x;
            s.l[i] = y;
            c = c1 | c2;
        }
        if (c || geq_p(s.l)) sub_p(s.l);
        return s;
    }
    __device__ __forceinline__ El sub(const El& a, const El& b) const {
        if constexpr (NL == 4) {
            uint32_t x[8], y[8], r[8];
            split8(a, x);
            split8(b, y);
            m32.sub(r, x, y);
            return join8(r);
        }
        El d;
        uint64_t borrow = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            uint64_t x = a.l[i] - b.l[i];
            uint64_t b1 = a.l[i] < b.l[i];
            uint64_t y = x - borrow;
            uint64_t b2 = x < borrow;
            d.l[i] = y;
            borrow = b1 | b2;
        }
        if (borrow) {
            uint64_t c = 0;
#pragma unroll
            for (int i = 0; i < NL; ++i) {
                uint64_t x = d.l[i] + p[i];
                uint64_t c1 = x < d.l[i];
                uint64_t y = x + c;
                uint64_t c2 = y < x;
                d.l[i] = y;
                c = c1 | c2;
            }
        }
        return d;
    }
    __device__ __forceinline__ El mul(const El& a, const El& b) const {
        if constexpr (NL == 4) {
            uint32_t x[8], y[8], r[8];
            split8(a, x);
            split8(b, y);
            m32.mul(r, x, y);
            return join8(r);
        }
        uint64_t t[NL + 2];
#pragma unroll
        for (int i = 0; i < NL + 2; ++i) t[i] = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < NL; ++j) mac64(t[j], a.l[j], b.l[i], c);
            uint64_t s = t[NL] + c;
            t[NL + 1] = s < c;
            t[NL] = s;
            uint64_t m = t[0] * inv;
            c = 0;
            uint64_t dump = t[0];
            mac64(dump, m, p[0], c);
#pragma unroll
            for (int j = 1; j < NL; ++j) {
                uint64_t tj = t[j];
                mac64(tj, m, p[j], c);
                t[j - 1] = tj;
            }
            s = t[NL] + c;
            uint64_t c2 = s < c;
            t[NL - 1] = s;
            t[NL] = t[NL + 1] + c2;
        }
        El r;
#pragma unroll
        for (int i = 0; i < NL; ++i) r.l[i] = t[i];
        if (t[NL] || geq_p(r.l)) sub_p(r.l);
        return r;
    }
    __device__ __forceinline__ El fold(const El& t0, const El& t1, const El& r) const { return add(t0, mul(sub(t1, t0), r)); }
    using FoldC = El;
    __device__ __forceinline__ FoldC fold_const(const El& r) const { return r; }
    __device__ __forceinline__ El fold_c(const El& t0, const El& t1, const FoldC& rc) const { return fold(t0, t1, rc); }
    __device__ __forceinline__ Lz msg_mul(const Lz& a, const Lz& b) const { return lz_mul(a, b); }
    __device__ __forceinline__ El msg_final(const Acc& a, int) const { return acc_final(a); }
    __device__ __forceinline__ Lz lz(const El& a) const { return a; }
    __device__ __forceinline__ Lz lz_diff(const El& a, const El& b) const { return sub(a, b); }
    __device__ __forceinline__ Lz lz_add(const Lz& a, const Lz& b) const { return add(a, b); }
    __device__ __forceinline__ Lz lz_sum(const Lz& a, const Lz& b) const { return add(a, b); }
    __device__ __forceinline__ Lz lz_mul(const Lz& a, const Lz& b) const { return mul(a, b); }
    __device__ __forceinline__ void acc_zero(Acc& a) const { a = zero(); }
    __device__ __forceinline__ void acc_add(Acc& a, const Lz& x) const { a = add(a, x); }
    __device__ __forceinline__ void acc_merge(Acc& a, const Acc& b) const { a = add(a, b); }
    __device__ __forceinline__ void acc_to_words(const Acc& a, uint64_t* w) const { to_words(a, w); }
    __device__ __forceinline__ void acc_from_words(Acc& a, const uint64_t* w) const { a = from_words(w); }
    __device__ __forceinline__ El acc_final(const Acc& a) const { return a; }
};

// ------------------------------------------------------------------------------------------
// vectorised global memory access: W consecutive u64 words (W*8 bytes, aligned to W*8 up to 32)
// ------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void ld_words(const uint64_t* __restrict__ ptr, uint64_t* w) {
    if constexpr (W % 4 == 0) {
#pragma unroll
        for (int i = 0; i < W; i += 4)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(w[i]), "=l"(w[i + 1]), "=l"(w[i + 2]), "=l"(w[i + 3])
                         : "l"(ptr + i));
    } else if constexpr (W % 2 == 0) {
#pragma unroll
        for (int i = 0; i < W; i += 2)
            asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(w[i]), "=l"(w[i + 1]) : "l"(ptr + i));
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(w[i]) : "l"(ptr + i));
    }
}
template <int W>
__device__ __forceinline__ void st_words(uint64_t* __restrict__ ptr, const uint64_t* w) {
    if constexpr (W % 4 == 0) {
#pragma unroll
        for (int i = 0; i < W; i += 4)
            asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(ptr + i), "l"(w[i]), "l"(w[i + 1]),
                         "l"(w[i + 2]), "l"(w[i + 3])
                         : "memory");
    } else if constexpr (W % 2 == 0) {
#pragma unroll
        for (int i = 0; i < W; i += 2)
            asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1,%2};" ::"l"(ptr + i), "l"(w[i]), "l"(w[i + 1]) : "memory");
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" ::"l"(ptr + i), "l"(w[i]) : "memory");
    }
}

}  // namespace scb
