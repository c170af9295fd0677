// lazy29.hpp -- 4-limb (250..255-bit) field arithmetic in radix 2^29 with lazy carries, host- and device-compilable.
//
// Why: on B200 the multiply-add with a 64-bit addend and NO carry (IMAD.WIDE.U32) issues at full rate, while the form a
// 32-bit-limb carry chain needs (IMAD.WIDE.U32.X, predicate carry in and out) issues at half that rate
// (scripts/imad_chain.cu, profiles/r02_imad_chain.jsonl).  With nine 29-bit limbs a column of a 9 x 9 schoolbook product
// sums eighteen 58-bit products in one 64-bit register without ever overflowing, so the whole Montgomery product is
// 162 plain IMAD.WIDE + 9 IMAD and a handful of shifts and adds per row on the other (ALU) pipe -- against 128
// half-rate IMAD.WIDE.X plus selects and conditional subtractions in the 32-bit-limb form (mont32.cuh, g4.cuh).
//
// Representation: L9 = nine uint32 limbs, value = sum l[j] 2^(29 j).  "Normalised": every limb below 2^29 except the
// top one, which holds the rest.  Lazy values may have limbs up to ~2^31 (sums / differences formed limb by limb).
// Elements stay in ark's Montgomery form (R = 2^256) in memory; the product here divides by R' = 2^261, i.e. it returns
// x y 2^-261 = MontMul_256(x, y) * 2^-5.  The missing factors 2^5 are repaired where that is free: the fold multiplies
// by the per-kernel constant r * 2^5, and a message product of K factors carries 2^(-5 (K-1)), which the HOST multiplies
// back into the K sums the kernel returns (exact field arithmetic).
//
// Bounds (p < 2^255, R' = 2^261): mont(x, y) < x y / 2^261 + p, so inputs up to ~10p give outputs below 2.6p and
// nothing here needs a conditional subtraction; only what is stored to HBM is made canonical (32-bit word domain).
// Column accumulators: a slot receives at most 9 x (A B + 2^58) + carries where A, B bound the limbs of the two operands:
// A B <= 2^60.2 keeps it below 2^64 (one operand normalised, the other with limbs below 2^31.2; or both below 2^30.1).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define L29_HD __host__ __device__ __forceinline__
#else
#define L29_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define L29_UNROLL _Pragma("unroll")
#else
#define L29_UNROLL  // host passes: plain loops (g++ would warn about the unknown pragma)
#endif

namespace scb {
namespace l29 {

constexpr int NL = 9;
constexpr uint32_t M29 = (1u << 29) - 1;

struct L9 {
    uint32_t l[NL];
};

// Field constants in limb form (kernel argument; built on the host by make_desc)
struct Desc29 {
    uint32_t p[NL];    // p, normalised
    uint32_t kp[NL];   // k p (k = 2^(258 - bits(p)), so 2^257 <= k p < 2^258) in "big limb" form: kp[j] >= 2^29 - 1 for j < 8
                       // and kp[8] >= 2^24, so that a[j] + kp[j] - b[j] never goes negative for a normalised b below 2^256
    uint32_t n0;       // -p^-1 mod 2^29
    uint32_t k;        // the multiple
};

// x < 2^256 as eight 32-bit words -> normalised limbs
L29_HD L9 from_words(const uint32_t (&w)[8]) {
    L9 r;
L29_UNROLL
    for (int j = 0; j < NL; ++j) {
        const int bit = 29 * j, q = bit >> 5, s = bit & 31;
        uint32_t v = w[q] >> s;
        if (s > 3 && q + 1 < 8) v |= w[q + 1] << (32 - s);  // the limb crosses into the next word
        r.l[j] = j == NL - 1 ? v : (v & M29);
    }
    return r;
}
// Limbs (lazy allowed) -> eight 32-bit words + the overflow above bit 256 (`top`), carries resolved on the way: a running
// 64-bit accumulator takes limb j at its bit position; word e is final once every limb starting below bit 32 (e + 1) is in.
L29_HD void to_words(const L9& a, uint32_t (&w)[8], uint32_t& top) {
    uint64_t acc = 0;
    int emitted = 0;
L29_UNROLL
    for (int j = 0; j < NL; ++j) {
L29_UNROLL
        for (int rep = 0; rep < 2; ++rep) {
            if (emitted < 8 && 32 * (emitted + 1) <= 29 * j) {
                w[emitted++] = (uint32_t)acc;
                acc >>= 32;
            }
        }
        acc += (uint64_t)a.l[j] << (29 * j - 32 * emitted);
    }
L29_UNROLL
    for (int rep = 0; rep < 8; ++rep) {
        if (emitted < 8) {
            w[emitted++] = (uint32_t)acc;
            acc >>= 32;
        }
    }
    top = (uint32_t)acc;
}
// carry propagation: same value, limbs below 2^29 (top limb takes the rest)
L29_HD L9 normalise(const L9& a) {
    L9 r;
    uint32_t c = 0;
L29_UNROLL
    for (int j = 0; j < NL - 1; ++j) {
        const uint64_t v = (uint64_t)a.l[j] + c;
        r.l[j] = (uint32_t)v & M29;
        c = (uint32_t)(v >> 29);
    }
    r.l[NL - 1] = a.l[NL - 1] + c;
    return r;
}
L29_HD L9 add(const L9& a, const L9& b) {
    L9 r;
L29_UNROLL
    for (int j = 0; j < NL; ++j) r.l[j] = a.l[j] + b.l[j];
    return r;
}
// a - b + k p for a normalised b below 2^256 (no limb goes negative); value below a + k p
L29_HD L9 sub_kp(const Desc29& d, const L9& a, const L9& b) {
    L9 r;
L29_UNROLL
    for (int j = 0; j < NL; ++j) r.l[j] = a.l[j] + d.kp[j] - b.l[j];
    return r;
}

// x y 2^-261 mod p as 64-bit columns (not yet carried): value = sum t[j] 2^(29 j) < x y / 2^261 + p, every t[j] < 2^63.7.
// Limb bounds: max limb(a) * max limb(b) <= 2^60.2.
L29_HD void mont_cols(const Desc29& d, const L9& a, const L9& b, uint64_t (&t)[NL]) {
L29_UNROLL
    for (int j = 0; j < NL; ++j) t[j] = 0;
L29_UNROLL
    for (int i = 0; i < NL; ++i) {
L29_UNROLL
        for (int j = 0; j < NL; ++j) t[j] += (uint64_t)a.l[j] * b.l[i];
        const uint32_t m = ((uint32_t)t[0] * d.n0) & M29;
L29_UNROLL
        for (int j = 0; j < NL; ++j) t[j] += (uint64_t)m * d.p[j];
        const uint64_t c = t[0] >> 29;  // t[0] is a multiple of 2^29 now
L29_UNROLL
        for (int j = 0; j < NL - 1; ++j) t[j] = t[j + 1];
        t[NL - 1] = 0;
        t[0] += c;
    }
}
// columns -> normalised limbs
L29_HD L9 carry_cols(uint64_t (&t)[NL]) {
    L9 r;
L29_UNROLL
    for (int j = 0; j < NL - 1; ++j) {
        r.l[j] = (uint32_t)t[j] & M29;
        t[j + 1] += t[j] >> 29;
    }
    r.l[NL - 1] = (uint32_t)t[NL - 1];
    return r;
}
L29_HD L9 mont(const Desc29& d, const L9& a, const L9& b) {
    uint64_t t[NL];
    mont_cols(d, a, b, t);
    return carry_cols(t);
}

// ---------------------------------------------------------------------------------------------- product scanning (r2b)
// The same product with the two phases separated: first all 81 partial products into 17 independent 64-bit columns
// (no instruction waits on another column), then the Montgomery reduction column by column.  Step k makes column k a
// multiple of 2^29 by adding m_k p (m_k from the low word of the column) and hands its upper bits to column k + 1.
// P0ONE: moduli with p = 1 (mod 2^29) (every field with 2-adicity >= 29, BLS12-381 Fr among them): n0 = -1, so
// m_k = -c_k mod 2^29 without a multiplication.
// Column bound: 9 (A B) from the products + 9 2^58 from the reduction + a carry below 2^36, A B <= 2^60.2 as above.
// The result occupies columns 9 .. 16 (column 17 of a 9 x 9 product is empty): eight 64-bit columns out.
template <bool P0ONE>
L29_HD void mont_ps_cols(const Desc29& d, const L9& a, const L9& b, uint64_t (&out)[NL - 1]) {
    uint64_t c[2 * NL - 1];
L29_UNROLL
    for (int k = 0; k < 2 * NL - 1; ++k) c[k] = 0;
L29_UNROLL
    for (int i = 0; i < NL; ++i)
L29_UNROLL
        for (int j = 0; j < NL; ++j) c[i + j] += (uint64_t)a.l[j] * b.l[i];
L29_UNROLL
    for (int k = 0; k < NL; ++k) {
        const uint32_t lo = (uint32_t)c[k];
        const uint32_t m = P0ONE ? ((0u - lo) & M29) : ((lo * d.n0) & M29);
L29_UNROLL
        for (int j = 0; j < NL; ++j) c[k + j] += (uint64_t)m * (P0ONE && j == 0 ? 1u : d.p[j]);
        if (k + 1 < 2 * NL - 1) c[k + 1] += c[k] >> 29;
    }
L29_UNROLL
    for (int j = 0; j < NL - 1; ++j) out[j] = c[NL + j];
}
// eight columns -> nine limbs, carries rippled (limbs below 2^29, the top limb takes the rest)
L29_HD L9 carry_cols8(uint64_t (&t)[NL - 1]) {
    L9 r;
L29_UNROLL
    for (int j = 0; j < NL - 2; ++j) {
        r.l[j] = (uint32_t)t[j] & M29;
        t[j + 1] += t[j] >> 29;
    }
    r.l[NL - 2] = (uint32_t)t[NL - 2] & M29;
    r.l[NL - 1] = (uint32_t)(t[NL - 2] >> 29);
    return r;
}
// eight columns -> nine limbs without a carry chain: every column is cut into its three 29-bit digits and limb j takes
// digit 0 of column j, digit 1 of column j - 1 and digit 2 of column j - 2.  Limbs below 2^30 + 2^6 (not normalised):
// fine as an operand of another product (9 2^60.02 + 9 2^58 < 2^64), not for limb-wise sums.
L29_HD L9 split_cols8(const uint64_t (&t)[NL - 1]) {
    uint32_t d0[NL - 1], d1[NL - 1], d2[NL - 1];
L29_UNROLL
    for (int j = 0; j < NL - 1; ++j) {
        d0[j] = (uint32_t)t[j] & M29;
        d1[j] = (uint32_t)(t[j] >> 29) & M29;
        d2[j] = (uint32_t)(t[j] >> 58);
    }
    L9 r;
L29_UNROLL
    for (int j = 0; j < NL - 1; ++j) r.l[j] = d0[j] + (j >= 1 ? d1[j - 1] : 0u) + (j >= 2 ? d2[j - 2] : 0u);
    r.l[NL - 1] = (uint32_t)(t[NL - 2] >> 29) + d2[NL - 3];  // digit 1 of the last column unmasked: it takes what is above
    return r;
}
template <bool P0ONE>
L29_HD L9 mont_ps(const Desc29& d, const L9& a, const L9& b) {
    uint64_t t[NL - 1];
    mont_ps_cols<P0ONE>(d, a, b, t);
    return carry_cols8(t);
}
template <bool P0ONE>
L29_HD L9 mont_ps_par(const Desc29& d, const L9& a, const L9& b) {
    uint64_t t[NL - 1];
    mont_ps_cols<P0ONE>(d, a, b, t);
    return split_cols8(t);
}

// ---------------------------------------------------------------------------------------------- host set-up
// p as four little-endian 64-bit limbs.  False if the modulus is outside the range the bounds above were derived for
// (the caller then keeps the 32-bit-limb kernel).
inline bool make_desc(const uint64_t p64[4], uint32_t bits, Desc29* out) {
    if (bits > 255 || bits < 250) return false;
    uint32_t w[8];
    for (int i = 0; i < 4; ++i) {
        w[2 * i] = (uint32_t)p64[i];
        w[2 * i + 1] = (uint32_t)(p64[i] >> 32);
    }
    const L9 pl = from_words(w);
    for (int j = 0; j < NL; ++j) out->p[j] = pl.l[j];
    uint32_t inv = 1;  // p^-1 mod 2^32 by Newton iteration (p odd)
    for (int i = 0; i < 6; ++i) inv *= 2u - pl.l[0] * inv;
    out->n0 = (0u - inv) & M29;
    const uint32_t e = 258 - bits;  // k = 2^e: 2^257 <= k p < 2^258
    out->k = 1u << e;
    uint64_t limbs[NL];
    for (int j = 0; j < NL; ++j) limbs[j] = pl.l[j];
    for (uint32_t s = 0; s < e; ++s) {  // double, then carry (the top limb keeps what is above bit 261)
        uint64_t c = 0;
        for (int j = 0; j < NL; ++j) {
            const uint64_t v = 2 * limbs[j] + c;
            if (j < NL - 1) {
                limbs[j] = v & M29;
                c = v >> 29;
            } else {
                limbs[j] = v;
            }
        }
    }
    for (int j = 0; j < NL - 1; ++j) {  // borrow one unit of limb j+1 into limb j (wraps resolve: arithmetic mod 2^64)
        limbs[j] += 1u << 29;
        limbs[j + 1] -= 1;
    }
    if (limbs[NL - 1] < (1u << 24) || limbs[NL - 1] >= (1ull << 30)) return false;
    for (int j = 0; j < NL; ++j) {
        if (j < NL - 1 && (limbs[j] < M29 || limbs[j] >= (1ull << 31))) return false;
        out->kp[j] = (uint32_t)limbs[j];
    }
    return true;
}

}  // namespace l29
}  // namespace scb
