// hostfield.hpp -- host-side prime-field arithmetic for the protocol layer (interpolation of the
// (d+1) round sums, verifier checks, serialization).  Same in-memory format as the device side and
// as [ARK] ark_ff::Fp<MontBackend<_,N>,N>: N little-endian u64 limbs, Montgomery form, R = 2^(64N).
// This is product code (the host half of the drop-in), not the oracle.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace scb {

constexpr int kHostMaxLimbs = 4;
typedef unsigned __int128 u128_t;

struct Fe {  // field element, limbs beyond n are zero
    std::array<uint64_t, kHostMaxLimbs> l{{0, 0, 0, 0}};
    bool operator==(const Fe& o) const { return l == o.l; }
    bool operator!=(const Fe& o) const { return l != o.l; }
};

class HostField {
   public:
    uint32_t n = 0;
    uint32_t bits = 0;
    uint64_t p[kHostMaxLimbs] = {0, 0, 0, 0};
    uint64_t inv = 0;  // -p^{-1} mod 2^64
    Fe one_, r2_;

    HostField() = default;
    HostField(uint32_t n_limbs, const uint64_t* modulus) {
        if (n_limbs < 1 || n_limbs > (uint32_t)kHostMaxLimbs) throw std::invalid_argument("n_limbs must be 1..4");
        if (!(modulus[0] & 1)) throw std::invalid_argument("modulus must be odd");
        n = n_limbs;
        for (uint32_t i = 0; i < n; ++i) p[i] = modulus[i];
        if (p[n - 1] == 0) throw std::invalid_argument("top limb of the modulus is zero");
        if (n == 1 && p[0] < 3) throw std::invalid_argument("modulus must be >= 3");
        bits = 64 * (n - 1) + (64 - __builtin_clzll(p[n - 1]));
        uint64_t x = 1;
        for (int i = 0; i < 6; ++i) x *= 2 - p[0] * x;  // Newton iteration: p^{-1} mod 2^64
        inv = (uint64_t)0 - x;
        Fe v;
        v.l[0] = 1;
        for (uint32_t i = 0; i < 128 * n; ++i) {
            v = dbl_raw(v);
            if (i + 1 == 64 * n) one_ = v;
        }
        r2_ = v;
    }

    uint32_t ser_bytes() const { return (bits + 7) / 8; }  // [ARK] uncompressed size of an Fp
    Fe zero() const { return Fe(); }
    Fe one() const { return one_; }
    bool is_zero(const Fe& a) const { return a == Fe(); }

    Fe add(const Fe& a, const Fe& b) const {
        Fe s;
        u128_t c = 0;
        for (uint32_t i = 0; i < n; ++i) {
            c += (u128_t)a.l[i] + b.l[i];
            s.l[i] = (uint64_t)c;
            c >>= 64;
        }
        if (c || geq_p(s)) sub_p(s);
        return s;
    }
    Fe sub(const Fe& a, const Fe& b) const {
        Fe d;
        uint64_t borrow = 0;
        for (uint32_t i = 0; i < n; ++i) {
            u128_t t = (u128_t)a.l[i] - b.l[i] - borrow;
            d.l[i] = (uint64_t)t;
            borrow = (uint64_t)(t >> 64) & 1;
        }
        if (borrow) {
            u128_t c = 0;
            for (uint32_t i = 0; i < n; ++i) {
                c += (u128_t)d.l[i] + p[i];
                d.l[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return d;
    }
    Fe neg(const Fe& a) const { return sub(zero(), a); }
    Fe mul(const Fe& a, const Fe& b) const {  // CIOS
        uint64_t t[kHostMaxLimbs + 2] = {0};
        for (uint32_t i = 0; i < n; ++i) {
            u128_t c = 0;
            for (uint32_t j = 0; j < n; ++j) {
                c += (u128_t)a.l[j] * b.l[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[n];
            t[n] = (uint64_t)c;
            t[n + 1] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * inv;
            c = ((u128_t)m * p[0] + t[0]) >> 64;
            for (uint32_t j = 1; j < n; ++j) {
                c += (u128_t)m * p[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[n];
            t[n - 1] = (uint64_t)c;
            t[n] = t[n + 1] + (uint64_t)(c >> 64);
        }
        Fe r;
        for (uint32_t i = 0; i < n; ++i) r.l[i] = t[i];
        if (t[n] || geq_p(r)) sub_p(r);
        return r;
    }
    Fe pow(Fe base, const uint64_t* e, uint32_t e_limbs) const {
        Fe acc = one_;
        for (int i = (int)e_limbs * 64 - 1; i >= 0; --i) {
            acc = mul(acc, acc);
            if ((e[i / 64] >> (i % 64)) & 1) acc = mul(acc, base);
        }
        return acc;
    }
    Fe pow_u64(const Fe& base, uint64_t e) const { return pow(base, &e, 1); }
    Fe inverse(const Fe& a) const {  // Fermat: a^(p-2); caller guarantees a != 0
        uint64_t e[kHostMaxLimbs];
        uint64_t borrow = 2;
        for (uint32_t i = 0; i < n; ++i) {
            e[i] = p[i] - borrow;
            borrow = p[i] < borrow ? 1 : 0;
        }
        return pow(a, e, n);
    }
    // small integer -> Montgomery form
    Fe from_u64(uint64_t v) const {
        Fe raw;
        raw.l[0] = v;
        if (n == 1) raw.l[0] = v % p[0];
        return mul(raw, r2_);
    }
    Fe to_mont(const Fe& canonical) const { return mul(canonical, r2_); }
    Fe from_mont(const Fe& m) const {
        Fe o;
        o.l[0] = 1;
        return mul(m, o);
    }
    bool is_canonical(const Fe& a) const { return !geq_p(a); }

    // [ARK] serialize_uncompressed of an Fp: canonical value, ceil(bits/8) little-endian bytes
    void serialize(const Fe& a, std::vector<uint8_t>& out) const {
        Fe c = from_mont(a);
        uint32_t nb = ser_bytes();
        for (uint32_t i = 0; i < nb; ++i) out.push_back((uint8_t)(c.l[i / 8] >> (8 * (i % 8))));
    }
    // returns false when the bytes do not encode a canonical element
    bool deserialize(const uint8_t* data, Fe& out) const {
        Fe c;
        uint32_t nb = ser_bytes();
        for (uint32_t i = 0; i < nb; ++i) c.l[i / 8] |= (uint64_t)data[i] << (8 * (i % 8));
        if (geq_p(c)) return false;
        out = to_mont(c);
        return true;
    }
    // [ARK] from_be_bytes_mod_order: big-endian integer reduced mod p (Horner, one byte at a time)
    Fe from_be_bytes_mod_order(const uint8_t* data, size_t len) const {
        Fe acc;  // canonical-value arithmetic carried out in Montgomery form
        Fe c256 = from_u64(256);
        for (size_t i = 0; i < len; ++i) acc = add(mul(acc, c256), from_u64(data[i]));
        return acc;
    }

    void load(const uint64_t* w, Fe& a) const {
        a = Fe();
        for (uint32_t i = 0; i < n; ++i) a.l[i] = w[i];
    }
    void store(const Fe& a, uint64_t* w) const {
        for (uint32_t i = 0; i < n; ++i) w[i] = a.l[i];
    }

    // 2a mod p on RAW (non-Montgomery) limbs: lets a caller build r 2^k mod p tables as plain integers (engine.cu: fold table)
    Fe double_raw(const Fe& a) const { return dbl_raw(a); }

   private:
    bool geq_p(const Fe& a) const {
        for (int i = (int)n - 1; i >= 0; --i) {
            if (a.l[i] > p[i]) return true;
            if (a.l[i] < p[i]) return false;
        }
        return true;
    }
    void sub_p(Fe& a) const {
        uint64_t borrow = 0;
        for (uint32_t i = 0; i < n; ++i) {
            u128_t t = (u128_t)a.l[i] - p[i] - borrow;
            a.l[i] = (uint64_t)t;
            borrow = (uint64_t)(t >> 64) & 1;
        }
    }
    Fe dbl_raw(const Fe& a) const {  // 2a mod p on raw (non-Montgomery) values
        Fe s;
        uint64_t c = 0;
        for (uint32_t i = 0; i < n; ++i) {
            s.l[i] = (a.l[i] << 1) | c;
            c = a.l[i] >> 63;
        }
        if (c || geq_p(s)) sub_p(s);
        return s;
    }
};

}  // namespace scb
