// fiat_shamir.hpp -- host mirror of the challenge derivation used by
// /root/reference/fiat-shamir/src/lib.rs:75-98,123-143:
//   hasher = DefaultFieldHasher<Sha256>::new(&[])      (empty DST, :78)
//   r_j    = hasher.hash_to_field::<1>(g_1 || ... || g_j)[0]   (cumulative byte string, :87-88)
// [ARK] DefaultFieldHasher<Sha256, 128>: L = ceil((bits(p) + 128) / 8); uniform bytes =
// expand_message_xmd(msg, DST, L) per RFC 9380 with SHA-256, EXCEPT that Z_pad has L zero bytes
// (ark's `block_size: len_per_base_elem`) instead of the hash's 64-byte block; DST' = DST || len(DST);
// the L bytes are read big-endian and reduced mod p (from_be_bytes_mod_order).
// The reference holds no vector for these bytes ("parity unpinned"); the Python
// oracle restates the same published algorithm on top of hashlib and must agree byte for byte.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hostfield.hpp"
#include "../options.hpp"

#if defined(__x86_64__) && defined(__GNUC__)
#include <cpuid.h>
#include <immintrin.h>
#define SCB_SHA_NI 1
#endif

namespace scb {

#ifdef SCB_SHA_NI
// SHA-256 compression with the x86 SHA extensions (same function as Sha256::compress_scalar; chosen at run time).
// The challenge derivation sits on the prover's critical path once the tables are small: two hash-to-field
// evaluations per turn-around of the resident kernels.
inline bool sha_ni_available() {
    if (opt(OPT_sha_scalar) != 0) return false;  // tests: force the portable compression function
    static const bool ok = [] {
        unsigned a = 0, b = 0, c = 0, d = 0;
        if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
        const bool sha = (b >> 29) & 1;
        if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
        const bool ssse3 = (c >> 9) & 1, sse41 = (c >> 19) & 1;
        return sha && ssse3 && sse41;
    }();
    return ok;
}
__attribute__((target("sha,sse4.1,ssse3"))) inline void sha256_compress_ni(uint32_t state[8], const uint8_t* blk, const uint32_t* k) {
    const __m128i mask = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i tmp = _mm_loadu_si128((const __m128i*)&state[0]);  // d c b a
    __m128i st1 = _mm_loadu_si128((const __m128i*)&state[4]);  // h g f e
    tmp = _mm_shuffle_epi32(tmp, 0xB1);                        // c d a b
    st1 = _mm_shuffle_epi32(st1, 0x1B);                        // e f g h
    __m128i st0 = _mm_alignr_epi8(tmp, st1, 8);                // a b e f
    st1 = _mm_blend_epi16(st1, tmp, 0xF0);                     // c d g h
    const __m128i save0 = st0, save1 = st1;
    __m128i m[4];
    for (int g = 0; g < 16; ++g) {
        if (g < 4) {
            m[g] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(blk + 16 * g)), mask);
        } else {  // W[4g..4g+3] from the previous sixteen words
            const __m128i x0 = m[g & 3], x1 = m[(g + 1) & 3], x2 = m[(g + 2) & 3], x3 = m[(g + 3) & 3];
            m[g & 3] = _mm_sha256msg2_epu32(_mm_add_epi32(_mm_sha256msg1_epu32(x0, x1), _mm_alignr_epi8(x3, x2, 4)), x3);
        }
        __m128i msg = _mm_add_epi32(m[g & 3], _mm_loadu_si128((const __m128i*)(k + 4 * g)));
        st1 = _mm_sha256rnds2_epu32(st1, st0, msg);
        msg = _mm_shuffle_epi32(msg, 0x0E);
        st0 = _mm_sha256rnds2_epu32(st0, st1, msg);
    }
    st0 = _mm_add_epi32(st0, save0);
    st1 = _mm_add_epi32(st1, save1);
    tmp = _mm_shuffle_epi32(st0, 0x1B);         // f e b a
    st1 = _mm_shuffle_epi32(st1, 0xB1);         // d c h g
    st0 = _mm_blend_epi16(tmp, st1, 0xF0);      // d c b a
    st1 = _mm_alignr_epi8(st1, tmp, 8);         // h g f e
    _mm_storeu_si128((__m128i*)&state[0], st0);
    _mm_storeu_si128((__m128i*)&state[4], st1);
}
#endif

class Sha256 {
   public:
    Sha256() { reset(); }
    void reset() {
        static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                       0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        std::memcpy(h_, iv, sizeof iv);
        len_ = 0;
        fill_ = 0;
    }
    void update(const uint8_t* data, size_t n) {
        len_ += n;
        while (n) {
            size_t take = 64 - fill_ < n ? 64 - fill_ : n;
            std::memcpy(buf_ + fill_, data, take);
            fill_ += take;
            data += take;
            n -= take;
            if (fill_ == 64) {
                compress(buf_);
                fill_ = 0;
            }
        }
    }
    void finalize(uint8_t out[32]) {
        const uint64_t bitlen = len_ * 8;
        buf_[fill_++] = 0x80;
        if (fill_ > 56) {
            std::memset(buf_ + fill_, 0, 64 - fill_);
            compress(buf_);
            fill_ = 0;
        }
        std::memset(buf_ + fill_, 0, 56 - fill_);
        for (int i = 0; i < 8; ++i) buf_[56 + i] = (uint8_t)(bitlen >> (56 - 8 * i));
        compress(buf_);
        fill_ = 0;
        for (int i = 0; i < 8; ++i) {
            out[4 * i] = (uint8_t)(h_[i] >> 24);
            out[4 * i + 1] = (uint8_t)(h_[i] >> 16);
            out[4 * i + 2] = (uint8_t)(h_[i] >> 8);
            out[4 * i + 3] = (uint8_t)h_[i];
        }
    }

   private:
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void compress(const uint8_t* blk) {
        alignas(16) static const uint32_t k[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
            0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
            0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
            0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
            0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
            0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
            0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#ifdef SCB_SHA_NI
        if (use_ni_) {
            sha256_compress_ni(h_, blk, k);
            return;
        }
#endif
        uint32_t w[64];
        for (int i = 0; i < 16; ++i)
            w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h_[0], b = h_[1], c = h_[2], d = h_[3], e = h_[4], f = h_[5], g = h_[6], h = h_[7];
        for (int i = 0; i < 64; ++i) {
            uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = h + S1 + ch + k[i] + w[i];
            uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h_[0] += a; h_[1] += b; h_[2] += c; h_[3] += d; h_[4] += e; h_[5] += f; h_[6] += g; h_[7] += h;
    }
    uint32_t h_[8];
    uint64_t len_;
    uint8_t buf_[64];
    size_t fill_;
#ifdef SCB_SHA_NI
    bool use_ni_ = sha_ni_available();

   public:
    void force_scalar(bool on) { use_ni_ = !on && sha_ni_available(); }  // tests: both code paths must agree
#endif
};

// [ARK] ExpanderXmd::expand (block_size = Z_pad length)
inline std::vector<uint8_t> expand_message_xmd(const uint8_t* msg, size_t msg_len, const std::vector<uint8_t>& dst, size_t n,
                                               size_t block_size) {
    const size_t b_len = 32;
    const size_t ell = (n + b_len - 1) / b_len;
    std::vector<uint8_t> dst_prime(dst);
    dst_prime.push_back((uint8_t)dst.size());
    std::vector<uint8_t> z_pad(block_size, 0);
    uint8_t lib_str[2] = {(uint8_t)(n >> 8), (uint8_t)n};
    uint8_t b0[32], bi[32];
    Sha256 h;
    h.update(z_pad.data(), z_pad.size());
    h.update(msg, msg_len);
    h.update(lib_str, 2);
    uint8_t zero = 0;
    h.update(&zero, 1);
    h.update(dst_prime.data(), dst_prime.size());
    h.finalize(b0);
    h.reset();
    h.update(b0, 32);
    uint8_t one = 1;
    h.update(&one, 1);
    h.update(dst_prime.data(), dst_prime.size());
    h.finalize(bi);
    std::vector<uint8_t> out(bi, bi + 32);
    for (size_t i = 2; i <= ell; ++i) {
        uint8_t x[32];
        for (int k = 0; k < 32; ++k) x[k] = b0[k] ^ bi[k];
        h.reset();
        h.update(x, 32);
        uint8_t ib = (uint8_t)i;
        h.update(&ib, 1);
        h.update(dst_prime.data(), dst_prime.size());
        h.finalize(bi);
        out.insert(out.end(), bi, bi + 32);
    }
    out.resize(n);
    return out;
}

// Incremental form of the same chain: the hash input of round j is the concatenation g_1 || ... || g_j
// (fiat-shamir/src/lib.rs:82-92), so the SHA-256 state after Z_pad || g_1 || ... || g_j is kept and only the
// new message bytes are compressed each round (identical digest, O(new bytes) instead of O(all bytes)).
class FsChain {
   public:
    explicit FsChain(const HostField& F) : F_(F), L_((F.bits + 128 + 7) / 8) {
        std::vector<uint8_t> z_pad(L_, 0);
        prefix_.update(z_pad.data(), z_pad.size());
    }
    void absorb(const uint8_t* data, size_t n) { prefix_.update(data, n); }
    // hash_to_field::<1>(everything absorbed so far)[0]
    Fe challenge() const {
        const size_t ell = (L_ + 31) / 32;
        const uint8_t dst_prime[1] = {0};  // DST = "" => DST' = I2OSP(0, 1)
        uint8_t lib_str[2] = {(uint8_t)(L_ >> 8), (uint8_t)L_};
        uint8_t b0[32], bi[32], zero = 0, one = 1;
        Sha256 h = prefix_;
        h.update(lib_str, 2);
        h.update(&zero, 1);
        h.update(dst_prime, 1);
        h.finalize(b0);
        h.reset();
        h.update(b0, 32);
        h.update(&one, 1);
        h.update(dst_prime, 1);
        h.finalize(bi);
        uint8_t uniform[96];
        std::memcpy(uniform, bi, 32);
        for (size_t i = 2; i <= ell && i <= 3; ++i) {
            uint8_t x[32];
            for (int k = 0; k < 32; ++k) x[k] = b0[k] ^ bi[k];
            h.reset();
            h.update(x, 32);
            uint8_t ib = (uint8_t)i;
            h.update(&ib, 1);
            h.update(dst_prime, 1);
            h.finalize(bi);
            std::memcpy(uniform + 32 * (i - 1), bi, 32);
        }
        return F_.from_be_bytes_mod_order(uniform, L_);
    }

   private:
    const HostField& F_;
    size_t L_;
    Sha256 prefix_;
};

// hasher.hash_to_field::<1>(msg)[0] with H::new(&[])
inline Fe hash_to_field(const HostField& F, const uint8_t* msg, size_t len) {
    const size_t L = (F.bits + 128 + 7) / 8;
    std::vector<uint8_t> uniform = expand_message_xmd(msg, len, std::vector<uint8_t>(), L, L);
    return F.from_be_bytes_mod_order(uniform.data(), uniform.size());
}

}  // namespace scb
