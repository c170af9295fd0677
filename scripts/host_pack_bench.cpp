// host_pack_bench.cpp -- what bounds the host lane of the narrowing upload (hostpack.hpp: pack21_host)?
// Times variants of the 8-byte -> 21-bit wire pack over T threads on tables far larger than the caches:
//   scalar      the loop of hostpack.hpp as it was (three loads, one word out)
//   scalar+pf   the same with a software prefetch D bytes ahead (crosses the 4 KB page the hardware prefetcher stops at)
//   avx2        12 entries -> 4 words per iteration with 256-bit loads
//   read-only   an OR-reduction of the source (the load-side ceiling of a thread)
//   memcpy      plain copy of the same bytes
// g++ -O3 -march=native -pthread scripts/host_pack_bench.cpp -o /tmp/host_pack_bench && /tmp/host_pack_bench [threads] [log2 entries]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>

static uint64_t pack_scalar(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n, uint64_t pm1) {
    uint64_t acc = 0, i = 0, j = 0;
    for (; i + 3 <= n; i += 3, ++j) {
        const uint64_t a = src[i], b = src[i + 1], c = src[i + 2];
        acc |= a | b | c | (pm1 - a) | (pm1 - b) | (pm1 - c);
        dst[j] = a | (b << 21) | (c << 42);
    }
    return acc;
}

template <int D>
static uint64_t pack_scalar_pf(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n, uint64_t pm1) {
    uint64_t acc = 0, i = 0, j = 0;
    for (; i + 24 <= n; i += 24, j += 8) {  // 24 entries = 192 B = three lines in, one line out
        _mm_prefetch((const char*)(src + i) + D, _MM_HINT_T0);
        _mm_prefetch((const char*)(src + i) + D + 64, _MM_HINT_T0);
        _mm_prefetch((const char*)(src + i) + D + 128, _MM_HINT_T0);
#pragma GCC unroll 8
        for (int k = 0; k < 8; ++k) {
            const uint64_t a = src[i + 3 * k], b = src[i + 3 * k + 1], c = src[i + 3 * k + 2];
            acc |= a | b | c | (pm1 - a) | (pm1 - b) | (pm1 - c);
            dst[j + k] = a | (b << 21) | (c << 42);
        }
    }
    for (; i + 3 <= n; i += 3, ++j) {
        const uint64_t a = src[i], b = src[i + 1], c = src[i + 2];
        acc |= a | b | c | (pm1 - a) | (pm1 - b) | (pm1 - c);
        dst[j] = a | (b << 21) | (c << 42);
    }
    return acc;
}

#ifdef __AVX2__
// 12 entries (three 256-bit loads) -> four words.  v0 = e0..e3, v1 = e4..e7, v2 = e8..e11; word w takes entries 3w..3w+2.
// Gather by permutes: A = (e0,e3,e6,e9), B = (e1,e4,e7,e10), C = (e2,e5,e8,e11); out = A | B << 21 | C << 42.
template <int D, bool NT>
static uint64_t pack_avx2(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n, uint64_t pm1) {
    __m256i acc = _mm256_setzero_si256();
    const __m256i pv = _mm256_set1_epi64x((long long)pm1);
    uint64_t i = 0, j = 0;
    for (; i + 12 <= n; i += 12, j += 4) {
        if (D) {
            _mm_prefetch((const char*)(src + i) + D, _MM_HINT_T0);
            _mm_prefetch((const char*)(src + i) + D + 64, _MM_HINT_T0);  // 96 B per iteration: every other one overlaps, cheap
        }
        const __m256i v0 = _mm256_loadu_si256((const __m256i*)(src + i));
        const __m256i v1 = _mm256_loadu_si256((const __m256i*)(src + i + 4));
        const __m256i v2 = _mm256_loadu_si256((const __m256i*)(src + i + 8));
        acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(v0, v1), v2));
        acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(_mm256_sub_epi64(pv, v0), _mm256_sub_epi64(pv, v1)), _mm256_sub_epi64(pv, v2)));
        // A: lanes (v0[0], v0[3], v1[2], v2[1]);  B: (v0[1], v1[0], v1[3], v2[2]);  C: (v0[2], v1[1], v2[0], v2[3])
        const __m256i a01 = _mm256_permute4x64_epi64(v0, 0x0C);                                   // (v0[0], v0[3], ., .)
        const __m256i a2 = _mm256_permute4x64_epi64(v1, 0x20);                                    // (., ., v1[2], .)
        const __m256i a3 = _mm256_permute4x64_epi64(v2, 0x40);                                    // (., ., ., v2[1])
        const __m256i A = _mm256_blend_epi32(_mm256_blend_epi32(a01, a2, 0x30), a3, 0xC0);
        const __m256i b0 = _mm256_permute4x64_epi64(v0, 0x01);                                    // (v0[1], ., ., .)
        const __m256i b12 = _mm256_permute4x64_epi64(v1, 0x30);                                   // (., v1[0], v1[3], .)
        const __m256i b3 = _mm256_permute4x64_epi64(v2, 0x80);                                    // (., ., ., v2[2])
        const __m256i B = _mm256_blend_epi32(_mm256_blend_epi32(b0, b12, 0x3C), b3, 0xC0);
        const __m256i c0 = _mm256_permute4x64_epi64(v0, 0x02);                                    // (v0[2], ., ., .)
        const __m256i c1 = _mm256_permute4x64_epi64(v1, 0x04);                                    // (., v1[1], ., .)
        const __m256i c23 = _mm256_permute4x64_epi64(v2, 0xC0);                                   // (., ., v2[0], v2[3])
        const __m256i C = _mm256_blend_epi32(_mm256_blend_epi32(c0, c1, 0x0C), c23, 0xF0);
        const __m256i w = _mm256_or_si256(A, _mm256_or_si256(_mm256_slli_epi64(B, 21), _mm256_slli_epi64(C, 42)));
        if (NT) _mm256_stream_si256((__m256i*)(dst + j), w);
        else _mm256_storeu_si256((__m256i*)(dst + j), w);
    }
    uint64_t t[4];
    _mm256_storeu_si256((__m256i*)t, acc);
    uint64_t r = t[0] | t[1] | t[2] | t[3];
    for (; i + 3 <= n; i += 3, ++j) {
        const uint64_t a = src[i], b = src[i + 1], c = src[i + 2];
        r |= a | b | c | (pm1 - a) | (pm1 - b) | (pm1 - c);
        dst[j] = a | (b << 21) | (c << 42);
    }
    if (NT) _mm_sfence();
    return r;
}
#endif

static uint64_t read_only(const uint64_t* __restrict__ src, uint64_t* __restrict__, uint64_t n, uint64_t) {
    uint64_t a = 0, b = 0, c = 0, d = 0;
    for (uint64_t i = 0; i + 4 <= n; i += 4) a |= src[i], b |= src[i + 1], c |= src[i + 2], d |= src[i + 3];
    return a | b | c | d;
}
static uint64_t copy_only(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n, uint64_t) {
    memcpy(dst, src, n * 8 / 3);  // the packed size, so that the store side matches
    return read_only(src, dst, n, 0);
}

typedef uint64_t (*fn_t)(const uint64_t*, uint64_t*, uint64_t, uint64_t);

int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
    const int lg = argc > 2 ? atoi(argv[2]) : 27;
    const uint64_t n = 1ull << lg, chunk = 1ull << 20, p = 1572869;
    uint64_t* src = (uint64_t*)aligned_alloc(4096, n * 8);
    std::vector<uint64_t*> stage(T);
    for (auto& s : stage) s = (uint64_t*)aligned_alloc(4096, 2 * chunk * 8 / 2);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                for (uint64_t i = n / T * t; i < n / T * (t + 1); ++i) src[i] = (i * 0x9E3779B97F4A7C15ull >> 20) % p;
                memset(stage[t], 0, chunk * 8);
            });
        for (auto& x : th) x.join();
    }
    std::vector<uint64_t> want((chunk + 2) / 3 + 8);
    uint64_t* got = (uint64_t*)aligned_alloc(4096, chunk * 4);
    struct V { const char* name; fn_t f; };
    std::vector<V> vs = {{"scalar", pack_scalar},
                         {"scalar+pf512", pack_scalar_pf<512>},
                         {"scalar+pf1024", pack_scalar_pf<1024>},
                         {"scalar+pf2048", pack_scalar_pf<2048>},
                         {"scalar+pf4096", pack_scalar_pf<4096>},
#ifdef __AVX2__
                         {"avx2", pack_avx2<0, false>},
                         {"avx2+pf1024", pack_avx2<1024, false>},
                         {"avx2+pf2048", pack_avx2<2048, false>},
                         {"avx2+pf4096", pack_avx2<4096, false>},
                         {"avx2+pf2048+nt", pack_avx2<2048, true>},
#endif
                         {"read-only", read_only},
                         {"copy(packed size)+read", copy_only}};
    pack_scalar(src, want.data(), chunk, p - 1);
    for (auto& v : vs) {
        if (strncmp(v.name, "read", 4) && strncmp(v.name, "copy", 4)) {
            memset(got, 0, chunk * 4);
            v.f(src, got, chunk, p - 1);
            const uint64_t words = chunk / 3;
            if (memcmp(got, want.data(), words * 8) != 0) {
                printf("{\"variant\": \"%s\", \"error\": \"output differs\"}\n", v.name);
                continue;
            }
        }
        double best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            std::vector<uint64_t> accs(T);
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    uint64_t a = 0;
                    for (uint64_t c = t; c < n / chunk; c += T) a |= v.f(src + c * chunk, stage[t] + (c / T & 1) * (chunk / 2), chunk, p - 1);
                    accs[t] = a;
                });
            for (auto& x : th) x.join();
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (s < best) best = s;
            if (accs[0] >> 63) printf("bad\n");
        }
        printf("{\"variant\": \"%s\", \"threads\": %d, \"entries_log2\": %d, \"source_GBps\": %.1f, \"per_thread_GBps\": %.2f}\n", v.name, T, lg,
               n * 8 / best / 1e9, n * 8 / best / 1e9 / T);
        fflush(stdout);
    }
    return 0;
}
