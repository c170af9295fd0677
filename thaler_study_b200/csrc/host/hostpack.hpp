// hostpack.hpp -- host side of the packed upload (scb_poly_product_from_host, upload_engine.inc).
//
// For the reference's own fields (Fp64, p < 2^28) every canonical Montgomery value fits 32 bits, yet ark's in-memory
// format spends 8 bytes on it, and an end-to-end proof of caller-owned tables is bound by those bytes crossing PCIe
// (2^28 x 3 x 8 B = 6.4 GB at ~54 GB/s = 119 ms against 2.7 ms of kernels).  The upload therefore runs two lanes over
// one shared list of table chunks:
//   * pack lane  -- host threads narrow a chunk to uint32 in a pinned staging buffer and queue a 4-byte-per-entry copy;
//   * raw lane   -- one thread queues the chunk as it is (8 bytes per entry) into a small device buffer and a kernel
//                   narrows it on the device (k_pack32).
// Pack workers take chunks from the front of the list, the raw lane from the back, until they meet: the link never
// idles, and the share of chunks that cross at 4 bytes per entry is whatever the host cores manage in that time.
// The raw lane keeps at most a few chunks in flight, so when the pack workers are fast their copies fill the queue.
//
// The scheduler is written against a back end so that tests can run it without a device (scb_host_pack_selftest).
#pragma once
#include <atomic>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace scb {

// Canonical-entry check shared by all lanes: acc |= v | (pm1 - v) with pm1 = p - 1.  For v < p both terms are below p;
// for v >= p the difference wraps and sets bit 63.  So "every entry < p"  <=>  acc >> 63 == 0 -- a real comparison with
// the modulus (not only with 2^bits(p)) at the price of one subtraction per entry.
static inline bool pack_acc_ok(uint64_t acc) { return (acc >> 63) == 0; }

// narrow n 8-byte entries to 4 bytes; returns the check accumulator (pack_acc_ok)
static inline uint64_t pack32_host(const uint64_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n, uint64_t pm1, uint32_t pf_bytes = 0) {
    uint64_t acc = 0, i = 0;
#if defined(__GNUC__)
    if (pf_bytes) {  // see pack21_host
        for (; i + 8 <= n; i += 8) {
            __builtin_prefetch((const char*)(src + i) + pf_bytes, 0, 3);
#pragma GCC unroll 8
            for (int k = 0; k < 8; ++k) {
                const uint64_t v = src[i + k];
                acc |= v | (pm1 - v);
                dst[i + k] = (uint32_t)v;
            }
        }
    }
#endif
    for (; i < n; ++i) {
        const uint64_t v = src[i];
        acc |= v | (pm1 - v);
        dst[i] = (uint32_t)v;
    }
    return acc;
}

#if defined(__SSE2__)
}  // namespace scb
#include <emmintrin.h>
namespace scb {
// the same with streaming stores: the staging buffer is written past the caches (no read-for-ownership of its lines)
static inline uint64_t pack32_host_nt(const uint64_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n, uint64_t pm1) {
    __m128i acc = _mm_setzero_si128();
    const __m128i pv = _mm_set1_epi64x((long long)pm1);
    uint64_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 2));
        acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(a, b), _mm_or_si128(_mm_sub_epi64(pv, a), _mm_sub_epi64(pv, b))));
        const __m128i lo = _mm_castps_si128(_mm_shuffle_ps(_mm_castsi128_ps(a), _mm_castsi128_ps(b), _MM_SHUFFLE(2, 0, 2, 0)));
        _mm_stream_si128((__m128i*)(dst + i), lo);  // staging buffers are 16-byte aligned, chunks are multiples of 4
    }
    uint64_t w[2];
    _mm_storeu_si128((__m128i*)w, acc);
    uint64_t r = w[0] | w[1];
    for (; i < n; ++i) {
        r |= src[i] | (pm1 - src[i]);
        dst[i] = (uint32_t)src[i];
    }
    _mm_sfence();
    return r;
}
#else
static inline uint64_t pack32_host_nt(const uint64_t* src, uint32_t* dst, uint64_t n, uint64_t pm1) { return pack32_host(src, dst, n, pm1); }
#endif

// Three entries below 2^21 per 64-bit word (entry i of a chunk in bits 21*(i%3) .. of word i/3; a last partial word is
// zero-filled): the wire format for fields of at most 21 bits, such as the reference's F_1572869.  ceil(n/3) words.
static inline uint64_t pack21_words(uint64_t n) { return (n + 2) / 3; }
// pf_bytes: software prefetch that far ahead of the loads (0: none).  The hardware prefetcher stops at every 4 KB page and a
// thread has only so many line fills in flight: on the GPU boxes' hosts a pack thread reads 5.2 GB/s without and 8.2 GB/s with a
// prefetch 4 KB ahead (16 threads: 83 -> 131 GB/s of source; scripts/host_pack_bench.cpp, profiles/r02_host_pack_bench.jsonl).
// Prefetches never fault, so running past the end of the chunk is harmless.
template <bool NT>
static inline uint64_t pack21_host(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n, uint64_t pm1, uint32_t pf_bytes = 0) {
    uint64_t acc = 0, i = 0, j = 0;
#if defined(__SSE2__) && defined(__x86_64__)
    if (pf_bytes) {
        for (; i + 24 <= n; i += 24, j += 8) {  // 24 entries = three 64-byte lines in, one line out
            const char* pf = (const char*)(src + i) + pf_bytes;
            _mm_prefetch(pf, _MM_HINT_T0);
            _mm_prefetch(pf + 64, _MM_HINT_T0);
            _mm_prefetch(pf + 128, _MM_HINT_T0);
#pragma GCC unroll 8
            for (int k = 0; k < 8; ++k) {
                const uint64_t a = src[i + 3 * k], b = src[i + 3 * k + 1], c = src[i + 3 * k + 2];
                acc |= a | b | c | (pm1 - a) | (pm1 - b) | (pm1 - c);
                const uint64_t w = a | (b << 21) | (c << 42);
                if (NT) _mm_stream_si64((long long*)(dst + j + k), (long long)w);
                else dst[j + k] = w;
            }
        }
    }
#endif
    for (; i + 3 <= n; i += 3, ++j) {
        const uint64_t a = src[i], b = src[i + 1], c = src[i + 2];
        acc |= a | b | c | (pm1 - a) | (pm1 - b) | (pm1 - c);
        const uint64_t w = a | (b << 21) | (c << 42);
#if defined(__SSE2__) && defined(__x86_64__)
        if (NT) _mm_stream_si64((long long*)(dst + j), (long long)w);
        else dst[j] = w;
#else
        dst[j] = w;
#endif
    }
    if (i < n) {
        const uint64_t a = src[i], b = i + 1 < n ? src[i + 1] : 0;
        acc |= a | b | (pm1 - a) | (pm1 - b);
        dst[j] = a | (b << 21);
    }
#if defined(__SSE2__) && defined(__x86_64__)
    if (NT) _mm_sfence();
#endif
    return acc;
}
static inline void unpack21_host(const uint64_t* src, uint32_t* dst, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) dst[i] = (uint32_t)((src[i / 3] >> (21 * (i % 3))) & 0x1fffffu);
}

struct PackUnits {  // chunks of all tables as one list: unit u = (table u / per_table, offset (u % per_table) * chunk)
    uint64_t per_table = 0, chunk = 0, total = 0;
    std::mutex mu;
    uint64_t front = 0, back = 0;
    PackUnits(uint32_t k, uint64_t len, uint64_t chunk_) : per_table(len / chunk_), chunk(chunk_), total(k * (len / chunk_)) { back = total; }
    bool take_front(uint64_t* u) {
        std::lock_guard<std::mutex> lk(mu);
        if (front >= back) return false;
        *u = front++;
        return true;
    }
    bool take_back(uint64_t* u) {
        std::lock_guard<std::mutex> lk(mu);
        if (front >= back) return false;
        *u = --back;
        return true;
    }
};

struct PackStats {
    uint64_t packed_units = 0, raw_units = 0;
};

// Back end B:
//   int  thread_enter();                                        per-thread set-up (device selection); 0 = ok
//   uint32_t* stage(int worker, int slot);                      pinned staging buffer (4 bytes per chunk entry, 8-byte aligned);
//                                                               holds uint32 entries or, with wire21, ceil(chunk/3) words
//   int  stage_wait(int worker, int slot);                      the copy queued from that buffer has completed
//   int  submit_packed(int worker, int slot, uint32_t table, uint64_t off, uint64_t n);
//   int  raw_wait(int slot);                                    the device buffer of that slot is free again
//   int  submit_raw(int slot, uint32_t table, uint64_t off, uint64_t n, const uint64_t* src);
// Every call returns 0 or an error code, which stops all lanes and becomes the return value (-2: no thread could be
// started).  Nothing is thrown.
template <class B>
int run_pack_upload(B& be, const uint64_t* const* tables, uint32_t k, uint64_t len, uint64_t chunk, int workers, int raw_slots,
                    uint64_t* or_acc, PackStats* stats, uint64_t pm1, bool streaming_stores = false, bool wire21 = false, uint32_t pf_bytes = 0) {
    PackUnits units(k, len, chunk);
    std::atomic<int> err{0};
    std::atomic<uint64_t> acc{0}, n_packed{0}, n_raw{0};
    auto pack_lane = [&](int w) {
        int rc = be.thread_enter();
        uint64_t u, a = 0, cnt = 0;
        int slot = 0;
        while (rc == 0 && err.load(std::memory_order_relaxed) == 0 && units.take_front(&u)) {
            const uint32_t t = (uint32_t)(u / units.per_table);
            const uint64_t off = (u % units.per_table) * chunk;
            rc = be.stage_wait(w, slot);
            if (rc != 0) break;
            const uint64_t* src = tables[t] + off;
            if (wire21) a |= streaming_stores ? pack21_host<true>(src, (uint64_t*)be.stage(w, slot), chunk, pm1, pf_bytes) : pack21_host<false>(src, (uint64_t*)be.stage(w, slot), chunk, pm1, pf_bytes);
            else a |= streaming_stores ? pack32_host_nt(src, be.stage(w, slot), chunk, pm1) : pack32_host(src, be.stage(w, slot), chunk, pm1, pf_bytes);
            rc = be.submit_packed(w, slot, t, off, chunk);
            slot ^= 1;
            ++cnt;
        }
        if (rc != 0) err.store(rc);
        acc.fetch_or(a);
        n_packed.fetch_add(cnt);
    };
    auto raw_lane = [&]() {
        int rc = be.thread_enter();
        uint64_t u, cnt = 0;
        int slot = 0;
        while (rc == 0 && err.load(std::memory_order_relaxed) == 0) {
            rc = be.raw_wait(slot);  // before taking a unit: a chunk is only claimed once it can be queued
            if (rc != 0 || !units.take_back(&u)) break;
            const uint32_t t = (uint32_t)(u / units.per_table);
            const uint64_t off = (u % units.per_table) * chunk;
            rc = be.submit_raw(slot, t, off, chunk, tables[t] + off);
            slot = (slot + 1) % raw_slots;
            ++cnt;
        }
        if (rc != 0) err.store(rc);
        n_raw.fetch_add(cnt);
    };
    std::vector<std::thread> th;
    th.reserve((size_t)workers + 1);
    try {
        for (int w = 0; w < workers; ++w) th.emplace_back(pack_lane, w);
        if (raw_slots > 0) th.emplace_back(raw_lane);
    } catch (...) {  // no more threads to be had: the lanes that did start finish the list
        if (th.empty()) err.store(-2);
    }
    for (auto& t : th) t.join();
    *or_acc = acc.load();
    if (stats) {
        stats->packed_units = n_packed.load();
        stats->raw_units = n_raw.load();
    }
    return err.load();
}

// memcpy back end: the "device" is host memory (self-test of the scheduler, no CUDA involved)
struct MemcpyPackBackend {
    uint64_t chunk;
    std::vector<std::vector<uint32_t>> staging;  // [worker * 2 + slot]
    std::vector<std::vector<uint32_t>>* dst;     // [table]
    std::atomic<uint64_t> raw_or{0};
    bool wire21 = false;
    uint64_t pm1 = ~0ull >> 1;
    MemcpyPackBackend(uint64_t chunk_, int workers, std::vector<std::vector<uint32_t>>* dst_) : chunk(chunk_), dst(dst_) {
        staging.resize((size_t)workers * 2);
        for (auto& s : staging) s.resize(chunk_ + 4);
    }
    int thread_enter() { return 0; }
    uint32_t* stage(int w, int slot) {  // 16-byte aligned inside the vector's allocation
        uint32_t* p = staging[(size_t)w * 2 + slot].data();
        return (uint32_t*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
    }
    int stage_wait(int, int) { return 0; }
    int submit_packed(int w, int slot, uint32_t t, uint64_t off, uint64_t n) {
        if (wire21) unpack21_host((const uint64_t*)stage(w, slot), (*dst)[t].data() + off, n);
        else std::copy(stage(w, slot), stage(w, slot) + n, (*dst)[t].begin() + off);
        return 0;
    }
    int raw_wait(int) { return 0; }
    int submit_raw(int, uint32_t t, uint64_t off, uint64_t n, const uint64_t* src) {
        raw_or.fetch_or(pack32_host(src, (*dst)[t].data() + off, n, pm1));
        return 0;
    }
};

}  // namespace scb
