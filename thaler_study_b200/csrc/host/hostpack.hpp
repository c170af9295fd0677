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

// narrow n 8-byte entries to 4 bytes; returns the OR of all entries (the caller checks the bits above the modulus)
static inline uint64_t pack32_host(const uint64_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n) {
    uint64_t acc = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t v = src[i];
        acc |= v;
        dst[i] = (uint32_t)v;
    }
    return acc;
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
//   uint32_t* stage(int worker, int slot);                      pinned staging buffer (chunk entries)
//   int  stage_wait(int worker, int slot);                      the copy queued from that buffer has completed
//   int  submit_packed(int worker, int slot, uint32_t table, uint64_t off, uint64_t n);
//   int  raw_wait(int slot);                                    the device buffer of that slot is free again
//   int  submit_raw(int slot, uint32_t table, uint64_t off, uint64_t n, const uint64_t* src);
// Every call returns 0 or an error code, which stops all lanes and becomes the return value.
template <class B>
int run_pack_upload(B& be, const uint64_t* const* tables, uint32_t k, uint64_t len, uint64_t chunk, int workers, int raw_slots,
                    uint64_t* or_acc, PackStats* stats) {
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
            a |= pack32_host(tables[t] + off, be.stage(w, slot), chunk);
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
    for (int w = 0; w < workers; ++w) th.emplace_back(pack_lane, w);
    if (raw_slots > 0) th.emplace_back(raw_lane);
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
    MemcpyPackBackend(uint64_t chunk_, int workers, std::vector<std::vector<uint32_t>>* dst_) : chunk(chunk_), dst(dst_) {
        staging.resize((size_t)workers * 2);
        for (auto& s : staging) s.resize(chunk_);
    }
    int thread_enter() { return 0; }
    uint32_t* stage(int w, int slot) { return staging[(size_t)w * 2 + slot].data(); }
    int stage_wait(int, int) { return 0; }
    int submit_packed(int w, int slot, uint32_t t, uint64_t off, uint64_t n) {
        std::copy(stage(w, slot), stage(w, slot) + n, (*dst)[t].begin() + off);
        return 0;
    }
    int raw_wait(int) { return 0; }
    int submit_raw(int, uint32_t t, uint64_t off, uint64_t n, const uint64_t* src) {
        raw_or.fetch_or(pack32_host(src, (*dst)[t].data() + off, n));
        return 0;
    }
};

}  // namespace scb
