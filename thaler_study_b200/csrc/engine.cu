// engine.cu -- level 1 of the C ABI (include/sumcheck_b200.h): fields, HBM-resident dense MLEs and
// the SumCheckPolynomial implementors, on top of the kernels in kernels.cuh.
//
// There is NO CPU fallback in this file: every compute entry point needs a CUDA device and returns
// SCB_ECUDA without one.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "internal.hpp"
#include "options.hpp"
#include "kernels.cuh"
#include "eqfix.cuh"
#include "gkr.cuh"
#include "packed.cuh"
#include "tail.cuh"
#include "persist.cuh"
#include "pairs.cuh"
#include "g4_launch.hpp"
#include "tri.cuh"
#include "sumcheck_b200.h"

using namespace scb;

// ------------------------------------------------------------------------------------------ errors
namespace scb {
static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
}  // namespace scb

#define CU_TRY(expr)                                                                          \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SCB_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)
#define ARG_TRY(cond, msg)        \
    do {                          \
        if (!(cond)) {            \
            set_error("%s", msg); \
            return SCB_EINVAL;    \
        }                         \
    } while (0)
#define RC_TRY(expr)                 \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != SCB_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------ context
static thread_local cudaStream_t g_stream = 0;
static std::atomic<uint64_t> g_launches{0};

struct Ctx {
    int dev = -1;
    int sms = 0;
    uint64_t* partials = nullptr;  // per-block partial sums
    unsigned int* ticket = nullptr;
    uint64_t* h_res = nullptr;     // mapped pinned host memory: kernels write round sums here directly
    uint64_t* h_msgs = nullptr;    // mapped pinned host memory, 64 KB: message sums of a whole batch of rounds (GKR layer)
    uint64_t* d_scratch = nullptr; // small device scratch (points, results)
    TailMailbox* mailbox = nullptr; // mapped pinned host memory shared with the persistent tail kernel
    PersistCtl* persist_ctl = nullptr; // device memory: barrier state of the grid-wide resident kernel
    unsigned int* d_flag = nullptr; // device word raised by k_check_canonical
    int coop = 0;                   // cooperative launches supported
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // bracket the resident kernels (scb_resident_stats)
};
static constexpr int kMaxGrid = 148 * 16;
static constexpr int kMaxDev = 16;
static Ctx g_ctx[kMaxDev];
static std::mutex g_ctx_mu;

static int get_ctx(Ctx** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device available: %s (this library has no CPU fallback)", cudaGetErrorString(e));
        return SCB_ECUDA;
    }
    ARG_TRY(dev >= 0 && dev < kMaxDev, "device index out of range");
    Ctx& c = g_ctx[dev];
    if (c.dev < 0) {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        if (c.dev < 0) {
            int sms = 0;
            CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            cudaMemPool_t pool;
            CU_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
            uint64_t thr = UINT64_MAX;  // keep freed table buffers cached in the pool (ping-pong folds)
            CU_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
            CU_TRY(cudaMalloc(&c.partials, (size_t)kMaxGrid * 2 * kMaxPts * kMaxLimbs * 8));  // also (K+1)^2 grid sums per block
            CU_TRY(cudaMalloc(&c.ticket, 64));
            CU_TRY(cudaMemset(c.ticket, 0, 64));
            CU_TRY(cudaHostAlloc(&c.h_res, 4096, cudaHostAllocMapped | cudaHostAllocPortable));
            CU_TRY(cudaHostAlloc(&c.h_msgs, 65536, cudaHostAllocMapped | cudaHostAllocPortable));
            CU_TRY(cudaMalloc(&c.d_scratch, 64 * 1024));
            CU_TRY(cudaHostAlloc(&c.mailbox, sizeof(TailMailbox), cudaHostAllocMapped | cudaHostAllocPortable));
            std::memset((void*)c.mailbox, 0, sizeof(TailMailbox));
            CU_TRY(cudaMalloc(&c.persist_ctl, sizeof(PersistCtl)));
            CU_TRY(cudaMalloc(&c.d_flag, 64));
            CU_TRY(cudaDeviceGetAttribute(&c.coop, cudaDevAttrCooperativeLaunch, dev));
            CU_TRY(cudaEventCreate(&c.ev0));
            CU_TRY(cudaEventCreate(&c.ev1));
            c.sms = sms;
            c.dev = dev;
        }
    }
    *out = &c;
    return SCB_OK;
}

static inline int grid_for(const Ctx* c, uint64_t items, int blocks_per_sm = 8) {
    uint64_t want = (items + kThreads - 1) / kThreads;
    uint64_t cap = (uint64_t)c->sms * blocks_per_sm;
    if (cap > (uint64_t)kMaxGrid) cap = kMaxGrid;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// Grid for a grid-stride kernel: exactly one resident wave (SMs x active blocks per SM, from the occupancy
// calculator, cached per kernel) so that no partial second wave runs at reduced occupancy.
#include <unordered_map>
static std::unordered_map<const void*, int> g_occ;
static std::mutex g_occ_mu;
// pref_bps > 0 overrides the blocks-per-SM count (measured sweet spots, profiles/r01_kernel_sweep.md):
// the fused fold+message kernel streams 2 reads : 1 write and peaks at 5 CTAs/SM, the read-only message
// kernel keeps improving up to two full waves.
template <class KernelT>
static int occ_grid(const Ctx* c, KernelT kernel, uint64_t items, size_t dyn_smem = 0, int pref_bps = 0) {
    int nb = 0;
    {
        std::lock_guard<std::mutex> lk(g_occ_mu);
        auto it = g_occ.find((const void*)kernel);
        if (it == g_occ.end()) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, dyn_smem) != cudaSuccess || nb < 1) nb = 1;
            g_occ[(const void*)kernel] = nb;
        } else {
            nb = it->second;
        }
    }
    if (pref_bps > 0) nb = pref_bps;
    const int bps_env = (int)opt(OPT_bps);
    if (bps_env > 0) nb = bps_env;
    uint64_t want = (items + kThreads - 1) / kThreads;
    uint64_t cap = (uint64_t)c->sms * nb;
    if (cap > (uint64_t)kMaxGrid) cap = kMaxGrid;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}
// the same for kernels whose dynamic shared memory varies from call to call (no cache)
template <class KernelT>
static int occ_grid_smem(const Ctx* c, KernelT kernel, uint64_t items, size_t dyn_smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, dyn_smem) != cudaSuccess || nb < 1) nb = 1;
    uint64_t want = (items + kThreads - 1) / kThreads;
    uint64_t cap = (uint64_t)c->sms * nb;
    if (cap > (uint64_t)kMaxGrid) cap = kMaxGrid;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}
// opt in to more than 48 KB of dynamic shared memory (once per kernel)
template <class KernelT>
static int allow_smem(KernelT kernel, size_t bytes) {
    if (bytes <= 32 * 1024) return SCB_OK;  // static + dynamic must stay under 48 KB without the opt-in
    static std::unordered_map<const void*, size_t> done;
    std::lock_guard<std::mutex> lk(g_occ_mu);
    auto it = done.find((const void*)kernel);
    if (it != done.end() && it->second >= bytes) return SCB_OK;
    CU_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done[(const void*)kernel] = bytes;
    return SCB_OK;
}
#define LAUNCH_CHECK()                                                                      \
    do {                                                                                    \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) {                                                           \
            set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SCB_ECUDA;                                                               \
        }                                                                                   \
    } while (0)


// ------------------------------------------------------------------------------------------ peers (multi-GPU)
struct scb_peers {
    uint32_t rank = 0, world = 1;
    uint64_t* win[kMaxRanks] = {nullptr};  // win[rank] = own window (cudaMalloc), others = IPC mappings
    size_t gather_bytes = 0;               // capacity of ONE gather buffer (all ranks' slabs of all tables)
    uint64_t seq = 0, gseq = 0;
    bool connected = false;
};
static thread_local scb_peers* g_cur_peers = nullptr;  // set around the round calls of a sharded prover
static constexpr size_t kStatusWord = 500;              // word of Ctx::h_res used as exchange status

static int peers_check(Ctx* c);
static PeerArg peer_arg(Ctx* c) {
    PeerArg a;
    std::memset(&a, 0, sizeof a);
    scb_peers* p = g_cur_peers;
    if (p && p->world > 1) {
        for (uint32_t g = 0; g < p->world; ++g) a.win[g] = p->win[g];
        a.rank = p->rank;
        a.world = p->world;
        a.seq = ++p->seq;
        a.status = c->h_res + kStatusWord;
        a.timeout_ns = 10ull * 1000 * 1000 * 1000;
    }
    return a;
}

// policy dispatch: binds `A` to the arithmetic policy of the field
#define DISPATCH_POLICY(pol, ...)                              \
    switch (pol) {                                             \
        case POL_SP: { using A = PolSP; __VA_ARGS__; } break;   \
        case POL_G1: { using A = PolG1; __VA_ARGS__; } break;   \
        case POL_G4: { using A = PolGN<4>; __VA_ARGS__; } break; \
        default: set_error("unsupported field policy"); return SCB_EINVAL; \
    }
#define DISPATCH_K(kk, ...)                                   \
    switch (kk) {                                             \
        case 1: { constexpr int K = 1; __VA_ARGS__; } break;  \
        case 2: { constexpr int K = 2; __VA_ARGS__; } break;  \
        case 3: { constexpr int K = 3; __VA_ARGS__; } break;  \
        case 4: { constexpr int K = 4; __VA_ARGS__; } break;  \
        default: set_error("number of tables must be 1..4"); return SCB_EINVAL; \
    }

// ------------------------------------------------------------------------------------------ buffers
struct DevBuf {
    uint64_t* ptr = nullptr;
    size_t bytes = 0;
    bool owned = false;
    cudaStream_t stream = 0;
    ~DevBuf() {
        if (owned && ptr) cudaFreeAsync(ptr, stream);
    }
};
typedef std::shared_ptr<DevBuf> BufRef;

static int alloc_buf(size_t bytes, BufRef* out) {
    auto b = std::make_shared<DevBuf>();
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, bytes < 32 ? 32 : bytes, g_stream);
    if (e != cudaSuccess) {
        set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? SCB_ENOMEM : SCB_ECUDA;
    }
    b->ptr = (uint64_t*)p;
    b->bytes = bytes;
    b->owned = true;
    b->stream = g_stream;
    *out = b;
    return SCB_OK;
}

struct Table {
    uint32_t nv = 0;
    BufRef buf;
    bool p32 = false;  // packed uint32 entries (internal intermediate of the small-prime policy, packed.cuh)
    uint64_t len() const { return 1ull << nv; }
};

struct scb_mle {
    std::shared_ptr<FieldImpl> f;
    Table t;
};

struct scb_poly {
    uint32_t kind = 0;
    std::shared_ptr<FieldImpl> f;
    std::vector<Table> t;
    uint32_t var_len = 0;  // triangle_counting::G::var_len
    bool allow_packed = false;  // descendants produced by fix_and_round may keep packed uint32 tables
    // triangle_counting::G while an x variable is left: M[z][x] = sum_y f2[z][y] f1[y][x] (tri.cuh), index (z << xn) | x
    bool has_aux = false;
    Table aux;
    // pairs.cuh, 21-bit triples: word i = t[0][i] | t[1][i] << 21 | t[2][i] << 42, written by the grid pass of a prover that will
    // run the first pair pass as its own launch.  Valid only for the buffers it was built from (w21_ok): copies of this struct
    // with other tables simply ignore it.
    // The slot is never copied: a clone or a folded descendant starts without it, so the 2^v x 8 bytes die with the handle they
    // were built on (the prover's private copy) instead of being kept alive by every copy of the struct.
    struct W21Slot {
        BufRef buf;
        const uint64_t* src[3] = {nullptr, nullptr, nullptr};
        W21Slot() = default;
        W21Slot(const W21Slot&) {}
        W21Slot& operator=(const W21Slot&) { return *this; }
    };
    mutable W21Slot w21s;
    bool w21_ok() const {
        if (!w21s.buf || t.size() != 3) return false;
        for (int k = 0; k < 3; ++k)
            if (t[k].p32 || !t[k].buf || t[k].buf->ptr != w21s.src[k] || t[k].nv != t[0].nv) return false;
        return true;
    }
    bool any_packed() const {
        for (const Table& x : t)
            if (x.p32) return true;
        return false;
    }
};

// packed-table helpers (defined next to the fused kernels below)
static int unpack_table(Ctx* c, const FieldImpl& f, const Table& in, Table* out);
static int plain_poly(const scb_poly* p, std::unique_ptr<scb_poly>* holder, const scb_poly** out);
#define PLAIN_POLY(p)                             \
    std::unique_ptr<scb_poly> plain_holder__;     \
    RC_TRY(plain_poly(p, &plain_holder__, &p))

static ElemArg elem_arg(const FieldImpl& f, const uint64_t* w) {
    ElemArg a;
    for (int i = 0; i < kMaxLimbs; ++i) a.w[i] = i < (int)f.d.n ? w[i] : 0;
    return a;
}
static bool elem_canonical(const FieldImpl& f, const uint64_t* w) {
    Fe e;
    f.h.load(w, e);
    return f.h.is_canonical(e);
}

// ------------------------------------------------------------------------------------------ library
extern "C" const char* scb_last_error(void) { return g_err.c_str(); }
extern "C" const char* scb_version(void) { return "sumcheck_b200 0.1 (sm_100a)"; }
extern "C" int scb_device_count(int* out) {
    ARG_TRY(out, "null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *out = n;
    return SCB_OK;
}
extern "C" int scb_set_stream(void* s) {
    g_stream = (cudaStream_t)s;
    return SCB_OK;
}
extern "C" int scb_synchronize(void) {
    CU_TRY(cudaStreamSynchronize(g_stream));
    return SCB_OK;
}
extern "C" int scb_set_option(const char* name, int64_t value) {
    ARG_TRY(name, "null argument");
    const int i = option_index(name);
    if (i < 0) {
        set_error("unknown option '%s'", name);
        return SCB_EINVAL;
    }
    option_table().v[i].store(value);
    return SCB_OK;
}
extern "C" int scb_get_option(const char* name, int64_t* out) {
    ARG_TRY(name && out, "null argument");
    const int i = option_index(name);
    if (i < 0) {
        set_error("unknown option '%s'", name);
        return SCB_EINVAL;
    }
    *out = option_table().v[i].load();
    return SCB_OK;
}
extern "C" const char* scb_option_name(uint32_t index) { return option_name((int)index); }
extern "C" void scb_reset_options(void) { option_table().reset(); }
extern "C" int scb_launch_count(uint64_t* out, int reset) {
    ARG_TRY(out, "null out");
    *out = reset ? g_launches.exchange(0) : g_launches.load();
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------ field
extern "C" int scb_field_create(uint32_t n_limbs, const uint64_t* modulus, scb_field** out) {
    ARG_TRY(modulus && out, "null argument");
    ARG_TRY(n_limbs == 1 || n_limbs == 4, "n_limbs must be 1 or 4");
    auto impl = std::make_shared<FieldImpl>();
    try {
        impl->h = HostField(n_limbs, modulus);
    } catch (const std::exception& ex) {
        set_error("scb_field_create: %s", ex.what());
        return SCB_EINVAL;
    }
    std::memset(&impl->d, 0, sizeof impl->d);
    for (uint32_t i = 0; i < n_limbs; ++i) {
        impl->d.p[i] = impl->h.p[i];
        impl->d.one[i] = impl->h.one_.l[i];
        impl->d.r2[i] = impl->h.r2_.l[i];
    }
    impl->d.inv = impl->h.inv;
    impl->d.n = n_limbs;
    impl->d.bits = impl->h.bits;
    impl->policy = n_limbs == 4 ? POL_G4 : (impl->h.bits <= 28 ? POL_SP : POL_G1);
    *out = new scb_field{impl};
    return SCB_OK;
}
extern "C" void scb_field_free(scb_field* f) { delete f; }
extern "C" int scb_field_n_limbs(const scb_field* f, uint32_t* out) {
    ARG_TRY(f && out, "null argument");
    *out = f->impl->d.n;
    return SCB_OK;
}
extern "C" int scb_field_modulus_bits(const scb_field* f, uint32_t* out) {
    ARG_TRY(f && out, "null argument");
    *out = f->impl->d.bits;
    return SCB_OK;
}
extern "C" int scb_field_policy(const scb_field* f, uint32_t* out) {
    ARG_TRY(f && out, "null argument");
    *out = f->impl->policy;
    return SCB_OK;
}
extern "C" int scb_field_to_mont(const scb_field* f, const uint64_t* canonical, uint64_t* mont, size_t count) {
    ARG_TRY(f && canonical && mont, "null argument");
    const HostField& h = f->impl->h;
    for (size_t i = 0; i < count; ++i) {
        Fe e;
        h.load(canonical + i * h.n, e);
        ARG_TRY(h.is_canonical(e), "scb_field_to_mont: value >= modulus");
        h.store(h.to_mont(e), mont + i * h.n);
    }
    return SCB_OK;
}
extern "C" int scb_field_from_mont(const scb_field* f, const uint64_t* mont, uint64_t* canonical, size_t count) {
    ARG_TRY(f && canonical && mont, "null argument");
    const HostField& h = f->impl->h;
    for (size_t i = 0; i < count; ++i) {
        Fe e;
        h.load(mont + i * h.n, e);
        h.store(h.from_mont(e), canonical + i * h.n);
    }
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------ MLE
static int mle_new(const std::shared_ptr<FieldImpl>& f, uint32_t nv, scb_mle** out, bool alloc) {
    ARG_TRY(nv <= 40, "num_vars too large");
    auto m = std::make_unique<scb_mle>();
    m->f = f;
    m->t.nv = nv;
    if (alloc) RC_TRY(alloc_buf((size_t)8 * f->d.n << nv, &m->t.buf));
    *out = m.release();
    return SCB_OK;
}

static int mle_upload_packed(Ctx* c, const FieldImpl& fi, uint32_t num_vars, const uint64_t* evals, Table* out, bool* done);  // upload_engine.inc
// every entry of a freshly uploaded table must be a canonical field element (< p): the lazy accumulators of the
// small-prime kernels size their head-room on it.  One read of the table on the device (upload_engine.inc); waits.
static int check_canonical_table(Ctx* c, const FieldImpl& fi, const uint64_t* d_ptr, uint64_t n);
extern "C" int scb_mle_from_host(const scb_field* f, uint32_t num_vars, const uint64_t* evals, scb_mle** out) {
    ARG_TRY(f && evals && out, "null argument");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    {
        // large small-prime tables cross PCIe narrowed and are widened on the device (upload_engine.inc)
        ARG_TRY(num_vars <= 40, "too many variables");
        Table t;
        bool done = false;
        RC_TRY(mle_upload_packed(c, *f->impl, num_vars, evals, &t, &done));
        if (done) {
            CU_TRY(cudaStreamSynchronize(g_stream));
            auto mm = std::make_unique<scb_mle>();
            mm->f = f->impl;
            mm->t = t;
            *out = mm.release();
            return SCB_OK;
        }
    }
    scb_mle* m = nullptr;
    RC_TRY(mle_new(f->impl, num_vars, &m, true));
    cudaError_t e = cudaMemcpyAsync(m->t.buf->ptr, evals, m->t.buf->bytes, cudaMemcpyHostToDevice, g_stream);
    if (e != cudaSuccess) {
        delete m;
        set_error("H2D copy failed: %s", cudaGetErrorString(e));
        return SCB_ECUDA;
    }
    const int rc = check_canonical_table(c, *f->impl, m->t.buf->ptr, m->t.len());  // waits: the caller may reuse `evals` on return
    if (rc != SCB_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return SCB_OK;
}
extern "C" int scb_mle_from_device(const scb_field* f, uint32_t num_vars, const uint64_t* d_evals, int copy, scb_mle** out) {
    ARG_TRY(f && d_evals && out, "null argument");
    ARG_TRY(((uintptr_t)d_evals & 31) == 0, "device table must be 32-byte aligned");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    scb_mle* m = nullptr;
    RC_TRY(mle_new(f->impl, num_vars, &m, copy != 0));
    if (copy) {
        cudaError_t e = cudaMemcpyAsync(m->t.buf->ptr, d_evals, m->t.buf->bytes, cudaMemcpyDeviceToDevice, g_stream);
        if (e != cudaSuccess) {
            delete m;
            set_error("D2D copy failed: %s", cudaGetErrorString(e));
            return SCB_ECUDA;
        }
    } else {
        m->t.buf = std::make_shared<DevBuf>();
        m->t.buf->ptr = const_cast<uint64_t*>(d_evals);
        m->t.buf->bytes = (size_t)8 * f->impl->d.n << num_vars;
        m->t.buf->owned = false;
    }
    *out = m;
    return SCB_OK;
}
extern "C" int scb_mle_synthetic(const scb_field* f, uint32_t num_vars, uint64_t seed, uint64_t start, scb_mle** out) {
    ARG_TRY(f && out, "null argument");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    scb_mle* m = nullptr;
    RC_TRY(mle_new(f->impl, num_vars, &m, true));
    const uint64_t n = 1ull << num_vars;
    if (f->impl->d.n == 1)
        k_synth_fill<1><<<grid_for(c, n), kThreads, 0, g_stream>>>(f->impl->d, seed, start, n, m->t.buf->ptr);
    else
        k_synth_fill<4><<<grid_for(c, n), kThreads, 0, g_stream>>>(f->impl->d, seed, start, n, m->t.buf->ptr);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        delete m;
        set_error("k_synth_fill launch failed: %s", cudaGetErrorString(e));
        return SCB_ECUDA;
    }
    *out = m;
    return SCB_OK;
}
extern "C" int scb_mle_clone(const scb_mle* m, scb_mle** out) {
    ARG_TRY(m && out, "null argument");
    *out = new scb_mle(*m);
    return SCB_OK;
}
extern "C" void scb_mle_free(scb_mle* m) { delete m; }
extern "C" int scb_mle_num_vars(const scb_mle* m, uint32_t* out) {
    ARG_TRY(m && out, "null argument");
    *out = m->t.nv;
    return SCB_OK;
}
extern "C" int scb_mle_device_ptr(const scb_mle* m, const uint64_t** out) {
    ARG_TRY(m && out, "null argument");
    *out = m->t.buf->ptr;
    return SCB_OK;
}

// one fold pass of one table (a4)
static int fold_table(Ctx* c, const FieldImpl& f, const Table& in, const uint64_t* r, Table* out) {
    ARG_TRY(in.nv >= 1, "cannot fix a variable of a 0-variable table");
    ARG_TRY(elem_canonical(f, r), "challenge is not a canonical field element");
    Table o;
    o.nv = in.nv - 1;
    RC_TRY(alloc_buf((size_t)8 * f.d.n << o.nv, &o.buf));
    const uint64_t n_out = o.len();
    const ElemArg ra = elem_arg(f, r);
    DISPATCH_POLICY(f.policy, {
        if (A::N == 1 && n_out >= 2) {
            k_fold<A, (A::N == 1 ? 2 : 1)><<<grid_for(c, n_out / 2), kThreads, 0, g_stream>>>(f.d, in.buf->ptr, o.buf->ptr, ra, n_out / 2);
        } else {
            k_fold<A, 1><<<grid_for(c, n_out), kThreads, 0, g_stream>>>(f.d, in.buf->ptr, o.buf->ptr, ra, n_out);
        }
    });
    LAUNCH_CHECK();
    *out = o;
    return SCB_OK;
}
static int build_eq_tables(Ctx* c, const FieldImpl& f, const uint64_t* bitpt, uint32_t lb, uint32_t v, BufRef* lo, BufRef* hi);
// Several variables at once: out[j] = sum_i eq(point; i) * t[..] in ONE pass over the table (eqfix.cuh).
// high == false: the LOW n variables ([ARK] fix_variables(point)); high == true: the TOP n variables, point[t] bound
// to variable nv-n+t (= relabel + fix_variables, matrix-multiplication/src/lib.rs:82-83).
static int fix_table_eq(Ctx* c, const FieldImpl& f, const Table& in, const uint64_t* point, uint32_t n, bool high, Table* out) {
    for (uint32_t j = 0; j < n; ++j) ARG_TRY(elem_canonical(f, point + (size_t)j * f.d.n), "challenge is not a canonical field element");
    BufRef eq, unused;
    RC_TRY(build_eq_tables(c, f, point, n, n, &eq, &unused));
    Table o;
    o.nv = in.nv - n;
    RC_TRY(alloc_buf((size_t)8 * f.d.n << o.nv, &o.buf));
    const uint64_t n_out = o.len();
    if (high) {
        // enough (j, row-slice) threads to fill the machine; slices are summed by a second tiny kernel
        uint32_t n_slices = 1;
        while (n_slices < (1u << n) && n_out * n_slices < (uint64_t)c->sms * 2048) n_slices *= 2;
        BufRef scratch;
        RC_TRY(alloc_buf((size_t)8 * f.d.n * n_out * n_slices, &scratch));
        DISPATCH_POLICY(f.policy, {
            k_fix_high_eq<A><<<grid_for(c, n_out * n_slices), kThreads, 0, g_stream>>>(f.d, in.buf->ptr, eq->ptr, n, scratch->ptr, n_out, n_slices);
            LAUNCH_CHECK();
            k_fix_high_finish<A><<<grid_for(c, n_out), kThreads, 0, g_stream>>>(f.d, scratch->ptr, o.buf->ptr, n_out, n_slices);
        });
    } else {
        DISPATCH_POLICY(f.policy, { k_fix_low_eq<A><<<grid_for(c, n_out * 32), kThreads, 0, g_stream>>>(f.d, in.buf->ptr, eq->ptr, n, o.buf->ptr, n_out); });
    }
    LAUNCH_CHECK();
    *out = o;
    return SCB_OK;
}
static uint32_t eq_fix_max_vars(const FieldImpl& f) { return f.d.n == 1 ? 12 : 10; }

static int fix_table(Ctx* c, const FieldImpl& f, const Table& in, const uint64_t* point, uint32_t n, Table* out) {
    ARG_TRY(n <= in.nv, "invalid size of partial point");  // [ARK] assert in fix_variables
    if (n >= 2 && n <= eq_fix_max_vars(f) && !in.p32) return fix_table_eq(c, f, in, point, n, false, out);
    Table cur = in;
    for (uint32_t i = 0; i < n; ++i) {
        Table nxt;
        RC_TRY(fold_table(c, f, cur, point + (size_t)i * f.d.n, &nxt));
        cur = nxt;
    }
    *out = cur;
    return SCB_OK;
}

extern "C" int scb_mle_fix_variables(const scb_mle* m, const uint64_t* partial_point, uint32_t n_point, scb_mle** out) {
    ARG_TRY(m && out && (partial_point || n_point == 0), "null argument");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    Table t;
    RC_TRY(fix_table(c, *m->f, m->t, partial_point, n_point, &t));
    auto r = new scb_mle{m->f, t};
    *out = r;
    return SCB_OK;
}

// eq tables over index bits [0, lb) and [lb, v) for the coordinates bitpt[j] (host memory, Montgomery limbs),
// built by one launch with the point as a kernel argument (eqfix.cuh)
static int build_eq_tables(Ctx* c, const FieldImpl& f, const uint64_t* bitpt, uint32_t lb, uint32_t v, BufRef* lo, BufRef* hi) {
    const uint32_t N = f.d.n;
    ARG_TRY(v <= (uint32_t)kMaxPointCoords, "too many coordinates");
    RC_TRY(alloc_buf((size_t)8 * N << lb, lo));
    RC_TRY(alloc_buf((size_t)8 * N << (v - lb), hi));
    PointArg pa;
    std::memset(&pa, 0, sizeof pa);
    if (v) std::memcpy(pa.w, bitpt, (size_t)8 * N * v);
    const uint32_t cap_bits = N == 1 ? 12 : 10;
    const size_t smem = (size_t)8 * N << cap_bits;
    const bool split = opt(OPT_eq_split) != 0;
    const uint32_t hb = v - lb;
    if (split && lb <= 18 && hb <= 18 && v >= 2) {
        // sub-tables of at most 2^9 entries in shared memory, the last level spread over the grid (eqfix.cuh)
        const uint32_t blocks_lo = lb <= 10 ? 1 : std::min(32u, 1u << (lb - 10)), blocks_hi = hb <= 10 ? 1 : std::min(64u, 1u << (hb - 10));
        const size_t smem2 = (size_t)8 * N * ((1u << (lb / 2)) + (1u << (lb - lb / 2)) > (1u << (hb / 2)) + (1u << (hb - hb / 2))
                                                  ? (1u << (lb / 2)) + (1u << (lb - lb / 2))
                                                  : (1u << (hb / 2)) + (1u << (hb - hb / 2)));
        DISPATCH_POLICY(f.policy, { k_eq_tables_split<A><<<blocks_lo + blocks_hi, 1024, smem2, g_stream>>>(f.d, pa, lb, v, (*lo)->ptr, (*hi)->ptr, blocks_lo); });
    } else {
        DISPATCH_POLICY(f.policy, { k_eq_tables<A><<<2, 1024, smem, g_stream>>>(f.d, pa, lb, v, (*lo)->ptr, (*hi)->ptr, cap_bits); });
    }
    LAUNCH_CHECK();
    return SCB_OK;
}

// MLE evaluation through eq tables (K4).  bitpt[j] = coordinate bound to index bit j (host memory).
// Result goes to host (h_out) or, if d_out != nullptr, stays on the device.
// One launch: every CTA builds the eq sub-tables in shared memory, warps take rows, last CTA adds up (eqfix.cuh).
// v_total > t.nv: `t` is this rank's slab of a table sharded by its top variables, row0 its first row; the finishing
// thread then adds the peer GPUs' sums (current peer group).  res: mapped host or device memory.
static bool eval_fused_ok(const FieldImpl& f, const Table& t, uint32_t v_total) {
    const uint32_t lb = f.d.n == 1 ? 8 : 10;
    return opt(OPT_mle_fused) != 0 && !t.p32 && t.nv >= lb && v_total >= t.nv && v_total - lb <= 27 && v_total <= (uint32_t)kMaxPointCoords;
}
static int eval_table_fused(Ctx* c, const FieldImpl& f, const Table& t, const uint64_t* bitpt, uint32_t v_total, uint64_t row0, uint64_t* res) {
    const uint32_t N = f.d.n;
    PointArg pa;
    std::memset(&pa, 0, sizeof pa);
    std::memcpy(pa.w, bitpt, (size_t)8 * N * v_total);
    if (f.policy == POL_G4 && opt(OPT_g4_kernel) == 3 && f.d.bits <= 255) {
        // 4-limb fields: the same launch with unreduced products in 544-bit register accumulators (g4_mle.cuh)
        const cudaError_t le = launch_mle_eval_fused_g4(opt(OPT_g4_p0one) != 0 && g4_p0one(f.d), c->sms, g_stream, f.d, pa, t.buf->ptr, t.nv, v_total, row0,
                                                        c->partials, c->ticket, res, peer_arg(c), kMaxGrid);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (le != cudaSuccess) {
            set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(le), __FILE__, __LINE__);
            return SCB_ECUDA;
        }
        return SCB_OK;
    }
    DISPATCH_POLICY(f.policy, {
        auto kern = k_mle_eval_fused<A>;
        constexpr size_t smem = MleFusedCfg<A>::smem_bytes;
        RC_TRY(allow_smem(kern, smem));
        const uint64_t n_rows = t.len() >> MleFusedCfg<A>::LB;
        // one-limb fields up to 2^25 entries: three CTAs per SM instead of the five that fit -- fewer table builds and partial sums on the
        // latency-bound path (2^24: 50.6 -> 45.6 us per call, scripts/kbench_mle_bps.py); larger tables keep the full wave for bandwidth
        const int pref = (A::N == 1 && t.nv <= 25) ? 3 : 0;
        kern<<<occ_grid(c, kern, n_rows * 32, smem, pref), kThreads, smem, g_stream>>>(f.d, pa, t.buf->ptr, t.nv, v_total, row0, c->partials, c->ticket, res,
                                                                             peer_arg(c));
    });
    LAUNCH_CHECK();
    return SCB_OK;
}

static int eval_table(Ctx* c, const FieldImpl& f, const Table& t, const uint64_t* bitpt, uint64_t* h_out, uint64_t* d_out) {
    const uint32_t v = t.nv, N = f.d.n;
    ARG_TRY(v <= 34, "table too large for MLE evaluation");
    for (uint32_t j = 0; j < v; ++j) ARG_TRY(elem_canonical(f, bitpt + (size_t)j * N), "point coordinate is not canonical");
    if (eval_fused_ok(f, t, v) && !(g_cur_peers && g_cur_peers->world > 1)) {
        RC_TRY(eval_table_fused(c, f, t, bitpt, v, 0, d_out ? d_out : c->h_res));
        if (!d_out) {
            CU_TRY(cudaStreamSynchronize(g_stream));
            std::memcpy(h_out, c->h_res, 8 * N);
        }
        return SCB_OK;
    }
    // low table (shared memory, copied by every CTA) over lb index bits, high table (L2) over the rest
    const uint32_t lb_env = (uint32_t)opt(OPT_mle_lb);
    const int u_env = (int)opt(OPT_mle_u);
    uint32_t lb_max = N == 1 ? 12 : 10;
    if (lb_env >= 2 && lb_env < lb_max) lb_max = lb_env;
    const uint32_t lb = v < lb_max ? v : lb_max;
    BufRef lo, hi;
    RC_TRY(build_eq_tables(c, f, bitpt, lb, v, &lo, &hi));
    uint64_t* res = d_out ? d_out : c->h_res;
    DISPATCH_POLICY(f.policy, {
        const size_t smem = (size_t)8 * N << lb;
        const uint64_t n = t.len();
        if (lb >= 2) {  // four entries of a row per group: 1.25 multiplications per entry
            if (A::N == 1 && u_env != 1) {
                auto kern = k_mle_dot<A, 4, (A::N == 1 ? 4 : 1)>;  // four 256-bit loads in flight per thread
                kern<<<occ_grid_smem(c, kern, (n / 4 + 3) / 4, smem), kThreads, smem, g_stream>>>(f.d, t.buf->ptr, lo->ptr, hi->ptr, lb, n / 4, c->partials, c->ticket, res);
            } else {
                auto kern = k_mle_dot<A, 4, 1>;
                kern<<<occ_grid_smem(c, kern, n / 4, smem), kThreads, smem, g_stream>>>(f.d, t.buf->ptr, lo->ptr, hi->ptr, lb, n / 4, c->partials, c->ticket, res);
            }
        } else {
            auto kern = k_mle_dot<A, 1, 1>;
            kern<<<occ_grid_smem(c, kern, n, smem), kThreads, smem, g_stream>>>(f.d, t.buf->ptr, lo->ptr, hi->ptr, lb, n, c->partials, c->ticket, res);
        }
    });
    LAUNCH_CHECK();
    if (!d_out) {
        CU_TRY(cudaStreamSynchronize(g_stream));
        std::memcpy(h_out, c->h_res, 8 * N);
    }
    return SCB_OK;
}

// T evaluations of one table, points in host memory as pts[t][j][limb] (coordinate j bound to index bit j): one
// kernel builds all eq tables, then one pass over the table per chunk of TC points.  d_out (device or mapped host
// memory) receives T elements; nothing waits.  Falls back to T single evaluations for shapes the staged tables do not fit.
static int eval_table_multi(Ctx* c, const FieldImpl& f, const Table& t, const uint64_t* pts, uint32_t T, uint64_t* d_out) {
    const uint32_t v = t.nv, N = f.d.n;
    const uint32_t cap_bits = N == 1 ? 12 : 10;
    const uint32_t TC = N == 1 ? 8 : 2;
    uint32_t lb = v / 2;
    while (lb > 0 && ((size_t)TC * 8 * N << lb) > 64 * 1024) --lb;  // TC low tables in 64 KB of shared memory
    // one-limb fields, tables up to 2^(8 + cap_bits) entries: the row-wise kernel with 256-entry rows (k_mle_rows_multi)
    const bool rows = N == 1 && opt(OPT_mle_rows_multi) != 0 && v >= (uint32_t)kRowsMultiLB && v - kRowsMultiLB <= cap_bits && !t.p32;
    if (rows) lb = kRowsMultiLB;
    if (v < 4 || lb < 2 || v - lb > cap_bits || t.p32) {
        for (uint32_t i = 0; i < T; ++i) RC_TRY(eval_table(c, f, t, pts + (size_t)i * v * N, nullptr, d_out + (size_t)i * N));
        return SCB_OK;
    }
    for (size_t i = 0; i < (size_t)T * v; ++i) ARG_TRY(elem_canonical(f, pts + i * N), "point coordinate is not canonical");
    BufRef d_pts, lo, hi;
    RC_TRY(alloc_buf((size_t)T * v * N * 8, &d_pts));
    RC_TRY(alloc_buf(((size_t)T * 8 * N) << lb, &lo));
    RC_TRY(alloc_buf(((size_t)T * 8 * N) << (v - lb), &hi));
    CU_TRY(cudaMemcpyAsync(d_pts->ptr, pts, (size_t)T * v * N * 8, cudaMemcpyHostToDevice, g_stream));
    const size_t eq_smem = (size_t)8 * N << cap_bits;
    DISPATCH_POLICY(f.policy, {
        auto kern = k_eq_tables_multi<A>;
        RC_TRY(allow_smem(kern, eq_smem));
        kern<<<2 * T, 1024, eq_smem, g_stream>>>(f.d, d_pts->ptr, lb, v, lo->ptr, hi->ptr);
    });
    LAUNCH_CHECK();
    const uint64_t n = t.len();
    const uint32_t step = rows && T > 8 ? 24 : TC;  // the row-wise kernel also exists for 24 points per pass
    for (uint32_t t0 = 0; t0 < T; t0 += step) {
        const uint32_t np = T - t0 < step ? T - t0 : step;
        const size_t smem = ((size_t)np * 8 * N) << lb;
        if (rows) {
            const int grid = grid_for(c, (n >> kRowsMultiLB) * 32, (int)opt(OPT_mle_rows_bps));
            const uint64_t* lo_p = lo->ptr + (((size_t)t0 << lb) * N);
            const uint64_t* hi_p = hi->ptr + (((size_t)t0 << (v - lb)) * N);
            uint64_t* out_p = d_out + (size_t)t0 * N;
#define SCB_ROWS_MULTI(POL, TCC)                                                                                                        \
    {                                                                                                                                   \
        auto kern = k_mle_rows_multi<POL, TCC>;                                                                                         \
        RC_TRY(allow_smem(kern, (size_t)TCC * 8 << kRowsMultiLB));                                                                      \
        kern<<<grid, kThreads, smem, g_stream>>>(f.d, t.buf->ptr, lo_p, hi_p, v, np, n, c->partials, c->ticket, out_p);                 \
    }
            if (f.policy == POL_SP) {
                if (step == 24) SCB_ROWS_MULTI(PolSP, 24) else SCB_ROWS_MULTI(PolSP, 8)
            } else {
                if (step == 24) SCB_ROWS_MULTI(PolG1, 24) else SCB_ROWS_MULTI(PolG1, 8)
            }
#undef SCB_ROWS_MULTI
            LAUNCH_CHECK();
            continue;
        }
        DISPATCH_POLICY(f.policy, {
            constexpr int KTC = A::N == 1 ? 8 : 2;
            auto kern = k_mle_dot_multi<A, KTC>;
            RC_TRY(allow_smem(kern, ((size_t)KTC * 8 * N) << lb));
            // every CTA first stages the low tables (up to 64 KB, so at most 3 CTAs fit an SM): one resident wave
            kern<<<grid_for(c, n / 4, 3), kThreads, smem, g_stream>>>(f.d, t.buf->ptr, lo->ptr + (((size_t)t0 << lb) * N), hi->ptr + (((size_t)t0 << (v - lb)) * N),
                                                              lb, v, np, n, c->partials, c->ticket, d_out + (size_t)t0 * N);
        });
        LAUNCH_CHECK();
    }
    return SCB_OK;
}

extern "C" int scb_mle_evaluate_many(const scb_mle* m, const uint64_t* points, uint32_t n_point, uint32_t n_points, uint64_t* out_elems) {
    ARG_TRY(m && out_elems && points, "null argument");
    ARG_TRY(n_point == m->t.nv, "point dimension does not match num_vars");
    ARG_TRY(n_points >= 1 && n_points <= 64, "between 1 and 64 points per call");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    const uint32_t N = m->f->d.n;
    RC_TRY(eval_table_multi(c, *m->f, m->t, points, n_points, c->h_msgs));  // mapped host memory: results land there
    CU_TRY(cudaStreamSynchronize(g_stream));
    std::memcpy(out_elems, c->h_msgs, (size_t)8 * N * n_points);
    return SCB_OK;
}
extern "C" int scb_mle_evaluate(const scb_mle* m, const uint64_t* point, uint32_t n_point, uint64_t* out_elem) {
    ARG_TRY(m && out_elem && (point || n_point == 0), "null argument");
    ARG_TRY(n_point == m->t.nv, "point dimension does not match num_vars");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    return eval_table(c, *m->f, m->t, point, out_elem, nullptr);  // LSB-first: bit j <-> point[j]
}
static int eval_table_be(Ctx* c, const FieldImpl& f, const Table& t, const uint64_t* r, uint64_t* out) {
    const uint32_t v = t.nv, N = f.d.n;
    std::vector<uint64_t> bitpt((size_t)N * (v ? v : 1));
    for (uint32_t j = 0; j < v; ++j) std::memcpy(&bitpt[(size_t)j * N], r + (size_t)(v - 1 - j) * N, 8 * N);  // r[0] <-> MSB
    return eval_table(c, f, t, bitpt.data(), out, nullptr);
}
extern "C" int scb_mle_evaluate_be(const scb_mle* m, const uint64_t* r, uint32_t n_r, uint64_t* out_elem) {
    ARG_TRY(m && out_elem && (r || n_r == 0), "null argument");
    ARG_TRY(n_r == m->t.nv, "point dimension does not match num_vars");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    return eval_table_be(c, *m->f, m->t, r, out_elem);
}
// MLE evaluation of a table sharded by its top log2(world) variables ("each GPU dots its slab with its
// slice of the eq table; one exchange of E bytes"): rank g holds entries [g 2^lv, (g+1) 2^lv).  The point has
// lv + log2(world) coordinates; big_endian != 0: r[0] <-> index MSB (multilinear-extensions), else LSB-first ([ARK]).
// One launch per rank; the exchange and the modular sum happen in the kernel's finishing thread over NVLink peer
// memory.  Every rank returns the same element.
extern "C" int scb_mle_evaluate_sharded(const scb_mle* slab, scb_peers* peers, const uint64_t* point, uint32_t n_point, int big_endian,
                                        uint64_t* out_elem) {
    ARG_TRY(slab && peers && out_elem && point, "null argument");
    ARG_TRY(peers->connected || peers->world == 1, "scb_peers_connect has not been called");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    const FieldImpl& f = *slab->f;
    const uint32_t N = f.d.n, lv = slab->t.nv, lg = 31 - __builtin_clz(peers->world);
    ARG_TRY(n_point == lv + lg, "point dimension does not match num_vars + log2(world)");
    ARG_TRY(eval_fused_ok(f, slab->t, n_point), "slab too small (or table too large) for the sharded evaluation: gather it first");
    std::vector<uint64_t> bitpt((size_t)N * n_point);
    for (uint32_t j = 0; j < n_point; ++j) {
        const uint64_t* src = point + (size_t)(big_endian ? n_point - 1 - j : j) * N;
        ARG_TRY(elem_canonical(f, src), "point coordinate is not canonical");
        std::memcpy(&bitpt[(size_t)j * N], src, 8 * N);
    }
    const uint32_t lb = N == 1 ? 8 : 10;
    scb_peers* prev = g_cur_peers;
    RC_TRY(scb_peers_set_current(peers));
    int rc = eval_table_fused(c, f, slab->t, bitpt.data(), n_point, (uint64_t)peers->rank << (lv - lb), c->h_res);
    if (rc == SCB_OK && cudaStreamSynchronize(g_stream) != cudaSuccess) {
        set_error("sharded MLE evaluation failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = SCB_ECUDA;
    }
    if (rc == SCB_OK) rc = peers_check(c);
    g_cur_peers = prev;
    if (rc == SCB_OK) std::memcpy(out_elem, c->h_res, 8 * N);
    return rc;
}
extern "C" int scb_mle_relabel(const scb_mle* m, uint32_t a, uint32_t b, uint32_t k, scb_mle** out) {
    ARG_TRY(m && out, "null argument");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    if (a > b) std::swap(a, b);
    if (a == b || k == 0) return scb_mle_clone(m, out);
    ARG_TRY(b + k <= m->t.nv, "invalid relabel argument");
    ARG_TRY(a + k <= b, "overlapped swap window is not allowed");
    scb_mle* r = nullptr;
    RC_TRY(mle_new(m->f, m->t.nv, &r, true));
    const uint64_t n = m->t.len();
    if (m->f->d.n == 1)
        k_relabel<1><<<grid_for(c, n), kThreads, 0, g_stream>>>(m->t.buf->ptr, r->t.buf->ptr, n, a, b, k);
    else
        k_relabel<4><<<grid_for(c, n), kThreads, 0, g_stream>>>(m->t.buf->ptr, r->t.buf->ptr, n, a, b, k);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        delete r;
        set_error("k_relabel launch failed: %s", cudaGetErrorString(e));
        return SCB_ECUDA;
    }
    *out = r;
    return SCB_OK;
}
static int table_to_host(const FieldImpl& f, const Table& t, uint64_t* out, size_t cap_elems) {
    ARG_TRY(cap_elems >= t.len(), "output buffer too small");
    CU_TRY(cudaMemcpyAsync(out, t.buf->ptr, (size_t)8 * f.d.n << t.nv, cudaMemcpyDeviceToHost, g_stream));
    CU_TRY(cudaStreamSynchronize(g_stream));
    return SCB_OK;
}
extern "C" int scb_mle_copy_to_device(const scb_mle* m, uint64_t* d_out) {
    ARG_TRY(m && d_out, "null argument");
    CU_TRY(cudaMemcpyAsync(d_out, m->t.buf->ptr, (size_t)8 * m->f->d.n << m->t.nv, cudaMemcpyDeviceToDevice, g_stream));
    return SCB_OK;
}
extern "C" int scb_mle_to_evaluations(const scb_mle* m, uint64_t* out, size_t cap_elems) {
    ARG_TRY(m && out, "null argument");
    return table_to_host(*m->f, m->t, out, cap_elems);
}

// multilinear-extensions free functions: evals in host memory, r[0] <-> index MSB
static int mle_eval_host_be(const scb_field* f, const uint64_t* evals, size_t n_evals, const uint64_t* r, uint32_t n_r, uint64_t* out) {
    ARG_TRY(f && evals && out && (r || n_r == 0), "null argument");
    ARG_TRY(n_r <= 34 && n_evals == ((size_t)1 << n_r), "evals.len() must be 2^r.len()");
    scb_mle* m = nullptr;
    RC_TRY(scb_mle_from_host(f, n_r, evals, &m));
    int rc = scb_mle_evaluate_be(m, r, n_r, out);
    scb_mle_free(m);
    return rc;
}
extern "C" int scb_vsbw_multilinear_from_evaluations(const scb_field* f, const uint64_t* evals, size_t n_evals, const uint64_t* r,
                                                     uint32_t n_r, uint64_t* out_elem) {
    return mle_eval_host_be(f, evals, n_evals, r, n_r, out_elem);
}
extern "C" int scb_cti_multilinear_from_evaluations(const scb_field* f, const uint64_t* evals, size_t n_evals, const uint64_t* r,
                                                    uint32_t n_r, uint64_t* out_elem) {
    // same field element as the vsbw form (exact arithmetic); one kernel serves both
    return mle_eval_host_be(f, evals, n_evals, r, n_r, out_elem);
}

// ------------------------------------------------------------------------------------------ polys
static uint32_t tri_xn(const scb_poly* p) { return p->t[0].nv > p->var_len ? p->t[0].nv - p->var_len : 0; }   // :53-55
static uint32_t tri_yn(const scb_poly* p) { return p->t[1].nv > p->var_len ? p->t[1].nv - p->var_len : 0; }   // :57-59
static uint32_t tri_zn(const scb_poly* p) { return p->t[2].nv < p->var_len ? p->t[2].nv : p->var_len; }       // :61-67

static uint32_t poly_num_vars(const scb_poly* p) {
    switch (p->kind) {
        case SCB_POLY_TRIANGLE_G: return tri_xn(p) + tri_yn(p) + tri_zn(p);  // triangle-counting/src/lib.rs:134-136
        case SCB_POLY_GKR_W: return p->t[0].nv;                              // round_polynomial.rs:92-94
        default: return p->t[0].nv;                                          // matrix-multiplication/src/lib.rs:133-135
    }
}
static uint32_t poly_n_points(const scb_poly* p) {
    switch (p->kind) {
        case SCB_POLY_PRODUCT: return (uint32_t)p->t.size() + 1;
        default: return 3;  // round degree 2
    }
}

extern "C" int scb_poly_product(const scb_mle* const* tables, uint32_t k, scb_poly** out) {
    ARG_TRY(tables && out, "null argument");
    ARG_TRY(k >= 1 && k <= (uint32_t)kMaxTables, "number of tables must be 1..4");
    auto p = std::make_unique<scb_poly>();
    p->kind = SCB_POLY_PRODUCT;
    for (uint32_t i = 0; i < k; ++i) {
        ARG_TRY(tables[i], "null table");
        ARG_TRY(tables[i]->f.get() == tables[0]->f.get() || std::memcmp(&tables[i]->f->d, &tables[0]->f->d, sizeof(FieldDesc)) == 0,
                "tables are over different fields");
        ARG_TRY(tables[i]->t.nv == tables[0]->t.nv, "tables must have the same number of variables");
        p->t.push_back(tables[i]->t);
    }
    p->f = tables[0]->f;
    ARG_TRY((uint64_t)k < p->f->h.p[0] || p->f->d.n > 1, "field characteristic too small to interpolate degree-K messages");
    *out = p.release();
    return SCB_OK;
}
extern "C" int scb_poly_matmul_g(const scb_mle* f_a, const scb_mle* f_b, scb_poly** out) {
    ARG_TRY(f_a && f_b && out, "null argument");
    const scb_mle* tabs[2] = {f_a, f_b};
    RC_TRY(scb_poly_product(tabs, 2, out));
    (*out)->kind = SCB_POLY_MATMUL_G;
    return SCB_OK;
}
extern "C" int scb_poly_matmul_g_new(const scb_field* f, uint32_t n, const uint64_t* a, const uint64_t* b, const uint64_t* point,
                                     scb_poly** out) {
    // matrix-multiplication/src/lib.rs:77-92
    ARG_TRY(f && a && b && point && out, "null argument");
    ARG_TRY(n >= 1 && n <= 15, "n out of range");
    const uint32_t N = f->impl->d.n;
    scb_mle *ma = nullptr, *mar = nullptr, *maf = nullptr, *mb = nullptr, *mbf = nullptr;
    int rc = scb_mle_from_host(f, 2 * n, a, &ma);                              // :81
    if (rc == SCB_OK && n >= 2 && n <= eq_fix_max_vars(*f->impl)) {
        // :82-83 relabel(0, n, n) + fix_variables(&point[..n]) == fixing the HIGH variable block directly (one pass)
        Ctx* c = nullptr;
        rc = get_ctx(&c);
        Table t;
        if (rc == SCB_OK) rc = fix_table_eq(c, *ma->f, ma->t, point, n, true, &t);
        if (rc == SCB_OK) maf = new scb_mle{ma->f, t};
    } else {
        if (rc == SCB_OK) rc = scb_mle_relabel(ma, 0, n, n, &mar);             // :82
        if (rc == SCB_OK) rc = scb_mle_fix_variables(mar, point, n, &maf);     // :83
    }
    if (rc == SCB_OK) rc = scb_mle_from_host(f, 2 * n, b, &mb);                // :85
    if (rc == SCB_OK) rc = scb_mle_fix_variables(mb, point + (size_t)n * N, n, &mbf);  // :86
    if (rc == SCB_OK) rc = scb_poly_matmul_g(maf, mbf, out);
    scb_mle_free(ma);
    scb_mle_free(mar);
    scb_mle_free(maf);
    scb_mle_free(mb);
    scb_mle_free(mbf);
    return rc;
}
// M = f2 x f1 as matrices (tri.cuh): computed once per polynomial, folded along x with it afterwards
static int triangle_build_aux(Ctx* c, scb_poly* p) {
    const uint32_t xn = p->t[0].nv > p->var_len ? p->t[0].nv - p->var_len : 0;
    const uint32_t yn = p->t[1].nv > p->var_len ? p->t[1].nv - p->var_len : 0;
    const uint32_t zn = p->t[2].nv < p->var_len ? p->t[2].nv : p->var_len;
    p->has_aux = false;
    if (xn == 0 || xn + zn > 30) return SCB_OK;
    const FieldImpl& f = *p->f;
    Table m;
    m.nv = xn + zn;
    RC_TRY(alloc_buf((size_t)8 * f.d.n << m.nv, &m.buf));
    const uint32_t X = 1u << xn, Y = 1u << yn, Z = 1u << zn;
    if (f.policy == POL_SP) {
        const dim3 grid((X + kMmTile - 1) / kMmTile, (Z + kMmTile - 1) / kMmTile);
        k_field_matmul_sp<<<grid, 256, 0, g_stream>>>(f.d, p->t[0].buf->ptr, p->t[1].buf->ptr, m.buf->ptr, X, Y, Z);
    } else {
        const dim3 grid((X + 15) / 16, (Z + 15) / 16);
        DISPATCH_POLICY(f.policy, { k_field_matmul_gen<A><<<grid, 256, 0, g_stream>>>(f.d, p->t[0].buf->ptr, p->t[1].buf->ptr, m.buf->ptr, X, Y, Z); });
    }
    LAUNCH_CHECK();
    p->aux = m;
    p->has_aux = true;
    return SCB_OK;
}
extern "C" int scb_poly_triangle_g_new(const scb_field* f, uint32_t num_vars, const uint8_t* adj, scb_poly** out) {
    // triangle-counting/src/lib.rs:32-51
    ARG_TRY(f && adj && out, "null argument");
    ARG_TRY(num_vars <= 30, "num_vars out of range");
    const uint32_t N = f->impl->d.n;
    const size_t len = (size_t)1 << num_vars;
    std::vector<uint64_t> host(len * N, 0);
    for (size_t i = 0; i < len; ++i)
        if (adj[i]) std::memcpy(&host[i * N], f->impl->d.one, 8 * N);  // F::one() / F::zero()
    scb_mle* g = nullptr;
    RC_TRY(scb_mle_from_host(f, num_vars, host.data(), &g));
    auto p = std::make_unique<scb_poly>();
    p->kind = SCB_POLY_TRIANGLE_G;
    p->f = f->impl;
    p->t = {g->t, g->t, g->t};  // f_a_1, f_a_2, f_a_3 are clones of one table (:46-48)
    p->var_len = num_vars / 2;  // :44
    scb_mle_free(g);
    if (opt(OPT_tri_tiled) != 0) {
        Ctx* c;
        RC_TRY(get_ctx(&c));
        RC_TRY(triangle_build_aux(c, p.get()));
    }
    *out = p.release();
    return SCB_OK;
}
extern "C" int scb_poly_gkr_w(const scb_mle* add_i, const scb_mle* mul_i, const scb_mle* w_b, const scb_mle* w_c, scb_poly** out) {
    ARG_TRY(add_i && mul_i && w_b && w_c && out, "null argument");
    ARG_TRY(add_i->t.nv == mul_i->t.nv, "add_i and mul_i must have the same number of variables");
    ARG_TRY(add_i->t.nv == w_b->t.nv + w_c->t.nv, "add_i must range over (b, c)");
    ARG_TRY(std::memcmp(&add_i->f->d, &mul_i->f->d, sizeof(FieldDesc)) == 0 && std::memcmp(&add_i->f->d, &w_b->f->d, sizeof(FieldDesc)) == 0 &&
                std::memcmp(&add_i->f->d, &w_c->f->d, sizeof(FieldDesc)) == 0,
            "tables are over different fields");
    auto p = std::make_unique<scb_poly>();
    p->kind = SCB_POLY_GKR_W;
    p->f = add_i->f;
    p->t = {add_i->t, mul_i->t, w_b->t, w_c->t};
    *out = p.release();
    return SCB_OK;
}
extern "C" int scb_poly_clone(const scb_poly* p, scb_poly** out) {
    ARG_TRY(p && out, "null argument");
    *out = new scb_poly(*p);
    return SCB_OK;
}
extern "C" void scb_poly_free(scb_poly* p) { delete p; }
extern "C" int scb_poly_kind_of(const scb_poly* p, uint32_t* out) {
    ARG_TRY(p && out, "null argument");
    *out = p->kind;
    return SCB_OK;
}
extern "C" int scb_poly_n_tables(const scb_poly* p, uint32_t* out) {
    ARG_TRY(p && out, "null argument");
    *out = (uint32_t)p->t.size();
    return SCB_OK;
}
extern "C" int scb_poly_table(const scb_poly* p, uint32_t idx, scb_mle** out) {
    ARG_TRY(p && out, "null argument");
    PLAIN_POLY(p);
    ARG_TRY(idx < p->t.size(), "table index out of range");
    *out = new scb_mle{p->f, p->t[idx]};
    return SCB_OK;
}
extern "C" int scb_poly_n_points(const scb_poly* p, uint32_t* out) {
    ARG_TRY(p && out, "null argument");
    *out = poly_n_points(p);
    return SCB_OK;
}
extern "C" int scb_poly_num_vars(const scb_poly* p, uint32_t* out) {
    ARG_TRY(p && out, "null argument");
    *out = poly_num_vars(p);
    return SCB_OK;
}

// SumCheckPolynomial::evaluate
extern "C" int scb_poly_evaluate(const scb_poly* p, const uint64_t* point, uint32_t n_point, uint64_t* out_elem) {
    ARG_TRY(p && out_elem && (point || n_point == 0), "null argument");
    PLAIN_POLY(p);
    ARG_TRY(n_point == poly_num_vars(p), "point dimension does not match num_vars");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    const FieldImpl& f = *p->f;
    const HostField& h = f.h;
    const uint32_t N = f.d.n;
    auto ev = [&](const Table& t, const uint64_t* pt, Fe* out) -> int {
        uint64_t w[kMaxLimbs];
        RC_TRY(eval_table(c, f, t, pt, w, nullptr));
        h.load(w, *out);
        return SCB_OK;
    };
    Fe res;
    if (p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G) {
        // matrix-multiplication/src/lib.rs:96-101
        res = h.one();
        for (size_t k = 0; k < p->t.size(); ++k) {
            Fe e;
            RC_TRY(ev(p->t[k], point, &e));
            res = k == 0 ? e : h.mul(res, e);
        }
    } else if (p->kind == SCB_POLY_TRIANGLE_G) {
        // triangle-counting/src/lib.rs:71-87
        const uint32_t xn = tri_xn(p), yn = tri_yn(p), zn = tri_zn(p);
        std::vector<uint64_t> xz((size_t)(xn + zn + 1) * N);
        std::memcpy(xz.data(), point, (size_t)8 * N * xn);
        std::memcpy(xz.data() + (size_t)xn * N, point + (size_t)(xn + yn) * N, (size_t)8 * N * zn);
        Fe e1, e2, e3;
        RC_TRY(ev(p->t[0], point, &e1));                      // x_y_point = point[..xn+yn]
        RC_TRY(ev(p->t[1], point + (size_t)xn * N, &e2));      // y_z_point = point[xn..]
        RC_TRY(ev(p->t[2], xz.data(), &e3));
        res = h.mul(h.mul(e1, e3), e2);
    } else {
        // gkr-protocol/src/round_polynomial.rs:48-57
        const uint32_t bn = p->t[2].nv;
        Fe ea, em, eb, ec;
        RC_TRY(ev(p->t[0], point, &ea));
        RC_TRY(ev(p->t[1], point, &em));
        RC_TRY(ev(p->t[2], point, &eb));
        RC_TRY(ev(p->t[3], point + (size_t)bn * N, &ec));
        res = h.add(h.mul(ea, h.add(eb, ec)), h.mul(em, h.mul(eb, ec)));
    }
    h.store(res, out_elem);
    return SCB_OK;
}

// SumCheckPolynomial::fix_variables
extern "C" int scb_poly_fix_variables(const scb_poly* p, const uint64_t* pp, uint32_t n, scb_poly** out) {
    ARG_TRY(p && out && (pp || n == 0), "null argument");
    PLAIN_POLY(p);
    Ctx* c;
    RC_TRY(get_ctx(&c));
    const FieldImpl& f = *p->f;
    const uint32_t N = f.d.n;
    auto q = std::make_unique<scb_poly>(*p);
    if (p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G) {
        // matrix-multiplication/src/lib.rs:103-108
        for (size_t k = 0; k < p->t.size(); ++k) RC_TRY(fix_table(c, f, p->t[k], pp, n, &q->t[k]));
    } else if (p->kind == SCB_POLY_TRIANGLE_G) {
        // triangle-counting/src/lib.rs:89-118
        const uint32_t xn = tri_xn(p), yn = tri_yn(p);
        const uint32_t n_xy = n < xn + yn ? n : xn + yn;
        const uint32_t n_yz = xn <= n ? n - xn : 0;
        const uint32_t n_x = n < xn ? n : xn;
        const uint32_t n_z = xn + yn <= n ? n - (xn + yn) : 0;
        std::vector<uint64_t> xz((size_t)(n_x + n_z + 1) * N);
        std::memcpy(xz.data(), pp, (size_t)8 * N * n_x);
        if (n_z) std::memcpy(xz.data() + (size_t)n_x * N, pp + (size_t)(xn + yn) * N, (size_t)8 * N * n_z);
        RC_TRY(fix_table(c, f, p->t[0], pp, n_xy, &q->t[0]));
        RC_TRY(fix_table(c, f, p->t[1], pp + (size_t)(n_yz ? xn : 0) * N, n_yz, &q->t[1]));
        RC_TRY(fix_table(c, f, p->t[2], xz.data(), n_x + n_z, &q->t[2]));
        // M folds along x like a table; once no x variable is left it is no longer needed
        q->has_aux = false;
        q->aux = Table();
        if (p->has_aux && n < xn) {
            RC_TRY(fix_table(c, f, p->aux, pp, n, &q->aux));
            q->has_aux = true;
        }
    } else {
        // gkr-protocol/src/round_polynomial.rs:59-76
        const uint32_t bn = p->t[2].nv;
        const uint32_t n_b = n < bn ? n : bn;
        const uint32_t n_c = bn <= n ? n - bn : 0;
        RC_TRY(fix_table(c, f, p->t[0], pp, n, &q->t[0]));
        RC_TRY(fix_table(c, f, p->t[1], pp, n, &q->t[1]));
        RC_TRY(fix_table(c, f, p->t[2], pp, n_b, &q->t[2]));
        RC_TRY(fix_table(c, f, p->t[3], pp + (size_t)(n_c ? bn : 0) * N, n_c, &q->t[3]));
    }
    *out = q.release();
    return SCB_OK;
}

// launch the message kernel of `p`; result (np elements) to `res` (mapped host or device memory)
static int launch_round_evals(Ctx* c, const scb_poly* p, uint64_t* res) {
    const FieldImpl& f = *p->f;
    ARG_TRY(poly_num_vars(p) >= 1, "polynomial has no variables left");
    if (p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G) {
        const uint64_t n_pairs = p->t[0].len() / 2;
        ARG_TRY(f.policy != POL_SP || p->t[0].nv <= 32, "table too large for the small-prime path");
        DISPATCH_POLICY(f.policy, DISPATCH_K(p->t.size(), {
            TabsIn<K> in;
            for (int k = 0; k < K; ++k) in.p[k] = p->t[k].buf->ptr;
            if (A::N == 1 && n_pairs >= 2) {
                constexpr int PV = A::N == 1 ? 2 : 1;
                auto kern = k_round_evals<A, K, PV>;
                kern<<<occ_grid(c, kern, n_pairs / PV, 0, A::kLight ? 16 : 0), kThreads, 0, g_stream>>>(f.d, in, n_pairs / PV, c->partials, c->ticket, res, peer_arg(c));
            } else {
                auto kern = k_round_evals<A, K, 1>;
                kern<<<occ_grid(c, kern, n_pairs), kThreads, 0, g_stream>>>(f.d, in, n_pairs, c->partials, c->ticket, res, peer_arg(c));
            }
        }));
    } else if (p->kind == SCB_POLY_TRIANGLE_G) {
        const uint32_t xn = tri_xn(p), yn = tri_yn(p), zn = tri_zn(p);
        const uint64_t n_t = xn > 0 ? 1ull << (xn - 1 + zn) : (yn > 0 ? 1ull << (yn - 1 + zn) : 1ull << (zn - 1));
        if (p->has_aux && xn > 0) {
            // x phase in matrix form: the message of the product of the tables M and f3 over adjacent x pairs (tri.cuh)
            DISPATCH_POLICY(f.policy, {
                TabsIn<2> in;
                in.p[0] = p->aux.buf->ptr;
                in.p[1] = p->t[2].buf->ptr;
                auto kern = k_round_evals<A, 2, 1>;
                kern<<<occ_grid(c, kern, n_t), kThreads, 0, g_stream>>>(f.d, in, n_t, c->partials, c->ticket, res, PeerArg{});
            });
        } else
        DISPATCH_POLICY(f.policy, {
            k_triangle_round<A><<<grid_for(c, n_t), kThreads, 0, g_stream>>>(f.d, p->t[0].buf->ptr, p->t[1].buf->ptr, p->t[2].buf->ptr, xn, yn, zn,
                                                                            c->partials, c->ticket, res);
        });
    } else {
        const uint32_t bn = p->t[2].nv, cn = p->t[3].nv;
        const uint64_t n_t = 1ull << (bn + cn - 1);
        DISPATCH_POLICY(f.policy, {
            k_gkrw_round<A><<<grid_for(c, n_t), kThreads, 0, g_stream>>>(f.d, p->t[0].buf->ptr, p->t[1].buf->ptr, p->t[2].buf->ptr, p->t[3].buf->ptr,
                                                                        bn, cn, c->partials, c->ticket, res);
        });
    }
    LAUNCH_CHECK();
    return SCB_OK;
}

static void g4_rebuild_evals(const HostField& h, uint32_t K, const Fe& g1, const Fe* S, Fe* ev);
static int round_evals_impl(const scb_poly* p, uint32_t n_points, uint64_t* h_out, uint64_t* d_out) {
    ARG_TRY(p && (h_out || d_out), "null argument");
    PLAIN_POLY(p);
    ARG_TRY(n_points >= 1 && n_points <= poly_n_points(p), "n_points out of range for this polynomial");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    const uint32_t N = p->f->d.n;
    if (d_out) {
        if (n_points == poly_n_points(p)) return launch_round_evals(c, p, d_out);
        RC_TRY(launch_round_evals(c, p, c->d_scratch));
        CU_TRY(cudaMemcpyAsync(d_out, c->d_scratch, (size_t)8 * N * n_points, cudaMemcpyDeviceToDevice, g_stream));
        return SCB_OK;
    }
    if ((p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G) && p->f->policy == POL_G4 && opt(OPT_g4_kernel) == 2 && g29_supported(p->f->d) &&
        p->t[0].nv >= 1) {
        // 4-limb fields: the message pass in radix-2^29 lazy-carry arithmetic (g29.cuh); sums at 0, inf, 2..K-1, 1
        const FieldImpl& f = *p->f;
        const uint32_t K = (uint32_t)p->t.size();
        const uint64_t* in[kMaxTables];
        for (uint32_t k = 0; k < K; ++k) in[k] = p->t[k].buf->ptr;
        const cudaError_t le = launch_round_evals_g29((int)K, (int)opt(OPT_bps), c->sms, g_stream, f.d, in, p->t[0].len() / 2, c->partials, c->ticket, c->h_res,
                                                      peer_arg(c), kMaxGrid);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (le != cudaSuccess) {
            set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(le), __FILE__, __LINE__);
            return SCB_ECUDA;
        }
        CU_TRY(cudaStreamSynchronize(g_stream));
        RC_TRY(peers_check(c));
        Fe S[kMaxPts + 1], ev[kMaxPts + 1];
        for (uint32_t x = 0; x <= K; ++x) f.h.load(c->h_res + (size_t)x * N, S[x]);
        if (K > 1) {
            const Fe fix = f.h.from_u64(1ull << (5 * (K - 1)));
            for (uint32_t x = 0; x <= K; ++x) S[x] = f.h.mul(S[x], fix);
        }
        g4_rebuild_evals(f.h, K, S[K], S, ev);
        for (uint32_t x = 0; x < n_points; ++x) f.h.store(ev[x], h_out + (size_t)x * N);
        return SCB_OK;
    }
    if ((p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G) && p->f->policy == POL_G4 && opt(OPT_g4_kernel) == 3 &&
        g4w_supported(p->f->d, (int)p->t.size()) && p->t[0].nv >= 1) {
        // 4-limb fields: the message pass with unreduced last products (k_round_evals_g4w); sums at 0, inf, 2..K-1, 1
        const FieldImpl& f = *p->f;
        const uint32_t K = (uint32_t)p->t.size();
        const uint64_t* in[kMaxTables];
        for (uint32_t k = 0; k < K; ++k) in[k] = p->t[k].buf->ptr;
        g_g4w_minb = (int)opt(OPT_g4_blocks4);
        const cudaError_t le = launch_round_evals_g4w((int)K, opt(OPT_g4_p0one) != 0 && g4_p0one(f.d), (int)opt(OPT_bps), c->sms, g_stream, f.d, in,
                                                      p->t[0].len() / 2, c->partials, c->ticket, c->h_res, peer_arg(c), kMaxGrid);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (le != cudaSuccess) {
            set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(le), __FILE__, __LINE__);
            return SCB_ECUDA;
        }
        CU_TRY(cudaStreamSynchronize(g_stream));
        RC_TRY(peers_check(c));
        Fe S[kMaxPts + 1], ev[kMaxPts + 1];
        for (uint32_t x = 0; x <= K; ++x) f.h.load(c->h_res + (size_t)x * N, S[x]);
        g4_rebuild_evals(f.h, K, S[K], S, ev);
        for (uint32_t x = 0; x < n_points; ++x) f.h.store(ev[x], h_out + (size_t)x * N);
        return SCB_OK;
    }
    RC_TRY(launch_round_evals(c, p, c->h_res));
    CU_TRY(cudaStreamSynchronize(g_stream));
    RC_TRY(peers_check(c));
    std::memcpy(h_out, c->h_res, (size_t)8 * N * n_points);
    return SCB_OK;
}
extern "C" int scb_poly_round_evals(const scb_poly* p, uint32_t n_points, uint64_t* out_elems) {
    return round_evals_impl(p, n_points, out_elems, nullptr);
}
extern "C" int scb_poly_round_evals_device(const scb_poly* p, uint32_t n_points, uint64_t* d_out) {
    return round_evals_impl(p, n_points, nullptr, d_out);
}

// small-prime fused fold+message with packed uint32 tables on either side (packed.cuh)
template <int K, bool IN32, bool OUT32, int QP>
static void launch_fold_sp(Ctx* c, const FieldDesc& d, TabsIn<K> in, TabsOut<K> o, ElemArg ra, uint64_t n_quads, uint64_t* res) {
    auto kern = k_fold_round_sp<K, IN32, OUT32, QP>;
    const int bps32 = (int)opt(OPT_bps32);
    const int pref = IN32 ? bps32 : 5;  // measured sweet spots (profiles/r01_kernel_sweep.md)
    kern<<<occ_grid(c, kern, n_quads / QP, 0, pref), kThreads, 0, g_stream>>>(d, in, o, ra, n_quads / QP, c->partials, c->ticket, res, peer_arg(c));
}

// Polynomials with packed tables only exist as descendants of a handle marked with scb_poly_allow_packed; every
// entry point other than the round/tail calls first converts them back to ark's 8-byte layout.
static int unpack_table(Ctx* c, const FieldImpl& f, const Table& in, Table* out) {
    if (!in.p32) {
        *out = in;
        return SCB_OK;
    }
    Table o;
    o.nv = in.nv;
    RC_TRY(alloc_buf((size_t)8 * f.d.n << in.nv, &o.buf));
    k_unpack32<<<grid_for(c, in.len()), kThreads, 0, g_stream>>>((const uint32_t*)in.buf->ptr, o.buf->ptr, in.len());
    LAUNCH_CHECK();
    *out = o;
    return SCB_OK;
}
static int plain_poly(const scb_poly* p, std::unique_ptr<scb_poly>* holder, const scb_poly** out) {
    if (!p->any_packed()) {
        *out = p;
        return SCB_OK;
    }
    Ctx* c;
    RC_TRY(get_ctx(&c));
    auto q = std::make_unique<scb_poly>(*p);
    for (size_t k = 0; k < p->t.size(); ++k) RC_TRY(unpack_table(c, *p->f, p->t[k], &q->t[k]));
    *holder = std::move(q);
    *out = holder->get();
    return SCB_OK;
}
extern "C" int scb_poly_allow_packed(scb_poly* p, int enable) {
    ARG_TRY(p, "null argument");
    p->allow_packed = enable != 0;
    return SCB_OK;
}

// The 4-limb kernel (g4.cuh) accumulates S_0 = sum prod lo, S_inf = sum prod (hi - lo) (the leading coefficient) and
// S_x = g(x) for x = 2..K-1; with the claim g(0) + g(1) the message values g(0..K) follow by exact field arithmetic:
// g(1) = claim - g(0) and, from the K-th finite difference  sum_i (-1)^(K-i) C(K,i) g(i) = K! * lead,
// g(K) = K! * lead - sum_{i<K} (-1)^(K-i) C(K,i) g(i).  Same field elements as summing every point.
static void g4_rebuild_evals(const HostField& h, uint32_t K, const Fe& g1, const Fe* S, Fe* ev) {
    ev[0] = S[0];
    ev[1] = g1;  // claim - g(0), or summed directly when no claim exists yet (round 0)
    if (K == 1) return;
    for (uint32_t x = 2; x < K; ++x) ev[x] = S[x];
    Fe fact = h.one(), acc = h.zero();
    for (uint32_t i = 2; i <= K; ++i) fact = h.mul(fact, h.from_u64(i));
    uint64_t binom = 1;  // C(K, i)
    for (uint32_t i = 0; i < K; ++i) {
        const Fe term = h.mul(h.from_u64(binom), ev[i]);
        acc = ((K - i) & 1) ? h.sub(acc, term) : h.add(acc, term);
        binom = binom * (K - i) / (i + 1);
    }
    ev[K] = h.sub(h.mul(fact, S[1]), acc);
}

// fused fold + message (product kinds: one kernel; mixed-arity kinds: fold kernels then message kernel).
// claim (optional, host results only) = g(0) + g(1) of the message about to be computed, i.e. g_{j-1}(r_{j-1}): lets the
// 4-limb kernel skip one point (g4.cuh).
static int fix_and_round_impl(const scb_poly* p, const uint64_t* r, uint32_t n_points, scb_poly** out, uint64_t* h_out, uint64_t* d_out,
                              const uint64_t* claim = nullptr) {
    ARG_TRY(p && r && out && (h_out || d_out), "null argument");
    ARG_TRY(n_points >= 1 && n_points <= poly_n_points(p), "n_points out of range for this polynomial");
    ARG_TRY(poly_num_vars(p) >= 2, "need at least two variables to fix one and send a message");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    const FieldImpl& f = *p->f;
    const uint32_t N = f.d.n;
    ARG_TRY(elem_canonical(f, r), "challenge is not a canonical field element");
    ARG_TRY(!claim || elem_canonical(f, claim), "claim is not a canonical field element");
    uint64_t* res = d_out ? (n_points == poly_n_points(p) ? d_out : c->d_scratch) : c->h_res;
    std::unique_ptr<scb_poly> q;
    if ((p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G) && f.policy == POL_G4 && claim && h_out && opt(OPT_g4_kernel) != 0 &&
        f.h.bits <= 255 && p->t[0].nv >= 2) {
        // 4-limb fields: leaner kernel, one point fewer, message rebuilt from the claim on the host
        const uint32_t K = (uint32_t)p->t.size();
        q = std::make_unique<scb_poly>(*p);
        const uint64_t* in[kMaxTables];
        uint64_t* o[kMaxTables];
        for (uint32_t k = 0; k < K; ++k) {
            q->t[k].nv = p->t[k].nv - 1;
            q->t[k].buf.reset();
            RC_TRY(alloc_buf((size_t)8 * N << q->t[k].nv, &q->t[k].buf));
            in[k] = p->t[k].buf->ptr;
            o[k] = q->t[k].buf->ptr;
        }
        // option g4_kernel: 3 (default) = carry chains with unreduced last products (k_fold_round_g4w), 1 = carry chains
        // (k_fold_round_g4), 2 = radix-2^29 lazy-carry arithmetic (g29.cuh) where the modulus allows
        const bool use29 = opt(OPT_g4_kernel) == 2 && g29_supported(f.d);
        const bool usew = opt(OPT_g4_kernel) == 3 && g4w_supported(f.d, (int)K);
        cudaError_t le;
        if (usew) {
            g_g4w_minb = (int)opt(OPT_g4_blocks4);
            // fold table of the challenge: t[i] = r 2^(32 i + 64) mod p as plain integers (64 + 7 x 32 modular doublings)
            g4::FoldTab ft;
            {
                Fe rr;
                f.h.load(r, rr);
                Fe x = f.h.from_mont(rr);
                for (int i = 0; i < 8; ++i) {
                    for (int s = 0; s < (i == 0 ? 64 : 32); ++s) x = f.h.double_raw(x);
                    for (int q = 0; q < 4; ++q) {
                        ft.t[i][2 * q] = (uint32_t)x.l[q];
                        ft.t[i][2 * q + 1] = (uint32_t)(x.l[q] >> 32);
                    }
                }
            }
            le = launch_fold_round_g4w((int)K, opt(OPT_g4_p0one) != 0 && g4_p0one(f.d), (int)opt(OPT_bps), c->sms, g_stream, f.d, in, o, ft,
                                       p->t[0].len() / 4, c->partials, c->ticket, c->h_res, peer_arg(c), kMaxGrid);
        } else if (use29) {
            Fe rr, r5;
            f.h.load(r, rr);
            r5 = f.h.mul(rr, f.h.from_u64(32));  // r * 2^5: the fold's product divides by 2^261 instead of 2^256
            uint64_t r5w[kMaxLimbs];
            f.h.store(r5, r5w);
            le = launch_fold_round_g29((int)K, (int)opt(OPT_g4_blocks), (int)opt(OPT_bps), c->sms, g_stream, f.d, in, o, elem_arg(f, r5w), p->t[0].len() / 4, c->partials, c->ticket,
                                       c->h_res, peer_arg(c), kMaxGrid);
        } else {
            le = launch_fold_round_g4((int)K, (int)opt(OPT_g4_blocks), (int)opt(OPT_bps), c->sms, g_stream, f.d, in, o, elem_arg(f, r), p->t[0].len() / 4, c->partials, c->ticket,
                                      c->h_res, peer_arg(c), kMaxGrid);
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (le != cudaSuccess) {
            set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(le), __FILE__, __LINE__);
            return SCB_ECUDA;
        }
        CU_TRY(cudaStreamSynchronize(g_stream));
        RC_TRY(peers_check(c));
        Fe S[kMaxPts], ev[kMaxPts], cl;
        for (int x = 0; x < g4_n_sums((int)K); ++x) f.h.load(c->h_res + (size_t)x * N, S[x]);
        if (use29 && K > 1) {  // every product of K factors came back short of 2^(5 (K-1))
            const Fe fix = f.h.from_u64(1ull << (5 * (K - 1)));
            for (int x = 0; x < g4_n_sums((int)K); ++x) S[x] = f.h.mul(S[x], fix);
        }
        f.h.load(claim, cl);
        g4_rebuild_evals(f.h, K, f.h.sub(cl, S[0]), S, ev);
        for (uint32_t x = 0; x < n_points; ++x) f.h.store(ev[x], h_out + (size_t)x * N);
        *out = q.release();
        return SCB_OK;
    }
    if (p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G) {
        ARG_TRY(f.policy != POL_SP || p->t[0].nv <= 32, "table too large for the small-prime path");
        q = std::make_unique<scb_poly>(*p);
        const uint64_t n_quads = p->t[0].len() / 4;
        // small-prime policy: the folded tables may be stored as packed uint32 when the handle allows it
        const bool in32 = p->t[0].p32;
        const bool out32 = f.policy == POL_SP && p->allow_packed;
        for (size_t k = 0; k < p->t.size(); ++k) {
            ARG_TRY(p->t[k].p32 == in32, "tables of one polynomial must share a layout");
            q->t[k].nv = p->t[k].nv - 1;
            q->t[k].buf.reset();
            q->t[k].p32 = out32;
            RC_TRY(alloc_buf((size_t)(out32 ? 4 : 8 * N) << q->t[k].nv, &q->t[k].buf));
        }
        const ElemArg ra = elem_arg(f, r);
        const int qp32 = opt(OPT_qp32) > 0 ? (int)opt(OPT_qp32) : 4;
        if (f.policy == POL_SP && (in32 || out32)) {
            DISPATCH_K(p->t.size(), {
                TabsIn<K> in;
                TabsOut<K> o;
                for (int k = 0; k < K; ++k) {
                    in.p[k] = p->t[k].buf->ptr;
                    o.p[k] = q->t[k].buf->ptr;
                }
                if (!in32) launch_fold_sp<K, false, true, 1>(c, f.d, in, o, ra, n_quads, res);  // out32 holds here
                else if (out32) {
                    if (n_quads >= 4 && qp32 >= 4) launch_fold_sp<K, true, true, 4>(c, f.d, in, o, ra, n_quads, res);
                    else if (n_quads >= 2 && qp32 >= 2) launch_fold_sp<K, true, true, 2>(c, f.d, in, o, ra, n_quads, res);
                    else launch_fold_sp<K, true, true, 1>(c, f.d, in, o, ra, n_quads, res);
                } else {
                    if (n_quads >= 2) launch_fold_sp<K, true, false, 2>(c, f.d, in, o, ra, n_quads, res);
                    else launch_fold_sp<K, true, false, 1>(c, f.d, in, o, ra, n_quads, res);
                }
            });
        } else
        DISPATCH_POLICY(f.policy, DISPATCH_K(p->t.size(), {
            TabsIn<K> in;
            TabsOut<K> o;
            for (int k = 0; k < K; ++k) {
                in.p[k] = p->t[k].buf->ptr;
                o.p[k] = q->t[k].buf->ptr;
            }
            auto kern = k_fold_round<A, K, 1>;
            kern<<<occ_grid(c, kern, n_quads, 0, A::kLight ? 5 : 0), kThreads, 0, g_stream>>>(f.d, in, o, ra, n_quads, c->partials, c->ticket, res, peer_arg(c));
        }));
        LAUNCH_CHECK();
    } else {
        scb_poly* qq = nullptr;
        RC_TRY(scb_poly_fix_variables(p, r, 1, &qq));
        q.reset(qq);
        RC_TRY(launch_round_evals(c, q.get(), res));
    }
    if (d_out) {
        if (res != d_out) CU_TRY(cudaMemcpyAsync(d_out, res, (size_t)8 * N * n_points, cudaMemcpyDeviceToDevice, g_stream));
    } else {
        CU_TRY(cudaStreamSynchronize(g_stream));
        RC_TRY(peers_check(c));
        std::memcpy(h_out, c->h_res, (size_t)8 * N * n_points);
    }
    *out = q.release();
    return SCB_OK;
}
extern "C" int scb_poly_fix_and_round_evals(const scb_poly* p, const uint64_t* r, uint32_t n_points, scb_poly** out, uint64_t* out_elems) {
    return fix_and_round_impl(p, r, n_points, out, out_elems, nullptr);
}
extern "C" int scb_poly_fix_and_round_evals_claim(const scb_poly* p, const uint64_t* r, const uint64_t* claim, uint32_t n_points, scb_poly** out,
                                                  uint64_t* out_elems) {
    ARG_TRY(claim, "null argument");
    return fix_and_round_impl(p, r, n_points, out, out_elems, nullptr, claim);
}
extern "C" int scb_poly_fix_and_round_evals_device(const scb_poly* p, const uint64_t* r, uint32_t n_points, scb_poly** out, uint64_t* d_out) {
    return fix_and_round_impl(p, r, n_points, out, nullptr, d_out);
}

// c_1 (Prover::new)
extern "C" int scb_poly_sum(const scb_poly* p, uint64_t* out_elem) {
    ARG_TRY(p && out_elem, "null argument");
    PLAIN_POLY(p);
    Ctx* c;
    RC_TRY(get_ctx(&c));
    const FieldImpl& f = *p->f;
    if (p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G) {
        const uint64_t n = p->t[0].len();
        ARG_TRY(f.policy != POL_SP || p->t[0].nv <= 32, "table too large for the small-prime path");
        DISPATCH_POLICY(f.policy, DISPATCH_K(p->t.size(), {
            TabsIn<K> in;
            for (int k = 0; k < K; ++k) in.p[k] = p->t[k].buf->ptr;
            if (A::N == 1 && n >= 4)
                k_product_sum<A, K, (A::N == 1 ? 4 : 1)><<<grid_for(c, n / 4), kThreads, 0, g_stream>>>(f.d, in, n / 4, c->partials, c->ticket, c->h_res);
            else
                k_product_sum<A, K, 1><<<grid_for(c, n), kThreads, 0, g_stream>>>(f.d, in, n, c->partials, c->ticket, c->h_res);
        }));
    } else if (p->kind == SCB_POLY_TRIANGLE_G) {
        const uint32_t xn = tri_xn(p), yn = tri_yn(p), zn = tri_zn(p);
        if (p->has_aux && xn > 0) {  // c_1 = sum_{x,z} M[z][x] f3[z][x]
            DISPATCH_POLICY(f.policy, {
                TabsIn<2> in;
                in.p[0] = p->aux.buf->ptr;
                in.p[1] = p->t[2].buf->ptr;
                const uint64_t n = 1ull << (xn + zn);
                k_product_sum<A, 2, 1><<<grid_for(c, n), kThreads, 0, g_stream>>>(f.d, in, n, c->partials, c->ticket, c->h_res);
            });
        } else
        DISPATCH_POLICY(f.policy, {
            k_triangle_sum<A><<<grid_for(c, 1ull << (xn + zn)), kThreads, 0, g_stream>>>(f.d, p->t[0].buf->ptr, p->t[1].buf->ptr, p->t[2].buf->ptr, xn, yn,
                                                                                        zn, c->partials, c->ticket, c->h_res);
        });
    } else {
        const uint32_t bn = p->t[2].nv, cn = p->t[3].nv;
        DISPATCH_POLICY(f.policy, {
            k_gkrw_sum<A><<<grid_for(c, 1ull << (bn + cn)), kThreads, 0, g_stream>>>(f.d, p->t[0].buf->ptr, p->t[1].buf->ptr, p->t[2].buf->ptr,
                                                                                    p->t[3].buf->ptr, bn, cn, c->partials, c->ticket, c->h_res);
        });
    }
    LAUNCH_CHECK();
    CU_TRY(cudaStreamSynchronize(g_stream));
    std::memcpy(out_elem, c->h_res, 8 * f.d.n);
    return SCB_OK;
}

// SumCheckPolynomial::to_evaluations
extern "C" int scb_poly_to_evaluations(const scb_poly* p, uint64_t* out, size_t cap_elems) {
    ARG_TRY(p && out, "null argument");
    PLAIN_POLY(p);
    Ctx* c;
    RC_TRY(get_ctx(&c));
    const FieldImpl& f = *p->f;
    const uint32_t nv = poly_num_vars(p);
    ARG_TRY(nv <= 34 && cap_elems >= ((size_t)1 << nv), "output buffer too small");
    const uint64_t n = 1ull << nv;
    BufRef tmp;
    RC_TRY(alloc_buf((size_t)8 * f.d.n << nv, &tmp));
    if (p->kind == SCB_POLY_PRODUCT || p->kind == SCB_POLY_MATMUL_G) {
        DISPATCH_POLICY(f.policy, DISPATCH_K(p->t.size(), {
            TabsIn<K> in;
            for (int k = 0; k < K; ++k) in.p[k] = p->t[k].buf->ptr;
            k_product_table<A, K><<<grid_for(c, n), kThreads, 0, g_stream>>>(f.d, in, tmp->ptr, n);
        }));
    } else if (p->kind == SCB_POLY_TRIANGLE_G) {
        DISPATCH_POLICY(f.policy, {
            k_triangle_table<A><<<grid_for(c, n), kThreads, 0, g_stream>>>(f.d, p->t[0].buf->ptr, p->t[1].buf->ptr, p->t[2].buf->ptr, tri_xn(p), tri_yn(p),
                                                                          tri_zn(p), tmp->ptr);
        });
    } else {
        DISPATCH_POLICY(f.policy, {
            k_gkrw_table<A><<<grid_for(c, n), kThreads, 0, g_stream>>>(f.d, p->t[0].buf->ptr, p->t[1].buf->ptr, p->t[2].buf->ptr, p->t[3].buf->ptr,
                                                                      p->t[2].nv, p->t[3].nv, tmp->ptr);
        });
    }
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(out, tmp->ptr, (size_t)8 * f.d.n << nv, cudaMemcpyDeviceToHost, g_stream));
    CU_TRY(cudaStreamSynchronize(g_stream));
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------ resident kernels, pair passes
#include "resident_engine.inc"

// ------------------------------------------------------------------------------------------ peers API
__global__ void __launch_bounds__(kThreads) k_peer_publish(PeerArg pa, const uint4* __restrict__ src, uint64_t n16, uint64_t dst_byte_off) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = src[i];
        for (uint32_t g = 0; g < pa.world; ++g) reinterpret_cast<uint4*>(reinterpret_cast<char*>(pa.win[g]) + dst_byte_off)[i] = v;
    }
}
__global__ void k_peer_gather_sync(PeerArg pa) {  // one thread: publish "my slabs are written", wait for everyone's
    __threadfence_system();
    for (uint32_t g = 0; g < pa.world; ++g) st_sys(pa.win[g] + kWinGatherFlags + pa.rank, pa.seq);
    const uint64_t t0 = globaltimer_ns();
    for (uint32_t g = 0; g < pa.world; ++g) {
        while (ld_sys(pa.win[pa.rank] + kWinGatherFlags + g) < pa.seq) {
            if (globaltimer_ns() - t0 > pa.timeout_ns) {
                st_sys(pa.status, 1);
                return;
            }
        }
    }
    __threadfence_system();
}

extern "C" int scb_peers_create(uint32_t rank, uint32_t world, size_t gather_bytes, scb_peers** out, uint8_t* handle_out) {
    ARG_TRY(out && handle_out, "null argument");
    ARG_TRY(world >= 1 && world <= (uint32_t)kMaxRanks && (world & (world - 1)) == 0 && rank < world, "world must be a power of two <= 8");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    auto p = std::make_unique<scb_peers>();
    p->rank = rank;
    p->world = world;
    p->gather_bytes = (gather_bytes + 255) & ~(size_t)255;
    const size_t bytes = (size_t)kWinGatherWords * 8 + 2 * p->gather_bytes;
    void* w = nullptr;
    CU_TRY(cudaMalloc(&w, bytes));  // plain cudaMalloc: pool (cudaMallocAsync) memory cannot be exported through IPC
    CU_TRY(cudaMemset(w, 0, bytes));
    CU_TRY(cudaDeviceSynchronize());
    p->win[rank] = (uint64_t*)w;
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, w));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handle_out, &h, 64);
    *out = p.release();
    return SCB_OK;
}
extern "C" int scb_peers_connect(scb_peers* p, const uint8_t* all_handles) {
    ARG_TRY(p && all_handles, "null argument");
    for (uint32_t g = 0; g < p->world; ++g) {
        if (g == p->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, all_handles + (size_t)g * 64, 64);
        void* w = nullptr;
        CU_TRY(cudaIpcOpenMemHandle(&w, h, cudaIpcMemLazyEnablePeerAccess));
        p->win[g] = (uint64_t*)w;
    }
    p->connected = true;
    return SCB_OK;
}
extern "C" void scb_peers_free(scb_peers* p) {
    if (!p) return;
    if (g_cur_peers == p) g_cur_peers = nullptr;
    for (uint32_t g = 0; g < p->world; ++g) {
        if (!p->win[g]) continue;
        if (g == p->rank) cudaFree(p->win[g]);
        else cudaIpcCloseMemHandle(p->win[g]);
    }
    delete p;
}
extern "C" int scb_peers_set_current(scb_peers* p) {
    ARG_TRY(!p || p->connected || p->world == 1, "scb_peers_connect has not been called");
    g_cur_peers = p;
    if (p) {
        Ctx* c;
        RC_TRY(get_ctx(&c));
        c->h_res[kStatusWord] = 0;
    }
    return SCB_OK;
}
static int peers_check(Ctx* c) {
    if (g_cur_peers && c->h_res[kStatusWord] != 0) {
        set_error("peer exchange timed out (a rank is missing or stuck)");
        return SCB_ENCCL;
    }
    return SCB_OK;
}
extern "C" int scb_peers_gather_capacity(const scb_peers* p, size_t* out_bytes) {
    ARG_TRY(p && out_bytes, "null argument");
    *out_bytes = p->gather_bytes;
    return SCB_OK;
}
// Consolidation: every rank writes its slabs into every peer's gather area (NVLink P2P stores); afterwards each
// rank holds the rank-order concatenation of all slabs and continues replicated, with no further communication.
extern "C" int scb_peers_gather_poly(scb_peers* p, const scb_poly* slab, scb_poly** out) {
    ARG_TRY(p && slab && out, "null argument");
    ARG_TRY(slab->kind == SCB_POLY_PRODUCT || slab->kind == SCB_POLY_MATMUL_G, "only product polynomials shard");
    Ctx* c;
    RC_TRY(get_ctx(&c));
    const FieldImpl& f = *slab->f;
    const uint32_t lg = 31 - __builtin_clz(p->world);
    const size_t K = slab->t.size();
    const bool p32 = slab->t[0].p32;
    const size_t table_bytes = (size_t)(p32 ? 4 : 8 * f.d.n) << slab->t[0].nv;
    const size_t region = (p->world * table_bytes + 255) & ~(size_t)255;
    ARG_TRY(K * region <= p->gather_bytes, "peer window too small for this consolidation (scb_peers_create gather_bytes)");
    PeerArg pa;
    std::memset(&pa, 0, sizeof pa);
    for (uint32_t g = 0; g < p->world; ++g) pa.win[g] = p->win[g];
    pa.rank = p->rank;
    pa.world = p->world;
    pa.seq = ++p->gseq;
    pa.status = c->h_res + kStatusWord;
    pa.timeout_ns = 10ull * 1000 * 1000 * 1000;
    const size_t area = (size_t)kWinGatherWords * 8 + (pa.seq & 1) * p->gather_bytes;
    auto q = std::make_unique<scb_poly>(*slab);
    for (size_t k = 0; k < K; ++k) {
        ARG_TRY(slab->t[k].p32 == p32 && slab->t[k].nv == slab->t[0].nv, "tables of one polynomial must share a layout");
        const size_t off = area + k * region + p->rank * table_bytes;
        if (table_bytes % 16 == 0) {
            k_peer_publish<<<grid_for(c, table_bytes / 16), kThreads, 0, g_stream>>>(pa, (const uint4*)slab->t[k].buf->ptr, table_bytes / 16, off);
            LAUNCH_CHECK();
        } else {  // tiny slabs: plain copies
            for (uint32_t g = 0; g < p->world; ++g)
                CU_TRY(cudaMemcpyAsync((char*)p->win[g] + off, slab->t[k].buf->ptr, table_bytes, cudaMemcpyDefault, g_stream));
        }
        q->t[k].nv = slab->t[k].nv + lg;
        q->t[k].buf = std::make_shared<DevBuf>();
        q->t[k].buf->ptr = (uint64_t*)((char*)p->win[p->rank] + area + k * region);
        q->t[k].buf->bytes = p->world * table_bytes;
        q->t[k].buf->owned = false;
    }
    k_peer_gather_sync<<<1, 1, 0, g_stream>>>(pa);
    LAUNCH_CHECK();
    CU_TRY(cudaStreamSynchronize(g_stream));
    if (c->h_res[kStatusWord] != 0) {
        set_error("peer gather timed out (a rank is missing or stuck)");
        return SCB_ENCCL;
    }
    *out = q.release();
    return SCB_OK;
}

// accessor used by protocol.cpp
extern "C" int scb_poly_field_impl(const scb_poly* p, const FieldImpl** out) {
    ARG_TRY(p && out, "null argument");
    *out = p->f.get();
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------ packed upload of host tables
#include "upload_engine.inc"

// ------------------------------------------------------------------------------------------ GKR (sparse wiring)
#include "gkr_engine.inc"
