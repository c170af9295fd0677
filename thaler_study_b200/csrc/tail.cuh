// tail.cuh -- persistent single-CTA kernel for the latency-bound tail of a product sum-check.
//
// Once the live tables are small (<= 2^kTailMaxVars entries) a round is a few microseconds of work and
// the per-round cost is dominated by launch + completion latency (hard part 5).
// Fiat-Shamir makes round j+1 depend on the hash of round j's message, so rounds cannot be batched;
// instead ONE kernel stays resident for all remaining rounds and talks to the host through a mailbox in
// mapped pinned memory: it posts the (d+1) round sums, the host derives the challenge with the unchanged
// transcript code (interpolation, serialization, SHA-256 -- the part that decides transcript bytes) and
// posts it back.  No kernel launches, no stream synchronisation, no fences: self-validating tagged words.
//
// Same arithmetic as k_fold_round (kernels.cuh); buffers written inside this kernel are read back with
// ld.global.cg (never through the non-coherent path).
#pragma once
#include <cstdint>

#include "kernels.cuh"

namespace scb {

constexpr uint32_t kTailMaxRounds = 32;

// Mapped pinned host memory.  Every 8-byte word validates itself: it carries a 32-bit payload (one half of a limb,
// or a whole limb for the small-prime policy) under a 32-bit tag, so neither side needs a fence or a second
// "ready" word -- posted PCIe writes may land in any order and the reader just waits until every word it needs
// shows the expected tag.  One PCIe posted write per word on the way out, one read round trip per poll on the way in.
constexpr uint32_t kMbAbortTag = 0xFFFFFFFFu;
struct TailMailbox {
    volatile uint64_t evals[2 * kMaxPts * kMaxLimbs];  // device -> host, tag = round + 1
    volatile uint64_t challenge[2 * kMaxLimbs];        // host -> device, tag = challenge number (>= 1) or kMbAbortTag
    volatile uint64_t dev_status;                      // device -> host: 0 running, 1 done, 2 timed out
    // %globaltimer stamps written by the grid-wide kernel: [2t] round t's sums posted, [2t+1] challenge t+1 received
    uint64_t stamp[2 * (kTailMaxRounds + 1)];
};
template <class A>
struct MailboxHalves {  // tagged words per limb
    static constexpr int value = A::kLight ? 1 : 2;
};

template <class A, int NP>
__device__ __forceinline__ void mailbox_post(TailMailbox* mb, uint32_t tag, const uint64_t (&w)[NP][A::N]) {
    constexpr int H = MailboxHalves<A>::value, N = A::N;
    const uint64_t hi = (uint64_t)tag << 32;
#pragma unroll
    for (int x = 0; x < NP; ++x)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            st_sys(&mb->evals[(x * N + i) * H], hi | (uint32_t)w[x][i]);
            if constexpr (H == 2) st_sys(&mb->evals[(x * N + i) * H + 1], hi | (w[x][i] >> 32));
        }
}
// Waits until every tagged word of a challenge shows `tag`; 0 = ok (r filled), 1 = abort tag seen or timeout.
// SYS: the words live in mapped host memory (one PCIe read round trip per poll), else in device memory.
template <class A, bool SYS>
__device__ __forceinline__ int tagged_wait(const volatile uint64_t* words, uint32_t tag, uint64_t timeout_ns, uint64_t* r) {
    constexpr int H = MailboxHalves<A>::value, N = A::N, W = N * H;
    const uint64_t t0 = globaltimer_ns();
    for (;;) {
        uint64_t c[W];
        if constexpr (W == 1) {
            if constexpr (SYS) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(c[0]) : "l"(words) : "memory");
            else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(c[0]) : "l"(words) : "memory");
        } else {
#pragma unroll
            for (int i = 0; i < W; i += 2) {
                if constexpr (SYS)
                    asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(c[i]), "=l"(c[i + 1]) : "l"(words + i) : "memory");
                else
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(c[i]), "=l"(c[i + 1]) : "l"(words + i) : "memory");
            }
        }
        bool ok = true, ab = false;
#pragma unroll
        for (int i = 0; i < W; ++i) {
            const uint32_t tg = (uint32_t)(c[i] >> 32);
            ok = ok && tg == tag;
            ab = ab || tg == kMbAbortTag;
        }
        if (ok) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if constexpr (H == 1) r[i] = (uint32_t)c[i];
                else r[i] = (uint64_t)(uint32_t)c[2 * i] | (c[2 * i + 1] << 32);
            }
            return 0;
        }
        if (ab || globaltimer_ns() - t0 > timeout_ns) return 1;
    }
}
template <class A>
__device__ __forceinline__ int mailbox_wait_challenge(TailMailbox* mb, uint32_t tag, uint64_t timeout_ns, uint64_t* r) {
    return tagged_wait<A, true>(mb->challenge, tag, timeout_ns, r);
}

template <int W>
__device__ __forceinline__ void ld_words_cg(const uint64_t* ptr, uint64_t* w) {
    if constexpr (W % 2 == 0) {
#pragma unroll
        for (int i = 0; i < W; i += 2)
            asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(w[i]), "=l"(w[i + 1]) : "l"(ptr + i) : "memory");
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(w[i]) : "l"(ptr + i) : "memory");
    }
}
// one quad (4 adjacent entries) of a table that is either in ark's 8-byte format or packed uint32 (PolSP only)
template <class A>
__device__ __forceinline__ void tail_ld_quad(const A& ar, const uint64_t* base, uint64_t i, bool w32, typename A::El (&t)[4]) {
    constexpr int N = A::N;
    if constexpr (A::kLight) {
        if (w32) {
            uint64_t w[2];
            ld_words_cg<2>(base + i * 2, w);
            t[0] = (uint32_t)w[0];
            t[1] = (uint32_t)(w[0] >> 32);
            t[2] = (uint32_t)w[1];
            t[3] = (uint32_t)(w[1] >> 32);
            return;
        }
    }
    uint64_t w[4 * N];
    ld_words_cg<4 * N>(base + i * 4 * N, w);
#pragma unroll
    for (int q = 0; q < 4; ++q) t[q] = ar.from_words(w + q * N);
}
template <class A>
__device__ __forceinline__ void tail_st_pair(const A& ar, uint64_t* base, uint64_t i, bool w32, const typename A::El& u0,
                                             const typename A::El& u1) {
    constexpr int N = A::N;
    if constexpr (A::kLight) {
        if (w32) {
            __stcg(base + i, (uint64_t)u0 | ((uint64_t)u1 << 32));
            return;
        }
    }
    uint64_t o[2 * N];
    ar.to_words(u0, o);
    ar.to_words(u1, o + N);
#pragma unroll
    for (int q = 0; q < 2 * N; ++q) __stcg(base + i * 2 * N + q, o[q]);
}

template <class A>
constexpr int tail_threads() {
    return A::kLight ? 1024 : (A::N == 1 ? 512 : 256);
}

// Runs `n_rounds` fused fold+message rounds on tables of 2^m entries (m >= 2, n_rounds <= m - 1).
// Round 0 folds by `r0`; round t > 0 folds by challenge number t posted by the host.
template <class A, int K>
__global__ void __launch_bounds__(tail_threads<A>(), 1)
    k_tail_rounds(FieldDesc f, TabsIn<K> in0, TabsOut<K> buf_a, TabsOut<K> buf_b, ElemArg r0, uint32_t m, uint32_t n_rounds,
                  TailMailbox* mb, uint64_t timeout_ns, int in0_w32) {
    // the internal ping-pong buffers are packed uint32 for the small-prime policy; the input may be either
    bool src_w32 = A::kLight && in0_w32 != 0;
    const bool buf_w32 = A::kLight;
    constexpr int NP = K + 1, N = A::N;
    const A ar(f);
    __shared__ uint64_t sm[32 * NP * A::AW];
    __shared__ uint64_t r_sm[kMaxLimbs];
    __shared__ int abort_sm;
    const uint64_t* src[K];
    uint64_t* dst[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        src[k] = in0.p[k];
        dst[k] = buf_a.p[k];
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) r_sm[i] = r0.w[i];
        abort_sm = 0;
    }
    __syncthreads();
    for (uint32_t t = 0; t < n_rounds; ++t) {
        if (t > 0) {
            if (threadIdx.x == 0) abort_sm = mailbox_wait_challenge<A>(mb, t, timeout_ns, r_sm);  // the host's challenge number t
            __syncthreads();
            if (abort_sm) {
                if (threadIdx.x == 0) st_sys(&mb->dev_status, 2);
                return;
            }
        }
        const typename A::FoldC r = ar.fold_const(ar.from_words(r_sm));
        typename A::Acc acc[NP];
#pragma unroll
        for (int x = 0; x < NP; ++x) ar.acc_zero(acc[x]);
        const uint64_t n_quads = 1ull << (m - 2);
        for (uint64_t i = threadIdx.x; i < n_quads; i += blockDim.x) {
            typename A::Lz prod[NP];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                typename A::El t4[4];
                tail_ld_quad<A>(ar, src[k], i, src_w32, t4);
                typename A::El u0 = ar.fold_c(t4[0], t4[1], r);
                typename A::El u1 = ar.fold_c(t4[2], t4[3], r);
                tail_st_pair<A>(ar, dst[k], i, buf_w32, u0, u1);
                pair_into_prod<A, NP>(ar, k == 0, u0, u1, prod);
            }
#pragma unroll
            for (int x = 0; x < NP; ++x) ar.acc_add(acc[x], prod[x]);
        }
        block_reduce<A, NP>(ar, acc, sm);
        if (threadIdx.x == 0) {
            uint64_t w[NP][N];
#pragma unroll
            for (int x = 0; x < NP; ++x) ar.to_words(ar.msg_final(acc[x], K), w[x]);
            mailbox_post<A, NP>(mb, t + 1, w);
        }
        __syncthreads();  // folded tables written by all threads are visible to the next round
        m -= 1;
        src_w32 = buf_w32;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint64_t* s = dst[k];
            dst[k] = (t & 1) ? buf_a.p[k] : buf_b.p[k];
            src[k] = s;
        }
    }
    if (threadIdx.x == 0) st_sys(&mb->dev_status, 1);
}

}  // namespace scb
