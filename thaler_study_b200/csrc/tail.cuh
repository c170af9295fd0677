// tail.cuh -- persistent single-CTA kernel for the latency-bound tail of a product sum-check.
//
// Once the live tables are small (<= 2^kTailMaxVars entries) a round is a few microseconds of work and
// the per-round cost is dominated by launch + completion latency (hard part 5).
// Fiat-Shamir makes round j+1 depend on the hash of round j's message, so rounds cannot be batched;
// instead ONE kernel stays resident for all remaining rounds and talks to the host through a mailbox in
// mapped pinned memory: it posts the (d+1) round sums, the host derives the challenge with the unchanged
// transcript code (interpolation, serialization, SHA-256 -- the part that decides transcript bytes) and
// posts it back.  No kernel launches, no stream synchronisation, two PCIe posted writes per round.
//
// Same arithmetic as k_fold_round (kernels.cuh); buffers written inside this kernel are read back with
// ld.global.cg (never through the non-coherent path).
#pragma once
#include <cstdint>

#include "kernels.cuh"

namespace scb {

constexpr uint32_t kTailMaxRounds = 32;

// Mapped pinned host memory.  seq_dev / seq_host count completed posts (monotone within one kernel).
struct TailMailbox {
    volatile uint64_t seq_dev;                    // device -> host: round sums of round `seq_dev` are valid
    uint64_t evals[kMaxPts * kMaxLimbs];
    volatile uint64_t seq_host;                   // host -> device: challenge number `seq_host` is valid
    uint64_t challenge[kMaxLimbs];
    volatile uint64_t abort_flag;                 // host -> device: give up (error on the host side)
    volatile uint64_t dev_status;                 // device -> host: 0 running, 1 done, 2 timed out
};

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
            if (threadIdx.x == 0) {  // wait for the host's challenge number t
                const uint64_t t0 = globaltimer_ns();
                int bad = 0;
                while (ld_sys(&mb->seq_host) < t) {
                    if (ld_sys(&mb->abort_flag) != 0 || globaltimer_ns() - t0 > timeout_ns) {
                        bad = 1;
                        break;
                    }
                }
                __threadfence_system();
                if (!bad) {
#pragma unroll
                    for (int i = 0; i < N; ++i) r_sm[i] = ld_sys(&mb->challenge[i]);
                }
                abort_sm = bad;
            }
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
#pragma unroll
            for (int x = 0; x < NP; ++x) {
                uint64_t w[N];
                ar.to_words(ar.msg_final(acc[x], K), w);
#pragma unroll
                for (int i = 0; i < N; ++i) st_sys(&mb->evals[x * N + i], w[i]);
            }
            __threadfence_system();
            st_sys(&mb->seq_dev, (uint64_t)t + 1);
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
