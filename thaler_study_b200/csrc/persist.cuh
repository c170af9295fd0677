// persist.cuh -- grid-wide resident kernel for ALL rounds after the first of a product sum-check.
//
// The per-round kernels (k_fold_round*, packed.cuh) pay a launch, a grid ramp-up and a stream synchronisation per
// round; from the third round on the tables are small enough that this fixed cost is comparable to the streaming
// time.  This kernel generalises the single-CTA tail (tail.cuh) to a full co-resident grid: one cooperative launch
// runs every remaining round.  Between rounds the CTAs meet at a ticket/flag barrier in device memory; the CTA that
// takes the last ticket of a round reduces the per-CTA partial sums, posts the round message sums to the host
// mailbox (mapped pinned memory), waits for the Fiat-Shamir challenge and publishes it to the grid, which is also
// the barrier release.  The host side (interpolation, serialization, SHA-256) is the unchanged transcript code.
//
// Arithmetic is identical to k_fold_round / k_fold_round_sp.  Tables written inside the kernel are read back with
// ld.global.cg (L2) only: a buffer is rewritten every second round, so L1 could hold stale lines.
#pragma once
#include <cstdint>

#include "tail.cuh"

namespace scb {

struct PersistCtl {  // device memory, zeroed before every launch
    // barrier release + challenge in one: tagged words like the mailbox's (tag = challenge number, or kMbAbortTag
    // in word 0 when the hand-shaking CTA gave up), so a waiting CTA needs one L2 read and no fence
    uint64_t challenge[2 * kMaxLimbs];
    unsigned int ticket[kTailMaxRounds + 1];
};

template <class A, int K>
constexpr int persist_blocks() {
    return A::kLight ? (K <= 3 ? 5 : 4) : fold_min_blocks<A, K, 1>();
}

__device__ __forceinline__ void st_gpu(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// publish challenge number `tag` to the waiting CTAs (every word validates itself)
template <class A>
__device__ __forceinline__ void grid_publish(PersistCtl* ctl, uint32_t tag, const uint64_t* r) {
    constexpr int H = MailboxHalves<A>::value;
    const uint64_t hi = (uint64_t)tag << 32;
#pragma unroll
    for (int i = 0; i < A::N; ++i) {
        st_gpu(&ctl->challenge[i * H], hi | (uint32_t)r[i]);
        if constexpr (H == 2) st_gpu(&ctl->challenge[i * H + 1], hi | (r[i] >> 32));
    }
}

template <int W, bool NC>
__device__ __forceinline__ void ld_words_sel(const uint64_t* ptr, uint64_t* w) {
    if constexpr (NC) {
        ld_words<W>(ptr, w);
    } else if constexpr (W % 4 == 0) {
#pragma unroll
        for (int i = 0; i < W; i += 4)
            asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(w[i]), "=l"(w[i + 1]), "=l"(w[i + 2]), "=l"(w[i + 3])
                         : "l"(ptr + i)
                         : "memory");
    } else {
        ld_words_cg<W>(ptr, w);
    }
}

// Small-prime streaming pass: IN32 ? (8 packed entries -> 4 packed outputs) : (4 ark entries -> 2 packed outputs)
// per table per thread-iteration; one 256-bit load either way.
template <int K, bool IN32, bool NC>
__device__ __forceinline__ void persist_pass_sp(const PolSP& ar, const PolSP::FoldC& r, const uint64_t* const (&src)[K], uint64_t* const (&dst)[K],
                                                uint64_t n_groups, uint64_t start, uint64_t stride, PolSP::Acc (&acc)[K + 1]) {
    using A = PolSP;
    constexpr int NP = K + 1, QP = IN32 ? 2 : 1;
    for (uint64_t g = start; g < n_groups; g += stride) {
        uint32_t t[K][4 * QP];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint64_t w[4];
            ld_words_sel<4, NC>(src[k] + g * 4, w);
            if constexpr (IN32) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    t[k][2 * q] = (uint32_t)w[q];
                    t[k][2 * q + 1] = (uint32_t)(w[q] >> 32);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) t[k][q] = (uint32_t)w[q];
            }
        }
        uint32_t u[K][2 * QP];
#pragma unroll
        for (int q = 0; q < QP; ++q) {
            A::Lz prod[NP];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                u[k][2 * q] = ar.fold_c(t[k][4 * q], t[k][4 * q + 1], r);
                u[k][2 * q + 1] = ar.fold_c(t[k][4 * q + 2], t[k][4 * q + 3], r);
                pair_into_prod<A, NP>(ar, k == 0, u[k][2 * q], u[k][2 * q + 1], prod);
            }
#pragma unroll
            for (int x = 0; x < NP; ++x) ar.acc_add(acc[x], prod[x]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint64_t o[QP];
#pragma unroll
            for (int q = 0; q < QP; ++q) o[q] = (uint64_t)u[k][2 * q] | ((uint64_t)u[k][2 * q + 1] << 32);
            st_words<QP>(dst[k] + g * QP, o);
        }
    }
}

// Runs `n_rounds` fused fold+message rounds on tables of 2^m entries (m >= 2, n_rounds <= m - 1).  Same mailbox
// protocol as k_tail_rounds.  Must be launched cooperatively (all CTAs co-resident).  Once a round needs a single
// CTA, CTA 0 finishes the proof alone (no tickets, no flags: the single-CTA tail) and every other CTA exits.
template <class A, int K>
__global__ void __launch_bounds__(kThreads, (persist_blocks<A, K>()))
    k_persist_rounds(FieldDesc f, TabsIn<K> in0, TabsOut<K> buf_a, TabsOut<K> buf_b, ElemArg r0, uint32_t m, uint32_t n_rounds, int in0_w32,
                     TailMailbox* mb, PersistCtl* ctl, uint64_t* partials, uint64_t timeout_ns, PeerArg peer) {
    constexpr int NP = K + 1, N = A::N, AW = A::AW;
    const bool buf_w32 = A::kLight;
    bool src_w32 = A::kLight && in0_w32 != 0;
    const A ar(f);
    __shared__ uint64_t sm[32 * NP * AW];
    __shared__ uint64_t r_sm[kMaxLimbs];
    __shared__ int flag_sm;  // 1: this CTA took the last ticket, 2: abort
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) r_sm[i] = r0.w[i];
        if (blockIdx.x == 0) st_sys(&mb->stamp[2 * kTailMaxRounds + 1], globaltimer_ns());  // kernel start
    }
    __syncthreads();
    bool have_r = true;  // r_sm already holds this round's challenge (round 0, or fetched by this CTA running alone)
    for (uint32_t t = 0; t < n_rounds; ++t) {
        // thread-iterations of this round and the CTAs that take part in it
        const uint64_t n_quads = 1ull << (m - 2);
        const bool fast32 = A::kLight && src_w32 && n_quads >= 2;
        const uint64_t n_groups = fast32 ? n_quads / 2 : n_quads;
        uint64_t active = (n_groups + blockDim.x - 1) / blockDim.x;
        if (active > gridDim.x) active = gridDim.x;
        const bool solo = active == 1;
        if (solo && blockIdx.x != 0) return;
        if (!have_r) {  // barrier + challenge: released by the CTA that finished round t-1
            if (threadIdx.x == 0) {
                flag_sm = tagged_wait<A, false>(ctl->challenge, t, 4 * timeout_ns, r_sm) ? 2 : 0;
                asm volatile("fence.acq_rel.gpu;" ::: "memory");  // round t-1's folded entries before this round's loads
            }
            __syncthreads();
            if (flag_sm == 2) return;
        }
        const uint64_t* src[K];
        uint64_t* dst[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            src[k] = t == 0 ? in0.p[k] : ((t & 1) ? buf_a.p[k] : buf_b.p[k]);
            dst[k] = (t & 1) ? buf_b.p[k] : buf_a.p[k];
        }
        const typename A::FoldC r = ar.fold_const(ar.from_words(r_sm));
        typename A::Acc acc[NP];
#pragma unroll
        for (int x = 0; x < NP; ++x) ar.acc_zero(acc[x]);
        if (blockIdx.x < active) {
            const uint64_t start = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = active * blockDim.x;
            bool done = false;
            if constexpr (A::kLight) {
                if (fast32) {
                    if (t == 0) persist_pass_sp<K, true, true>(ar, r, src, dst, n_groups, start, stride, acc);
                    else persist_pass_sp<K, true, false>(ar, r, src, dst, n_groups, start, stride, acc);
                    done = true;
                } else if (!src_w32 && t == 0) {
                    persist_pass_sp<K, false, true>(ar, r, src, dst, n_groups, start, stride, acc);
                    done = true;
                }
            }
            if (!done) {
                for (uint64_t i = start; i < n_groups; i += stride) {
                    typename A::Lz prod[NP];
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        typename A::El t4[4], u0, u1;
                        if constexpr (A::kLight) {
                            tail_ld_quad<A>(ar, src[k], i, src_w32, t4);
                            u0 = ar.fold_c(t4[0], t4[1], r);
                            u1 = ar.fold_c(t4[2], t4[3], r);
                            tail_st_pair<A>(ar, dst[k], i, buf_w32, u0, u1);
                        } else {  // ark layout on both sides: vector loads (L2-coherent) and vector stores
                            uint64_t w[4 * N], o[2 * N];
                            ld_words_sel<4 * N, false>(src[k] + i * 4 * N, w);
#pragma unroll
                            for (int q = 0; q < 4; ++q) t4[q] = ar.from_words(w + q * N);
                            u0 = ar.fold_c(t4[0], t4[1], r);
                            u1 = ar.fold_c(t4[2], t4[3], r);
                            ar.to_words(u0, o);
                            ar.to_words(u1, o + N);
                            st_words<2 * N>(dst[k] + i * 2 * N, o);
                        }
                        pair_into_prod<A, NP>(ar, k == 0, u0, u1, prod);
                    }
#pragma unroll
                    for (int x = 0; x < NP; ++x) ar.acc_add(acc[x], prod[x]);
                }
            }
        }
        if (blockIdx.x < active) {
            __threadfence();  // folded entries are visible device-wide before this CTA's ticket / next round's loads
            block_reduce<A, NP>(ar, acc, sm);
            bool finisher = solo;
            if (!solo) {
                if (threadIdx.x == 0) {
#pragma unroll
                    for (int x = 0; x < NP; ++x) {
                        uint64_t w[AW];
                        ar.acc_to_words(acc[x], w);
#pragma unroll
                        for (int i = 0; i < AW; ++i) __stcg(&partials[((size_t)blockIdx.x * NP + x) * AW + i], w[i]);
                    }
                    __threadfence();
                    const unsigned int tk = atomicAdd(&ctl->ticket[t], 1u);
                    flag_sm = (tk == (unsigned int)active - 1) ? 1 : 0;
                }
                __syncthreads();
                finisher = flag_sm == 1;
                if (finisher) {  // last CTA of the round: add everybody's partial sums
                    __threadfence();
#pragma unroll
                    for (int x = 0; x < NP; ++x) ar.acc_zero(acc[x]);
                    for (unsigned int b = threadIdx.x; b < (unsigned int)active; b += blockDim.x) {
#pragma unroll
                        for (int x = 0; x < NP; ++x) {
                            uint64_t w[AW];
#pragma unroll
                            for (int i = 0; i < AW; ++i) w[i] = __ldcg(&partials[((size_t)b * NP + x) * AW + i]);
                            typename A::Acc o;
                            ar.acc_from_words(o, w);
                            ar.acc_merge(acc[x], o);
                        }
                    }
                    block_reduce<A, NP>(ar, acc, sm);
                }
            }
            if (finisher && threadIdx.x == 0) {  // hand the sums to the host, fetch the next challenge
                uint64_t w[NP][N];
#pragma unroll
                for (int x = 0; x < NP; ++x) ar.to_words(ar.msg_final(acc[x], K), w[x]);
                if (peer.world > 1) {  // sharded prover: add the peer GPUs' sums of this round (NVLink peer windows)
                    PeerArg pa = peer;
                    pa.seq += t;
                    peer_exchange_sum<A, NP>(ar, pa, w);
                }
                mailbox_post<A, NP>(mb, t + 1, w);
                st_sys(&mb->stamp[2 * t], globaltimer_ns());
                int bad = 0;
                if (t + 1 < n_rounds) {
                    bad = mailbox_wait_challenge<A>(mb, t + 1, timeout_ns, r_sm);
                    if (bad) {
                        st_gpu(&ctl->challenge[0], (uint64_t)kMbAbortTag << 32);
                        st_sys(&mb->dev_status, 2);
                    } else {
                        // every CTA fenced its stores before its ticket and this thread fenced after taking the
                        // last one, so the release below orders all of round t's folded entries before round t+1
                        if (!solo) grid_publish<A>(ctl, t + 1, r_sm);
                        st_sys(&mb->stamp[2 * t + 1], globaltimer_ns());
                    }
                } else {
                    st_sys(&mb->dev_status, 1);
                }
                if (solo) flag_sm = bad ? 2 : 0;
            }
            if (solo) {
                __syncthreads();  // r_sm holds the next challenge; this CTA's stores are ordered before its next loads
                if (flag_sm == 2) return;
            }
        }
        have_r = solo;
        m -= 1;
        src_w32 = buf_w32;
    }
}

}  // namespace scb
