// kernels.cuh -- the sm_100a kernels of the sum-check prover hot path.
//
// Every kernel is HBM-streaming integer work (no tensor cores): coalesced 128/256-bit loads
// (LDG.E.128 / LDG.E.256), Montgomery arithmetic in registers (field.cuh), warp-shuffle +
// shared-memory tree reduction, and a last-block-done final reduction so one launch yields the
// round's field sums.  Reference semantics per kernel are cited at each definition
// (paths relative to /root/reference).
#pragma once
#include <cstdint>

#include "field.cuh"

namespace scb {

constexpr int kMaxTables = 4;             // K <= 4  (degree <= 4 round polynomials)
constexpr int kMaxPts = kMaxTables + 1;   // sums at X = 0..K
constexpr int kThreads = 256;

template <int K>
struct TabsIn {
    const uint64_t* p[K];
};
template <int K>
struct TabsOut {
    uint64_t* p[K];
};
struct ElemArg {  // one field element passed by value
    uint64_t w[kMaxLimbs];
};

__device__ __forceinline__ uint64_t ld_sys(const volatile uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(volatile uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}


// ------------------------------------------------------------------------------------------
// Peer exchange window (multi-GPU).  Every rank owns one window in its own HBM, mapped into every
// peer through CUDA IPC; NVLink P2P stores write straight into it.  Word layout (uint64_t):
//   [0, 512)    sums  [slot 2][rank 8][32]    partial sums posted BY rank g: (d+1) limbs-wide round sums, or the
//                                             (K+1)^2 one-limb grid sums of the two-rounds-per-pass kernels
//   [512, 528)  flags [slot 2][rank 8]        sequence number of the post
//   [528, 536)  gather flags [rank 8]
//   [1024, ...) gather area (table slabs at consolidation)
// ------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 8;
constexpr int kWinSumStride = 32;                           // words per (slot, rank): >= kMaxPts * kMaxLimbs and >= 25
constexpr int kWinFlags = 2 * kMaxRanks * kWinSumStride;    // 512
constexpr int kWinGatherFlags = kWinFlags + 2 * kMaxRanks;  // 528
constexpr int kWinGatherWords = 1024;                       // gather area starts at byte 8192
static_assert(kWinSumStride >= kMaxPts * kMaxLimbs, "window slot too small");
struct PeerArg {
    uint64_t* win[kMaxRanks];  // win[g] = rank g's window as mapped in THIS process
    uint32_t rank;
    uint32_t world;            // <= 1: no exchange
    uint64_t seq;              // exchange number (monotone, identical on all ranks)
    uint64_t* status;          // mapped host word: set to 1 on timeout
    uint64_t timeout_ns;
};

// Posts this rank's NP final sums into every peer's window, waits for all peers' posts of the same exchange,
// and adds the rows mod p: the per-round "all-gather + modular sum" of the sharded prover, done by the finishing
// thread of the round kernel itself over NVLink peer memory (no separate collective launch).
template <class A, int NP>
__device__ __forceinline__ void peer_exchange_sum(const A& ar, const PeerArg& peer, uint64_t (&w)[NP][A::N]) {
    constexpr int N = A::N;
    const int slot = (int)(peer.seq & 1);
    for (uint32_t g = 0; g < peer.world; ++g) {
        uint64_t* dst = peer.win[g] + (slot * kMaxRanks + peer.rank) * kWinSumStride;
#pragma unroll
        for (int x = 0; x < NP; ++x)
#pragma unroll
            for (int i = 0; i < N; ++i) st_sys(dst + x * N + i, w[x][i]);
    }
    __threadfence_system();
    for (uint32_t g = 0; g < peer.world; ++g) st_sys(peer.win[g] + kWinFlags + slot * kMaxRanks + peer.rank, peer.seq);
    uint64_t* me = peer.win[peer.rank];
    const uint64_t t0 = globaltimer_ns();
    for (uint32_t g = 0; g < peer.world; ++g) {
        while (ld_sys(me + kWinFlags + slot * kMaxRanks + g) != peer.seq) {
            if (globaltimer_ns() - t0 > peer.timeout_ns) {
                st_sys(peer.status, 1);
                return;
            }
        }
    }
    __threadfence_system();
    for (uint32_t g = 0; g < peer.world; ++g) {
        if (g == peer.rank) continue;
        const uint64_t* src = me + (slot * kMaxRanks + g) * kWinSumStride;
#pragma unroll
        for (int x = 0; x < NP; ++x) {
            uint64_t o[N], a[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                o[i] = ld_sys(src + x * N + i);
                a[i] = w[x][i];
            }
            ar.to_words(ar.add(ar.from_words(a), ar.from_words(o)), w[x]);
        }
    }
}

// The same exchange done by a whole WARP (all 32 lanes of one warp call it with the rank's NP * N canonical words in
// shared memory `v`; the totals are left there): the world * NP * N remote stores go out in parallel instead of one
// after the other from a single thread (128 dependent-issue stores for 8 ranks x 16 grid sums were ~6 us of the ~12 us a
// sharded pass spent exchanging), lane g raises / polls rank g's flag, lane x adds element x of every peer's row.
template <class A, int NP>
__device__ __forceinline__ void peer_exchange_warp(const A& ar, const PeerArg& peer, uint64_t* v) {
    constexpr int N = A::N, W = NP * N;
    static_assert(W <= 32 && NP <= 32, "one lane per word / per element");
    const int lane = threadIdx.x & 31;
    const int slot = (int)(peer.seq & 1);
    for (uint32_t idx = lane; idx < peer.world * W; idx += 32) {
        const uint32_t g = idx / W, i = idx % W;
        st_sys(peer.win[g] + (slot * kMaxRanks + peer.rank) * kWinSumStride + i, v[i]);
    }
    __threadfence_system();
    __syncwarp();
    uint64_t* me = peer.win[peer.rank];
    if ((uint32_t)lane < peer.world) {
        st_sys(peer.win[lane] + kWinFlags + slot * kMaxRanks + peer.rank, peer.seq);
        const uint64_t t0 = globaltimer_ns();
        while (ld_sys(me + kWinFlags + slot * kMaxRanks + lane) != peer.seq) {
            if (globaltimer_ns() - t0 > peer.timeout_ns) {
                st_sys(peer.status, 1);
                break;
            }
        }
    }
    __threadfence_system();
    __syncwarp();
    if (lane < NP) {
        uint64_t a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = v[lane * N + i];
        for (uint32_t g = 0; g < peer.world; ++g) {
            if (g == peer.rank) continue;
            const uint64_t* src = me + (slot * kMaxRanks + g) * kWinSumStride + lane * N;
            uint64_t o[N];
#pragma unroll
            for (int i = 0; i < N; ++i) o[i] = ld_sys(src + i);
            ar.to_words(ar.add(ar.from_words(a), ar.from_words(o)), a);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) v[lane * N + i] = a[i];
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// block-wide reduction of NP accumulators; result valid in thread 0
// ------------------------------------------------------------------------------------------
template <class A, int NP>
__device__ __forceinline__ void block_reduce(const A& ar, typename A::Acc (&acc)[NP], uint64_t* sm) {
    constexpr int AW = A::AW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int x = 0; x < NP; ++x) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            uint64_t w[AW];
            ar.acc_to_words(acc[x], w);
#pragma unroll
            for (int i = 0; i < AW; ++i) w[i] = __shfl_xor_sync(0xffffffffu, w[i], off);
            typename A::Acc o;
            ar.acc_from_words(o, w);
            ar.acc_merge(acc[x], o);
        }
    }
    __syncthreads();  // sm may still be read by a previous call
    if (lane == 0) {
#pragma unroll
        for (int x = 0; x < NP; ++x) {
            uint64_t w[AW];
            ar.acc_to_words(acc[x], w);
#pragma unroll
            for (int i = 0; i < AW; ++i) sm[(warp * NP + x) * AW + i] = w[i];
        }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int x = 0; x < NP; ++x) {
            typename A::Acc a;
            if (lane < nwarps) {
                uint64_t w[AW];
#pragma unroll
                for (int i = 0; i < AW; ++i) w[i] = sm[(lane * NP + x) * AW + i];
                ar.acc_from_words(a, w);
            } else {
                ar.acc_zero(a);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                uint64_t w[AW];
                ar.acc_to_words(a, w);
#pragma unroll
                for (int i = 0; i < AW; ++i) w[i] = __shfl_xor_sync(0xffffffffu, w[i], off);
                typename A::Acc o;
                ar.acc_from_words(o, w);
                ar.acc_merge(a, o);
            }
            acc[x] = a;
        }
    }
}

// Grid-wide finish: every block publishes its partial sums; the last block to arrive (atomic
// ticket) adds the partials of all blocks and writes NP canonical elements to `out`
// (device memory or mapped pinned host memory).  Exact field addition is associative and
// commutative, so the result is bit-identical for any grid size or arrival order.
template <class A, int NP>
__device__ __forceinline__ void grid_reduce_finish(const A& ar, typename A::Acc (&acc)[NP], uint64_t* partials,
                                                   unsigned int* ticket, uint64_t* out, int msg_k = 0,
                                                   const PeerArg* peer = nullptr) {
    constexpr int AW = A::AW;
    __shared__ uint64_t sm[32 * NP * AW];
    __shared__ bool is_last;
    block_reduce<A, NP>(ar, acc, sm);
    if (gridDim.x > 1) {  // a single CTA already holds the total in thread 0: no partials, no ticket, no second tree
                          // (the small launches of a GKR layer spent most of their 11-14 us here)
        if (threadIdx.x == 0) {
#pragma unroll
            for (int x = 0; x < NP; ++x) {
                uint64_t w[AW];
                ar.acc_to_words(acc[x], w);
#pragma unroll
                for (int i = 0; i < AW; ++i) __stcg(&partials[((size_t)blockIdx.x * NP + x) * AW + i], w[i]);
            }
            __threadfence();
            unsigned int t = atomicAdd(ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
#pragma unroll
        for (int x = 0; x < NP; ++x) ar.acc_zero(acc[x]);
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
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
    if constexpr (NP * A::N <= 32) {
        if (peer != nullptr && peer->world > 1) {  // sharded: warp 0 exchanges the sums with the peer GPUs (CTA-uniform branch)
            __syncthreads();  // sm is free again
            if (threadIdx.x == 0) {
#pragma unroll
                for (int x = 0; x < NP; ++x) ar.to_words(ar.msg_final(acc[x], msg_k), sm + x * A::N);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                peer_exchange_warp<A, NP>(ar, *peer, sm);
                if (threadIdx.x < NP * A::N) out[threadIdx.x] = sm[threadIdx.x];
                __syncwarp();
                if (threadIdx.x == 0) {
                    *ticket = 0;
                    __threadfence_system();
                }
            }
            return;
        }
    }
    if (threadIdx.x == 0) {
        uint64_t w[NP][A::N];
#pragma unroll
        for (int x = 0; x < NP; ++x) ar.to_words(ar.msg_final(acc[x], msg_k), w[x]);
        if (peer != nullptr && peer->world > 1) peer_exchange_sum<A, NP>(ar, *peer, w);
#pragma unroll
        for (int x = 0; x < NP; ++x)
#pragma unroll
            for (int i = 0; i < A::N; ++i) out[x * A::N + i] = w[x][i];
        *ticket = 0;  // re-arm for the next launch on this stream
        __threadfence_system();
    }
}

// accumulate the round-message contribution of one hypercube pair of one table:
// values at X = 0..NP-1 are lo + X*(hi-lo), obtained by repeated addition (the reference's
// `two * a[i] - a[i-1]` at X = 2, matrix-multiplication/src/lib.rs:117-118, is the same element)
template <class A, int NP>
__device__ __forceinline__ void pair_into_prod(const A& ar, bool first, const typename A::El& lo,
                                               const typename A::El& hi, typename A::Lz (&prod)[NP]) {
    typename A::Lz d = ar.lz_diff(hi, lo);
    typename A::Lz v = ar.lz(lo);
#pragma unroll
    for (int x = 0; x < NP; ++x) {
        if (x == 1) v = ar.lz(hi);  // exact hi instead of lo + d keeps the lazy bound tight
        else if (x > 1) v = ar.lz_add(v, d);
        prod[x] = first ? v : ar.msg_mul(prod[x], v);
    }
}

// ------------------------------------------------------------------------------------------
// K2a  round message of a product of K dense MLEs over the same variables.
// Replaces G::to_univariate's pass over adjacent pairs (matrix-multiplication/src/lib.rs:110-122)
// generalised to K tables / X = 0..K.  PV adjacent hypercube pairs per thread-iteration
// (PV = 2 for one-limb fields: one 256-bit load per table per iteration).
// ------------------------------------------------------------------------------------------
template <class A, int K, int PV>
constexpr int round_min_blocks() {  // resident CTAs per SM the register budget is tuned for
    if (A::N > 1) return K <= 3 ? 2 : 1;
    if (A::kLight) return (PV == 1 || K <= 2) ? 8 : 6;
    return K <= 2 ? 6 : 4;
}
template <class A, int K, int U>
constexpr int fold_min_blocks() {
    if (A::N > 1) return K <= 3 ? 2 : 1;
    if (A::kLight) return U == 1 ? (K <= 3 ? 8 : 6) : 5;
    return K <= 2 ? 6 : (U == 1 ? 4 : 3);
}

template <class A, int K, int PV>
__global__ void __launch_bounds__(kThreads, (round_min_blocks<A, K, PV>())) k_round_evals(FieldDesc f, TabsIn<K> in, uint64_t n_groups,
                                                                            uint64_t* partials, unsigned int* ticket, uint64_t* out,
                                                                            PeerArg peer) {
    constexpr int NP = K + 1, N = A::N;
    const A ar(f);
    typename A::Acc acc[NP];
#pragma unroll
    for (int x = 0; x < NP; ++x) ar.acc_zero(acc[x]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_groups; i += stride) {
        if constexpr (N == 1) {
            uint64_t w[K][2 * PV];
#pragma unroll
            for (int k = 0; k < K; ++k) ld_words<2 * PV>(in.p[k] + i * 2 * PV, w[k]);
#pragma unroll
            for (int e = 0; e < PV; ++e) {
                typename A::Lz prod[NP];
#pragma unroll
                for (int k = 0; k < K; ++k)
                    pair_into_prod<A, NP>(ar, k == 0, ar.from_words(&w[k][2 * e]), ar.from_words(&w[k][2 * e + 1]), prod);
#pragma unroll
                for (int x = 0; x < NP; ++x) ar.acc_add(acc[x], prod[x]);
            }
        } else {
            typename A::Lz prod[NP];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint64_t w[2 * N];
                ld_words<2 * N>(in.p[k] + i * 2 * N, w);
                pair_into_prod<A, NP>(ar, k == 0, ar.from_words(w), ar.from_words(w + N), prod);
            }
#pragma unroll
            for (int x = 0; x < NP; ++x) ar.acc_add(acc[x], prod[x]);
        }
    }
    grid_reduce_finish<A, NP>(ar, acc, partials, ticket, out, K, &peer);
}

// ------------------------------------------------------------------------------------------
// K3+K2  fused fold(r) + next round message.  Replaces, for round j >= 1,
//   self.g = self.g.fix_variables(&[r_prev]); self.g.to_univariate()
// (sum-check-protocol/src/lib.rs:105-112) with ONE pass: each thread-iteration loads 4 adjacent
// entries of every table (one 256-bit load for one-limb fields), folds them to 2 ([ARK] fix_variables:
// t[b] = t[2b] + r (t[2b+1]-t[2b])), stores the 2 folded entries (canonical, one 128-bit store) and
// accumulates the message of that folded pair.  U independent quads per iteration (more loads in flight).
// Traffic: read M, write M/2 per table -- the 4*K*2^v*E total of SURVEY 8d.
// ------------------------------------------------------------------------------------------
template <class A, int K, int U>
__global__ void __launch_bounds__(kThreads, (fold_min_blocks<A, K, U>()))
    k_fold_round(FieldDesc f, TabsIn<K> in, TabsOut<K> outp, ElemArg rarg, uint64_t n_quads, uint64_t* partials, unsigned int* ticket,
                 uint64_t* out, PeerArg peer) {
    constexpr int NP = K + 1, N = A::N;
    static_assert(N == 1 || U == 1, "unrolling is only implemented for one-limb fields");
    const A ar(f);
    const typename A::FoldC r = ar.fold_const(ar.from_words(rarg.w));
    typename A::Acc acc[NP];
#pragma unroll
    for (int x = 0; x < NP; ++x) ar.acc_zero(acc[x]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_quads; i0 += stride * U) {
        if constexpr (N == 1) {
            uint64_t w[U][K][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t i = i0 + (uint64_t)u * stride;
                if (u == 0 || i < n_quads) {
#pragma unroll
                    for (int k = 0; k < K; ++k) ld_words<4>(in.p[k] + i * 4, w[u][k]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t i = i0 + (uint64_t)u * stride;
                if (u == 0 || i < n_quads) {
                    typename A::Lz prod[NP];
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        typename A::El u0 = ar.fold_c(ar.from_words(&w[u][k][0]), ar.from_words(&w[u][k][1]), r);
                        typename A::El u1 = ar.fold_c(ar.from_words(&w[u][k][2]), ar.from_words(&w[u][k][3]), r);
                        uint64_t o[2];
                        ar.to_words(u0, &o[0]);
                        ar.to_words(u1, &o[1]);
                        st_words<2>(outp.p[k] + i * 2, o);
                        pair_into_prod<A, NP>(ar, k == 0, u0, u1, prod);
                    }
#pragma unroll
                    for (int x = 0; x < NP; ++x) ar.acc_add(acc[x], prod[x]);
                }
            }
        } else {
            const uint64_t i = i0;
            typename A::Lz prod[NP];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint64_t w[4 * N];
                ld_words<4 * N>(in.p[k] + i * 4 * N, w);
                typename A::El u0 = ar.fold_c(ar.from_words(w), ar.from_words(w + N), r);
                typename A::El u1 = ar.fold_c(ar.from_words(w + 2 * N), ar.from_words(w + 3 * N), r);
                uint64_t o[2 * N];
                ar.to_words(u0, o);
                ar.to_words(u1, o + N);
                st_words<2 * N>(outp.p[k] + i * 2 * N, o);
                pair_into_prod<A, NP>(ar, k == 0, u0, u1, prod);
            }
#pragma unroll
            for (int x = 0; x < NP; ++x) ar.acc_add(acc[x], prod[x]);
        }
    }
    grid_reduce_finish<A, NP>(ar, acc, partials, ticket, out, K, &peer);
}

// ------------------------------------------------------------------------------------------
// K3  fold of one table: [ARK] DenseMultilinearExtension::fix_variables(&[r]).
// out has n_out = len/2 entries; VEC outputs per thread-iteration.
// ------------------------------------------------------------------------------------------
template <class A, int VEC>
__global__ void __launch_bounds__(kThreads) k_fold(FieldDesc f, const uint64_t* __restrict__ in, uint64_t* __restrict__ outp,
                                                   ElemArg rarg, uint64_t n_groups) {
    constexpr int N = A::N;
    const A ar(f);
    const typename A::FoldC r = ar.fold_const(ar.from_words(rarg.w));
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_groups; i += stride) {
        uint64_t w[2 * VEC * N], o[VEC * N];
        ld_words<2 * VEC * N>(in + i * 2 * VEC * N, w);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            ar.to_words(ar.fold_c(ar.from_words(w + (2 * e) * N), ar.from_words(w + (2 * e + 1) * N), r), o + e * N);
        st_words<VEC * N>(outp + i * VEC * N, o);
    }
}

// ------------------------------------------------------------------------------------------
// c_1 = sum over the hypercube of the product (Prover::new, sum-check-protocol/src/lib.rs:89,
// with G::to_evaluations matrix-multiplication/src/lib.rs:137-146) without materialising the
// 2^v-entry product table.  VEC entries per thread-iteration.
// ------------------------------------------------------------------------------------------
template <class A, int K, int VEC>
__global__ void __launch_bounds__(kThreads) k_product_sum(FieldDesc f, TabsIn<K> in, uint64_t n_groups, uint64_t* partials,
                                                          unsigned int* ticket, uint64_t* out) {
    constexpr int N = A::N;
    const A ar(f);
    typename A::Acc acc[1];
    ar.acc_zero(acc[0]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_groups; i += stride) {
        typename A::Lz prod[VEC];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint64_t w[VEC * N];
            ld_words<VEC * N>(in.p[k] + i * VEC * N, w);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                typename A::Lz v = ar.lz(ar.from_words(w + e * N));
                prod[e] = k == 0 ? v : ar.lz_mul(prod[e], v);
            }
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) ar.acc_add(acc[0], prod[e]);
    }
    grid_reduce_finish<A, 1>(ar, acc, partials, ticket, out);
}

// to_evaluations(): the elementwise product table (matrix-multiplication/src/lib.rs:137-146)
template <class A, int K>
__global__ void __launch_bounds__(kThreads) k_product_table(FieldDesc f, TabsIn<K> in, uint64_t* __restrict__ outp, uint64_t n) {
    constexpr int N = A::N;
    const A ar(f);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        typename A::El prod = ar.zero();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint64_t w[N];
            ld_words<N>(in.p[k] + i * N, w);
            typename A::El v = ar.from_words(w);
            prod = k == 0 ? v : ar.mul(prod, v);
        }
        uint64_t o[N];
        ar.to_words(prod, o);
        st_words<N>(outp + i * N, o);
    }
}

// ------------------------------------------------------------------------------------------
// K4  MLE evaluation through eq/chi tables built by doubling
// (multilinear-extensions/src/lib.rs:6-24; [ARK] DenseMultilinearExtension::evaluate for the
// LSB-first order).  The 2^v-entry chi table of the reference is never materialised: it is the
// outer product of a table over the low `lb` index bits and one over the high v-lb bits.
//
// The two tables are built by k_eq_tables (eqfix.cuh).
// ------------------------------------------------------------------------------------------
// k_mle_dot: sum_i evals[i] * lo[i & (2^lb-1)] * hi[i >> lb].  VEC consecutive entries (same row)
// per group: the VEC products with the low table are added up before the one multiplication by the row's high-table
// entry (1 + 1/VEC multiplications per entry).  U groups per thread-iteration with all their loads issued first;
// lo table staged in shared memory.
template <class A, int VEC, int U>
__global__ void __launch_bounds__(kThreads) k_mle_dot(FieldDesc f, const uint64_t* __restrict__ evals,
                                                      const uint64_t* __restrict__ lo_tab, const uint64_t* __restrict__ hi_tab,
                                                      uint32_t lb, uint64_t n_groups, uint64_t* partials, unsigned int* ticket,
                                                      uint64_t* out) {
    constexpr int N = A::N;
    extern __shared__ uint64_t lo_sm[];
    const A ar(f);
    const uint64_t lo_words = (1ull << lb) * N;
    for (uint64_t i = threadIdx.x; i < lo_words; i += blockDim.x) lo_sm[i] = lo_tab[i];
    __syncthreads();
    typename A::Acc acc[1];
    ar.acc_zero(acc[0]);
    const uint64_t groups_per_row_mask = ((1ull << lb) / VEC) - 1;
    const uint32_t row_shift = lb - (VEC == 4 ? 2 : (VEC == 2 ? 1 : 0));
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g0 < n_groups; g0 += stride * U) {
        uint64_t w[U][VEC * N];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t g = g0 + (uint64_t)u * stride;
            if (u == 0 || g < n_groups) ld_words<VEC * N>(evals + g * VEC * N, w[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t g = g0 + (uint64_t)u * stride;
            if (u == 0 || g < n_groups) {
                const uint64_t il = (g & groups_per_row_mask) * VEC;
                const uint64_t ih = g >> row_shift;
                typename A::Lz s;
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    typename A::Lz m = ar.lz_mul(ar.lz(ar.from_words(w[u] + e * N)), ar.lz(ar.from_words(lo_sm + (il + e) * N)));
                    s = e == 0 ? m : ar.lz_add(s, m);
                }
                uint64_t hw[N];
#pragma unroll
                for (int i = 0; i < N; ++i) hw[i] = __ldg(hi_tab + ih * N + i);
                ar.acc_add(acc[0], ar.lz_mul(s, ar.lz(ar.from_words(hw))));
            }
        }
    }
    grid_reduce_finish<A, 1>(ar, acc, partials, ticket, out);
}

// ------------------------------------------------------------------------------------------
// [ARK] DenseMultilinearExtension::relabel(a, b, k): swap index bit-blocks [a,a+k) <-> [b,b+k)
// (matrix-multiplication/src/lib.rs:82 uses relabel(0, n, n) = matrix transpose)
// ------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(kThreads) k_relabel(const uint64_t* __restrict__ in, uint64_t* __restrict__ outp, uint64_t n,
                                                      uint32_t a, uint32_t b, uint32_t k) {
    const uint64_t mask = (1ull << k) - 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t x = ((i >> a) ^ (i >> b)) & mask;
        const uint64_t j = i ^ ((x << a) | (x << b));
        uint64_t w[N];
        ld_words<N>(in + j * N, w);
        st_words<N>(outp + i * N, w);
    }
}

// ------------------------------------------------------------------------------------------
// synthetic tables (bench / tests): entry i = limbs splitmix64(((seed<<40)+start+i)*N + l), top
// limb masked to the modulus bit length, one conditional subtraction of p.  Same stream as
// oracle/pyoracle.py::synth_element and oracle/oracle.c::orc_synth_fill.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64_dev(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <int N>
__global__ void __launch_bounds__(kThreads) k_synth_fill(FieldDesc f, uint64_t seed, uint64_t start, uint64_t n, uint64_t* __restrict__ outp) {
    const uint32_t topbits = f.bits - 64 * (N - 1);
    const uint64_t topmask = topbits >= 64 ? ~0ull : ((1ull << topbits) - 1);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t v[N];
        const uint64_t base = ((seed << 40) + start + i) * N;
#pragma unroll
        for (int l = 0; l < N; ++l) v[l] = splitmix64_dev(base + l);
        v[N - 1] &= topmask;
        bool ge = true;
#pragma unroll
        for (int l = 0; l < N; ++l) {
            if (v[l] > f.p[l]) ge = true;
            else if (v[l] < f.p[l]) ge = false;
        }
        if (ge) {
            uint64_t borrow = 0;
#pragma unroll
            for (int l = 0; l < N; ++l) {
                uint64_t d = v[l] - f.p[l];
                uint64_t b1 = v[l] < f.p[l];
                uint64_t d2 = d - borrow;
                uint64_t b2 = d < borrow;
                v[l] = d2;
                borrow = b1 | b2;
            }
        }
        st_words<N>(outp + i * N, v);
    }
}


// ==========================================================================================
// Mixed-arity implementors: the tables are indexed by different variable
// subsets and stay L2-resident; work is integer-bound.  Values of a table pair at X = 0,1,2 are
// lo, hi, 2hi-lo; both polynomials have round degree 2, so three sums determine the
// same coefficients as the reference's 4-point IFFT (triangle-counting/src/lib.rs:120-131,
// gkr-protocol/src/round_polynomial.rs:78-89).
// ==========================================================================================
template <class A>
__device__ __forceinline__ typename A::El ld_el(const A& ar, const uint64_t* __restrict__ tab, uint64_t idx) {
    uint64_t w[A::N];
    ld_words<A::N>(tab + idx * A::N, w);
    return ar.from_words(w);
}
// lazy values of a linear function through (0, lo), (1, hi) at X = 0, 1, 2
template <class A>
__device__ __forceinline__ void lin3(const A& ar, const typename A::El& lo, const typename A::El& hi, typename A::Lz (&v)[3]) {
    v[0] = ar.lz(lo);
    v[1] = ar.lz(hi);
    v[2] = ar.lz_add(ar.lz(hi), ar.lz_diff(hi, lo));
}

// ---- triangle_counting::G -------------------------------------------------------------------
// State: f1 over (x: xn bits, y: yn bits) at index (y << xn) | x, f2 over (y, z) at (z << yn) | y,
// f3 over (x, z) at (z << xn) | x  (triangle-counting/src/lib.rs:150-157,170-172).
//
// c_1 = sum_{x,y,z} f1 f2 f3 (to_evaluations :138-165 summed, Prover::new) as
// sum_{x,z} f3(x,z) * sum_y f1(x,y) f2(y,z): one thread per (x,z), loop over y.
template <class A>
__global__ void __launch_bounds__(kThreads) k_triangle_sum(FieldDesc f, const uint64_t* __restrict__ f1, const uint64_t* __restrict__ f2,
                                                           const uint64_t* __restrict__ f3, uint32_t xn, uint32_t yn, uint32_t zn,
                                                           uint64_t* partials, unsigned int* ticket, uint64_t* out) {
    const A ar(f);
    typename A::Acc acc[1];
    ar.acc_zero(acc[0]);
    const uint64_t n_xz = 1ull << (xn + zn), xmask = (1ull << xn) - 1, ny = 1ull << yn;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_xz; t += stride) {
        const uint64_t x = t & xmask, z = t >> xn;
        typename A::Acc inner;
        ar.acc_zero(inner);
        for (uint64_t y = 0; y < ny; ++y)
            ar.acc_add(inner, ar.lz_mul(ar.lz(ld_el(ar, f1, (y << xn) | x)), ar.lz(ld_el(ar, f2, (z << yn) | y))));
        ar.acc_add(acc[0], ar.lz_mul(ar.lz(ar.acc_final(inner)), ar.lz(ld_el(ar, f3, t))));
    }
    grid_reduce_finish<A, 1>(ar, acc, partials, ticket, out);
}

// Round message at X = 0,1,2 over variable 0, which is an x bit while xn > 0 (f1, f3 depend on
// it), then a y bit (f1, f2), then a z bit (f2, f3) -- the fold schedule of :89-118.
//   x phase: thread per (x', z):  sum_y f1_X(x',y) f2(y,z), times f3_X(x',z)
//   y phase: thread per (y', z):  f1_X(y') f2_X(y',z) f3(z)
//   z phase: thread per z':       f1 * f2_X(z') f3_X(z')
template <class A>
__global__ void __launch_bounds__(kThreads) k_triangle_round(FieldDesc f, const uint64_t* __restrict__ f1, const uint64_t* __restrict__ f2,
                                                             const uint64_t* __restrict__ f3, uint32_t xn, uint32_t yn, uint32_t zn,
                                                             uint64_t* partials, unsigned int* ticket, uint64_t* out) {
    const A ar(f);
    typename A::Acc acc[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) ar.acc_zero(acc[x]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t tid0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (xn > 0) {
        const uint32_t xh = xn - 1;
        const uint64_t n_t = 1ull << (xh + zn), xmask = (1ull << xh) - 1, ny = 1ull << yn;
        for (uint64_t t = tid0; t < n_t; t += stride) {
            const uint64_t xp = t & xmask, z = t >> xh;
            typename A::Acc inner[3];
#pragma unroll
            for (int x = 0; x < 3; ++x) ar.acc_zero(inner[x]);
            for (uint64_t y = 0; y < ny; ++y) {
                typename A::Lz a[3];
                lin3(ar, ld_el(ar, f1, (y << xn) | (2 * xp)), ld_el(ar, f1, (y << xn) | (2 * xp + 1)), a);
                const typename A::Lz b = ar.lz(ld_el(ar, f2, (z << yn) | y));
#pragma unroll
                for (int x = 0; x < 3; ++x) ar.acc_add(inner[x], ar.lz_mul(a[x], b));
            }
            typename A::Lz c[3];
            lin3(ar, ld_el(ar, f3, (z << xn) | (2 * xp)), ld_el(ar, f3, (z << xn) | (2 * xp + 1)), c);
#pragma unroll
            for (int x = 0; x < 3; ++x) ar.acc_add(acc[x], ar.lz_mul(c[x], ar.lz(ar.acc_final(inner[x]))));
        }
    } else if (yn > 0) {
        const uint32_t yh = yn - 1;
        const uint64_t n_t = 1ull << (yh + zn), ymask = (1ull << yh) - 1;
        for (uint64_t t = tid0; t < n_t; t += stride) {
            const uint64_t yp = t & ymask, z = t >> yh;
            typename A::Lz a[3], b[3];
            lin3(ar, ld_el(ar, f1, 2 * yp), ld_el(ar, f1, 2 * yp + 1), a);
            lin3(ar, ld_el(ar, f2, (z << yn) | (2 * yp)), ld_el(ar, f2, (z << yn) | (2 * yp + 1)), b);
            const typename A::Lz c = ar.lz(ld_el(ar, f3, z));
#pragma unroll
            for (int x = 0; x < 3; ++x) ar.acc_add(acc[x], ar.lz_mul(ar.lz_mul(a[x], b[x]), c));
        }
    } else {
        const uint64_t n_t = 1ull << (zn - 1);
        const typename A::Lz a = ar.lz(ld_el(ar, f1, 0));
        for (uint64_t t = tid0; t < n_t; t += stride) {
            typename A::Lz b[3], c[3];
            lin3(ar, ld_el(ar, f2, 2 * t), ld_el(ar, f2, 2 * t + 1), b);
            lin3(ar, ld_el(ar, f3, 2 * t), ld_el(ar, f3, 2 * t + 1), c);
#pragma unroll
            for (int x = 0; x < 3; ++x) ar.acc_add(acc[x], ar.lz_mul(ar.lz_mul(b[x], c[x]), a));
        }
    }
    grid_reduce_finish<A, 3>(ar, acc, partials, ticket, out);
}

// to_evaluations() in the reference's order (x outer, y, z inner; :147-161)
template <class A>
__global__ void __launch_bounds__(kThreads) k_triangle_table(FieldDesc f, const uint64_t* __restrict__ f1, const uint64_t* __restrict__ f2,
                                                             const uint64_t* __restrict__ f3, uint32_t xn, uint32_t yn, uint32_t zn,
                                                             uint64_t* __restrict__ outp) {
    const A ar(f);
    const uint64_t n = 1ull << (xn + yn + zn);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t z = i & ((1ull << zn) - 1), y = (i >> zn) & ((1ull << yn) - 1), x = i >> (zn + yn);
        typename A::El v = ar.mul(ar.mul(ld_el(ar, f1, (y << xn) | x), ld_el(ar, f2, (z << yn) | y)), ld_el(ar, f3, (z << xn) | x));
        uint64_t o[A::N];
        ar.to_words(v, o);
        st_words<A::N>(outp + i * A::N, o);
    }
}

// ---- gkr_protocol::round_polynomial::W ------------------------------------------------------
// add, mul over (b: bn bits, c: cn bits) at index (c << bn) | b; w_b over b; w_c over c
// (gkr-protocol/src/round_polynomial.rs:96-118,123-125).
// term(b,c) = add (w_b + w_c) + mul (w_b w_c)
template <class A>
__global__ void __launch_bounds__(kThreads) k_gkrw_sum(FieldDesc f, const uint64_t* __restrict__ add, const uint64_t* __restrict__ mul,
                                                       const uint64_t* __restrict__ wb, const uint64_t* __restrict__ wc, uint32_t bn, uint32_t cn,
                                                       uint64_t* partials, unsigned int* ticket, uint64_t* out) {
    const A ar(f);
    typename A::Acc acc[1];
    ar.acc_zero(acc[0]);
    const uint64_t n = 1ull << (bn + cn), bmask = (1ull << bn) - 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const typename A::El b = ld_el(ar, wb, i & bmask), c = ld_el(ar, wc, i >> bn);
        ar.acc_add(acc[0], ar.lz_mul(ar.lz(ld_el(ar, add, i)), ar.lz(ar.add(b, c))));
        ar.acc_add(acc[0], ar.lz_mul(ar.lz(ld_el(ar, mul, i)), ar.lz_mul(ar.lz(b), ar.lz(c))));
    }
    grid_reduce_finish<A, 1>(ar, acc, partials, ticket, out);
}
// Round message at X = 0,1,2: variable 0 is a b bit while bn > 0 (add, mul, w_b fold), then a c bit
// (add, mul, w_c fold) -- the schedule of :59-76.
template <class A>
__global__ void __launch_bounds__(kThreads) k_gkrw_round(FieldDesc f, const uint64_t* __restrict__ add, const uint64_t* __restrict__ mul,
                                                         const uint64_t* __restrict__ wb, const uint64_t* __restrict__ wc, uint32_t bn, uint32_t cn,
                                                         uint64_t* partials, unsigned int* ticket, uint64_t* out) {
    const A ar(f);
    typename A::Acc acc[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) ar.acc_zero(acc[x]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_t = 1ull << (bn + cn - 1);
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_t; t += stride) {
        typename A::Lz a[3], m[3], vb[3], vc[3];
        lin3(ar, ld_el(ar, add, 2 * t), ld_el(ar, add, 2 * t + 1), a);
        lin3(ar, ld_el(ar, mul, 2 * t), ld_el(ar, mul, 2 * t + 1), m);
        if (bn > 0) {
            const uint64_t bp = t & ((1ull << (bn - 1)) - 1), c = t >> (bn - 1);
            lin3(ar, ld_el(ar, wb, 2 * bp), ld_el(ar, wb, 2 * bp + 1), vb);
            vc[0] = vc[1] = vc[2] = ar.lz(ld_el(ar, wc, c));
        } else {
            lin3(ar, ld_el(ar, wc, 2 * t), ld_el(ar, wc, 2 * t + 1), vc);
            vb[0] = vb[1] = vb[2] = ar.lz(ld_el(ar, wb, 0));
        }
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            ar.acc_add(acc[x], ar.lz_mul(a[x], ar.lz_add(vb[x], vc[x])));
            ar.acc_add(acc[x], ar.lz_mul(m[x], ar.lz_mul(vb[x], vc[x])));
        }
    }
    grid_reduce_finish<A, 3>(ar, acc, partials, ticket, out);
}
// to_evaluations() in the reference's order (b outer, c inner; :106-114)
template <class A>
__global__ void __launch_bounds__(kThreads) k_gkrw_table(FieldDesc f, const uint64_t* __restrict__ add, const uint64_t* __restrict__ mul,
                                                         const uint64_t* __restrict__ wb, const uint64_t* __restrict__ wc, uint32_t bn, uint32_t cn,
                                                         uint64_t* __restrict__ outp) {
    const A ar(f);
    const uint64_t n = 1ull << (bn + cn);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t c = i & ((1ull << cn) - 1), b = i >> cn;
        const uint64_t bc = (c << bn) | b;
        const typename A::El vb = ld_el(ar, wb, b), vc = ld_el(ar, wc, c);
        typename A::El v = ar.add(ar.mul(ld_el(ar, add, bc), ar.add(vb, vc)), ar.mul(ld_el(ar, mul, bc), ar.mul(vb, vc)));
        uint64_t o[A::N];
        ar.to_words(v, o);
        st_words<A::N>(outp + i * A::N, o);
    }
}

}  // namespace scb
