// gkr.cuh -- kernels of the linear-time GKR layer prover (BASELINE configs[4] "layered circuit,
// width 2^20, depth 16").
//
// The reference (gkr-protocol/src/lib.rs:373-436) materialises DENSE wiring tables add_i / mul_i of
// 2^(k_i + 2 k_{i+1}) entries -- 2^60 at width 2^20 -- and sum-checks
//     f(b, c) = add~(r_i, b, c) (W(b) + W(c)) + mul~(r_i, b, c) W(b) W(c)         (round_polynomial.rs:13-21)
// over (b, c) with the generic W polynomial.  The round messages are polynomials, so any way of computing the same
// sums gives the same coefficients; here the wiring stays a GATE LIST and each layer is two k-round sum-checks of
// the form  P*Q + S  over 2^k-entry tables (the two-phase algorithm of Xie et al. / Thaler sect. 4.6.6):
//   phase 1 (b rounds): f summed over c  = W(b) h1(b) + h2(b),
//        h1(b) = sum over gates a with in0 = b of eq(r_i, a) * (add ? 1 : W(in1_a)),
//        h2(b) = sum over add gates a with in0 = b of eq(r_i, a) * W(in1_a);
//   phase 2 (c rounds, b bound to u): f(u, c) = W(c) Q(c) + S(c),
//        A(c) = sum over add gates with in1 = c of eq(r_i, a) eq(u, in0_a),  M(c) likewise over mul gates,
//        Q = A + W~(u) M,  S = W~(u) A.
// Gates are grouped by input (CSR built once per circuit), so the tables are gathered, not scattered: no atomics, and
// the result is a deterministic exact field sum.
#pragma once
#include <cstdint>

#include "kernels.cuh"
#include "persist.cuh"

namespace scb {

struct EqPair {  // eq(point; idx) = lo[idx & (2^lb - 1)] * hi[idx >> lb]
    const uint64_t* lo;
    const uint64_t* hi;
    uint32_t lb;
};
// eq tables are small (at most 2^12 + 2^14 entries) and read by every thread of a gate-list kernel at random positions:
// L1-allocating loads (ld_el's streaming loads bypass L1: ncu showed a 0.03 % L1 hit rate in k_gkr_wiring_eval)
template <class A>
__device__ __forceinline__ typename A::El ld_el_cached(const A& ar, const uint64_t* __restrict__ tab, uint64_t idx) {
    uint64_t w[A::N];
#pragma unroll
    for (int q = 0; q < A::N; ++q) w[q] = __ldg(tab + idx * A::N + q);
    return ar.from_words(w);
}
template <class A>
__device__ __forceinline__ typename A::El eq_at(const A& ar, const EqPair& e, uint64_t idx) {
    return ar.mul(ld_el_cached(ar, e.lo, idx & ((1ull << e.lb) - 1)), ld_el_cached(ar, e.hi, idx >> e.lb));
}

// Circuit::evaluate, one layer (gkr-protocol/src/circuit.rs:108-116)
template <class A>
__global__ void __launch_bounds__(kThreads) k_gkr_eval_layer(FieldDesc f, const uint8_t* __restrict__ types, const uint32_t* __restrict__ in0,
                                                             const uint32_t* __restrict__ in1, const uint64_t* __restrict__ w_in,
                                                             uint64_t* __restrict__ w_out, uint64_t n_gates) {
    const A ar(f);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n_gates; a += stride) {
        const typename A::El x = ld_el(ar, w_in, in0[a]), y = ld_el(ar, w_in, in1[a]);
        uint64_t o[A::N];
        ar.to_words(types[a] ? ar.mul(x, y) : ar.add(x, y), o);
        st_words<A::N>(w_out + a * A::N, o);
    }
}

// phase-1 tables h1, h2 over b (one thread per b, its gates through the CSR by in0).  Position t of the CSR order gives
// the gate id and type (idx0[t], type in bit 31) and the gate's other input (oth0[t] = in1 of that gate) from coalesced
// arrays, so the only gathers left are the eq lookups and W[in1].
constexpr uint32_t kGateIdMask = 0x7fffffffu;
template <class A>
__global__ void __launch_bounds__(kThreads) k_gkr_phase1(FieldDesc f, EqPair eq_r, const uint32_t* __restrict__ off0, const uint32_t* __restrict__ idx0,
                                                         const uint32_t* __restrict__ oth0,
                                                         const uint64_t* __restrict__ w, uint64_t* __restrict__ h1, uint64_t* __restrict__ h2,
                                                         uint64_t n_b) {
    const A ar(f);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_b; b += stride) {
        typename A::El s1 = ar.zero(), s2 = ar.zero();
        for (uint32_t t = off0[b]; t < off0[b + 1]; ++t) {
            const uint32_t at = idx0[t], a = at & kGateIdMask;
            const typename A::El e = eq_at(ar, eq_r, a);
            const typename A::El ew = ar.mul(e, ld_el(ar, w, oth0[t]));
            if (at >> 31) {
                s1 = ar.add(s1, ew);
            } else {
                s1 = ar.add(s1, e);
                s2 = ar.add(s2, ew);
            }
        }
        uint64_t o[A::N];
        ar.to_words(s1, o);
        st_words<A::N>(h1 + b * A::N, o);
        ar.to_words(s2, o);
        st_words<A::N>(h2 + b * A::N, o);
    }
}

// phase-2 tables Q = A + wu*M, S = wu*A over c (one thread per c, its gates through the CSR by in1)
template <class A>
__global__ void __launch_bounds__(kThreads) k_gkr_phase2(FieldDesc f, EqPair eq_r, EqPair eq_u, const uint32_t* __restrict__ off1,
                                                         const uint32_t* __restrict__ idx1,
                                                         const uint32_t* __restrict__ oth1, ElemArg wu_arg, uint64_t* __restrict__ q,
                                                         uint64_t* __restrict__ s, uint64_t n_c, const uint64_t* wu_dev = nullptr) {
    const A ar(f);
    uint64_t wu_w[A::N];  // W~(u): by value, or left in device memory by the previous kernel of a batched layer proof
#pragma unroll
    for (int i = 0; i < A::N; ++i) wu_w[i] = wu_dev ? __ldcg(wu_dev + i) : wu_arg.w[i];
    const typename A::El wu = ar.from_words(wu_w);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_c; c += stride) {
        typename A::El sa = ar.zero(), sm = ar.zero();
        for (uint32_t t = off1[c]; t < off1[c + 1]; ++t) {
            const uint32_t at = idx1[t], a = at & kGateIdMask;
            const typename A::El e = ar.mul(eq_at(ar, eq_r, a), eq_at(ar, eq_u, oth1[t]));  // oth1[t] = in0 of that gate
            if (at >> 31) sm = ar.add(sm, e);
            else sa = ar.add(sa, e);
        }
        uint64_t o[A::N];
        ar.to_words(ar.add(sa, ar.mul(wu, sm)), o);
        st_words<A::N>(q + c * A::N, o);
        ar.to_words(ar.mul(wu, sa), o);
        st_words<A::N>(s + c * A::N, o);
    }
}

// ---- small-prime fields: the same two tables by SCATTER (r2b).  The CSR kernels above are one thread per output with a
// loop over its gates: three dependent levels of uncoalesced loads, a third of the lanes active on average, 83-90 % of
// the stall samples on the load scoreboard (68-74 us per launch at width 2^20).  For p < 2^28 a sum of 2^26 canonical
// residues fits 64 bits with room to spare, so a thread per GATE (coalesced gate list, eq(r; a) coalesced in a, one
// gather of W) adds its contribution with a 64-bit integer atomic -- exact and order-independent, so the tables are the
// same field elements -- and a second small kernel reduces modulo p (phase 2: and forms Q, S).  The destination tables
// must be zero on entry.
__device__ __forceinline__ uint32_t eq_at_sp(const PolSP& ar, const EqPair& e, uint64_t idx) {  // L1-allocating loads: small, hot tables
    return ar.mul((uint32_t)__ldg(e.lo + (idx & ((1ull << e.lb) - 1))), (uint32_t)__ldg(e.hi + (idx >> e.lb)));
}
__global__ void __launch_bounds__(kThreads) k_gkr_phase1_scatter_sp(FieldDesc f, EqPair eq_r, const uint8_t* __restrict__ types,
                                                                    const uint32_t* __restrict__ in0, const uint32_t* __restrict__ in1,
                                                                    const uint64_t* __restrict__ w, unsigned long long* __restrict__ h1,
                                                                    unsigned long long* __restrict__ h2, uint64_t n_gates) {
    const PolSP ar(f);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n_gates; a += stride) {
        const uint32_t b = in0[a];
        const uint32_t e = eq_at_sp(ar, eq_r, a);
        const uint32_t ew = ar.mul(e, (uint32_t)w[in1[a]]);
        if (types[a]) {
            atomicAdd(h1 + b, (unsigned long long)ew);
        } else {
            atomicAdd(h1 + b, (unsigned long long)e);
            atomicAdd(h2 + b, (unsigned long long)ew);
        }
    }
}
__global__ void __launch_bounds__(kThreads) k_gkr_phase1_finish_sp(FieldDesc f, uint64_t* __restrict__ h1, uint64_t* __restrict__ h2, uint64_t n_b) {
    const uint64_t p = f.p[0];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_b; b += stride) {
        h1[b] %= p;
        h2[b] %= p;
    }
}
// phase 2: sa[c] (into q) and sm[c] (into s) as integer sums, then Q = sa + wu sm, S = wu sa in place
__global__ void __launch_bounds__(kThreads) k_gkr_phase2_scatter_sp(FieldDesc f, EqPair eq_r, EqPair eq_u, const uint8_t* __restrict__ types,
                                                                    const uint32_t* __restrict__ in0, const uint32_t* __restrict__ in1,
                                                                    unsigned long long* __restrict__ q, unsigned long long* __restrict__ s, uint64_t n_gates) {
    const PolSP ar(f);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n_gates; a += stride) {
        const uint32_t e = ar.mul(eq_at_sp(ar, eq_r, a), eq_at_sp(ar, eq_u, in0[a]));
        atomicAdd((types[a] ? s : q) + in1[a], (unsigned long long)e);
    }
}
__global__ void __launch_bounds__(kThreads) k_gkr_phase2_finish_sp(FieldDesc f, ElemArg wu_arg, uint64_t* __restrict__ q, uint64_t* __restrict__ s,
                                                                   uint64_t n_c, const uint64_t* wu_dev) {
    const PolSP ar(f);
    const uint32_t wu = (uint32_t)(wu_dev ? __ldcg(wu_dev) : wu_arg.w[0]);
    const uint64_t p = f.p[0];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_c; c += stride) {
        const uint32_t sa = (uint32_t)(q[c] % p), sm = (uint32_t)(s[c] % p);
        q[c] = ar.add(sa, ar.mul(wu, sm));
        s[c] = ar.mul(wu, sa);
    }
}

// add~_i(r, b*, c*) and mul~_i(r, b*, c*) from the gate list (verifier's final check, gkr-protocol/src/lib.rs:155):
// out[0] = sum over add gates of eq(r,a) eq(b*,in0) eq(c*,in1), out[1] the same over mul gates
template <class A>
__global__ void __launch_bounds__(kThreads) k_gkr_wiring_eval(FieldDesc f, EqPair eq_r, EqPair eq_b, EqPair eq_c, const uint8_t* __restrict__ types,
                                                              const uint32_t* __restrict__ in0, const uint32_t* __restrict__ in1, uint64_t n_gates,
                                                              uint64_t* partials, unsigned int* ticket, uint64_t* out) {
    const A ar(f);
    typename A::Acc acc[2];
    ar.acc_zero(acc[0]);
    ar.acc_zero(acc[1]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n_gates; a += stride) {
        const typename A::El e = ar.mul(ar.mul(eq_at(ar, eq_r, a), eq_at(ar, eq_b, in0[a])), eq_at(ar, eq_c, in1[a]));
        if (types[a]) ar.acc_add(acc[1], ar.lz(e));
        else ar.acc_add(acc[0], ar.lz(e));
    }
    grid_reduce_finish<A, 2>(ar, acc, partials, ticket, out);
}

// ---- sum-check of g = P*Q + S over tables of the same variables (degree 2: sums at X = 0, 1, 2) ----
template <class A>
__device__ __forceinline__ void pqs_accumulate(const A& ar, const typename A::El (&p)[2], const typename A::El (&q)[2],
                                               const typename A::El (&s)[2], typename A::Acc (&acc)[3]) {
    typename A::Lz pv[3], qv[3], sv[3];
    lin3(ar, p[0], p[1], pv);
    lin3(ar, q[0], q[1], qv);
    lin3(ar, s[0], s[1], sv);
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        ar.acc_add(acc[x], ar.lz_mul(pv[x], qv[x]));
        ar.acc_add(acc[x], sv[x]);
    }
}
template <class A>
__global__ void __launch_bounds__(kThreads) k_pqs_round(FieldDesc f, const uint64_t* __restrict__ P, const uint64_t* __restrict__ Q,
                                                        const uint64_t* __restrict__ S, uint64_t n_pairs, uint64_t* partials, unsigned int* ticket,
                                                        uint64_t* out) {
    const A ar(f);
    typename A::Acc acc[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) ar.acc_zero(acc[x]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
        const typename A::El p[2] = {ld_el(ar, P, 2 * i), ld_el(ar, P, 2 * i + 1)};
        const typename A::El q[2] = {ld_el(ar, Q, 2 * i), ld_el(ar, Q, 2 * i + 1)};
        const typename A::El s[2] = {ld_el(ar, S, 2 * i), ld_el(ar, S, 2 * i + 1)};
        pqs_accumulate(ar, p, q, s, acc);
    }
    grid_reduce_finish<A, 3>(ar, acc, partials, ticket, out);
}
// fused: fold the three tables by r (t[b] = t[2b] + r (t[2b+1] - t[2b])) and accumulate the next message
template <class A>
__global__ void __launch_bounds__(kThreads) k_pqs_fold_round(FieldDesc f, const uint64_t* __restrict__ P, const uint64_t* __restrict__ Q,
                                                             const uint64_t* __restrict__ S, uint64_t* __restrict__ Po, uint64_t* __restrict__ Qo,
                                                             uint64_t* __restrict__ So, ElemArg rarg, uint64_t n_quads, uint64_t* partials,
                                                             unsigned int* ticket, uint64_t* out) {
    const A ar(f);
    const typename A::El r = ar.from_words(rarg.w);
    typename A::Acc acc[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) ar.acc_zero(acc[x]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += stride) {
        typename A::El v[3][2];
        const uint64_t* in[3] = {P, Q, S};
        uint64_t* outp[3] = {Po, Qo, So};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                v[k][h] = ar.fold(ld_el(ar, in[k], 4 * i + 2 * h), ld_el(ar, in[k], 4 * i + 2 * h + 1), r);
                uint64_t o[A::N];
                ar.to_words(v[k][h], o);
                st_words<A::N>(outp[k] + (2 * i + h) * A::N, o);
            }
        }
        pqs_accumulate(ar, v[0], v[1], v[2], acc);
    }
    grid_reduce_finish<A, 3>(ar, acc, partials, ticket, out);
}

// ---- R rounds per pass, challenges known up front (scb_gkr_prover_prove_layer) ----
// A layer's challenges are public coins: they do not depend on the prover's messages, so when they are handed over up
// front nothing but the DATA dependency orders the rounds -- and a block of 2^R consecutive entries of a table is closed
// under R rounds of pairing and folding.  One thread-iteration loads such a block of P, Q and S, accumulates message t
// from its 2^(R-1) pairs, folds by challenge t, accumulates message t+1 from the 2^(R-2) folded pairs, ... and stores
// ONE entry per table: R messages (3 R sums) per pass over the tables instead of one, no barrier between them, and the
// table shrinks by 2^R per launch.  A k-round phase is ceil(k / R) ordinary launches (R = 4: five for k = 20) instead of
// k grid-wide barriers of ~8 us each.  Same sums, same field elements (gkr-protocol/src/round_polynomial.rs:59-90).
template <class A, int R>
__global__ void __launch_bounds__(kThreads) k_pqs_multi(FieldDesc f, const uint64_t* __restrict__ P, const uint64_t* __restrict__ Q,
                                                        const uint64_t* __restrict__ S, uint64_t* __restrict__ Po, uint64_t* __restrict__ Qo,
                                                        uint64_t* __restrict__ So, const uint64_t* __restrict__ challenges, uint64_t n_blocks,
                                                        uint64_t* partials, unsigned int* ticket, uint64_t* out) {
    constexpr int N = A::N, B = 1 << R;
    const A ar(f);
    typename A::El r[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        uint64_t rw[N];
#pragma unroll
        for (int i = 0; i < N; ++i) rw[i] = __ldg(challenges + (size_t)s * N + i);
        r[s] = ar.from_words(rw);
    }
    typename A::Acc acc[3 * R];
#pragma unroll
    for (int x = 0; x < 3 * R; ++x) ar.acc_zero(acc[x]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_blocks; i += stride) {
        typename A::El v[3][B];
        const uint64_t* in[3] = {P, Q, S};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if constexpr (N == 1 && B >= 4) {  // 256-bit loads: a thread's block is contiguous (8-byte loads would re-fetch each sector)
#pragma unroll
                for (int e = 0; e < B; e += 4) {
                    uint64_t w[4];
                    ld_words<4>(in[k] + (i * B + e), w);
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[k][e + q] = ar.from_words(&w[q]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < B; ++e) v[k][e] = ld_el(ar, in[k], i * B + e);
            }
        }
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const int pairs = B >> (s + 1);
            typename A::Acc a3[3] = {acc[3 * s], acc[3 * s + 1], acc[3 * s + 2]};
#pragma unroll
            for (int j = 0; j < pairs; ++j) {
                const typename A::El pp[2] = {v[0][2 * j], v[0][2 * j + 1]};
                const typename A::El qq[2] = {v[1][2 * j], v[1][2 * j + 1]};
                const typename A::El ss[2] = {v[2][2 * j], v[2][2 * j + 1]};
                pqs_accumulate(ar, pp, qq, ss, a3);
            }
            acc[3 * s] = a3[0];
            acc[3 * s + 1] = a3[1];
            acc[3 * s + 2] = a3[2];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int j = 0; j < pairs; ++j) v[k][j] = ar.fold(v[k][2 * j], v[k][2 * j + 1], r[s]);
        }
        uint64_t* outp[3] = {Po, Qo, So};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            uint64_t o[N];
            ar.to_words(v[k][0], o);
            st_words<N>(outp[k] + i * N, o);
        }
    }
    grid_reduce_finish<A, 3 * R>(ar, acc, partials, ticket, out);
}

// ---- the last rounds of a phase in ONE CTA (r2b): once the three tables have at most 2^kPqsTailMaxVars entries they fit
// shared memory, and the remaining rounds (message, fold, message, ...) need nothing but block barriers -- one launch
// instead of ceil(m / R) launches of k_pqs_multi whose 11-14 us each were launch and reduction latency, not work.
// Round t (t = 0 .. m-1): message of the tables as they are (three sums to out_slots[3 t ..]), then the fold by
// challenges[t]; the single entries left after the last fold go to Po / Qo / So (W~(u) for phase 2 of a GKR layer).
// One-limb fields.  The fold is done in place in chunks of blockDim pairs in increasing order: chunk i reads entries
// [2 i B, 2 (i + 1) B) and writes [i B, (i + 1) B), which no later chunk reads.
constexpr int kPqsTailMaxVars = 12;  // 3 x 2^12 x 8 B = 96 KB of shared memory (k = 20: two k_pqs_multi passes, then the tail)
template <class A>
__global__ void __launch_bounds__(kThreads) k_pqs_tail(FieldDesc f, const uint64_t* __restrict__ P, const uint64_t* __restrict__ Q,
                                                       const uint64_t* __restrict__ S, uint32_t m, const uint64_t* __restrict__ challenges,
                                                       uint64_t* out_slots, uint64_t* __restrict__ Po, uint64_t* __restrict__ Qo, uint64_t* __restrict__ So) {
    static_assert(A::N == 1, "one-limb fields only");
    constexpr int AW = A::AW;
    extern __shared__ uint64_t tail_sm[];  // [3][2^m]
    __shared__ uint64_t red_sm[32 * 3 * AW];
    const A ar(f);
    const uint32_t n0 = 1u << m;
    uint64_t* T[3] = {tail_sm, tail_sm + n0, tail_sm + 2 * n0};
    const uint64_t* in[3] = {P, Q, S};
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (uint32_t i = threadIdx.x; i < n0; i += blockDim.x) T[k][i] = in[k][i];
    __syncthreads();
    for (uint32_t t = 0; t < m; ++t) {
        const uint32_t pairs = n0 >> (t + 1);
        typename A::Acc acc[3];
#pragma unroll
        for (int x = 0; x < 3; ++x) ar.acc_zero(acc[x]);
        for (uint32_t j = threadIdx.x; j < pairs; j += blockDim.x) {
            const typename A::El pp[2] = {ar.from_words(T[0] + 2 * j), ar.from_words(T[0] + 2 * j + 1)};
            const typename A::El qq[2] = {ar.from_words(T[1] + 2 * j), ar.from_words(T[1] + 2 * j + 1)};
            const typename A::El ss[2] = {ar.from_words(T[2] + 2 * j), ar.from_words(T[2] + 2 * j + 1)};
            pqs_accumulate(ar, pp, qq, ss, acc);
        }
        block_reduce<A, 3>(ar, acc, red_sm);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int x = 0; x < 3; ++x) ar.to_words(ar.msg_final(acc[x], 0), out_slots + (size_t)(3 * t + x));
        }
        const uint64_t rw = __ldg(challenges + t);
        const typename A::El r = ar.from_words(&rw);
        for (uint32_t j0 = 0; j0 < pairs; j0 += blockDim.x) {  // in-place fold, chunk by chunk
            const uint32_t j = j0 + threadIdx.x;
            typename A::El v[3];
            if (j < pairs) {
#pragma unroll
                for (int k = 0; k < 3; ++k) v[k] = ar.fold(ar.from_words(T[k] + 2 * j), ar.from_words(T[k] + 2 * j + 1), r);
            }
            __syncthreads();  // every read of this chunk (and the message pass above) is done
            if (j < pairs) {
#pragma unroll
                for (int k = 0; k < 3; ++k) ar.to_words(v[k], T[k] + j);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Po[0] = T[0][0];
        Qo[0] = T[1][0];
        So[0] = T[2][0];
        __threadfence_system();  // out_slots may be mapped host memory
    }
}

// ---- all rounds of one P*Q + S sum-check in one cooperative launch, challenges known up front ----
// Round 0 is the message of the tables as they are; round t >= 1 folds by challenges[t-1] and accumulates the next
// message (the bodies of k_pqs_round / k_pqs_fold_round).  The CTAs meet at a ticket/flag barrier in device memory
// between rounds (PersistCtl, persist.cuh); the CTA that takes the last ticket writes the three sums of the round to
// its slot and releases the next round.  Nothing talks to the host: the launch is fully asynchronous.  Tables written
// inside the kernel are read back through L2 (ld.global.cg).  final_fold: after the last message, P (two entries by
// then) is folded by challenges[n_msgs-1] into out_final -- W~(u) for the second phase of a GKR layer.
template <class A>
__device__ __forceinline__ typename A::El ld_el_cg(const A& ar, const uint64_t* tab, uint64_t idx) {
    uint64_t w[A::N];
    ld_words_cg<A::N>(tab + idx * A::N, w);
    return ar.from_words(w);
}
template <class A>
constexpr int pqs_persist_blocks() {
    return A::kLight ? 4 : (A::N == 1 ? 3 : 1);
}
template <class A>
__global__ void __launch_bounds__(kThreads, (pqs_persist_blocks<A>()))
    k_pqs_persist(FieldDesc f, const uint64_t* P0, const uint64_t* Q0, const uint64_t* S0, uint64_t* buf_a, uint64_t* buf_b,
                  const uint64_t* challenges, uint32_t m, uint32_t n_msgs, int final_fold, uint64_t* out_slots, uint64_t* out_final, PersistCtl* ctl,
                  uint64_t* partials) {
    constexpr int N = A::N, AW = A::AW;
    const A ar(f);
    __shared__ uint64_t sm[32 * 3 * AW];
    __shared__ int flag_sm;
    const uint64_t cap_a = (m >= 1 ? (1ull << (m - 1)) : 1) * N, cap_b = (m >= 2 ? (1ull << (m - 2)) : 1) * N;  // words per table
    bool have_release = true;  // round 0 needs no barrier; a CTA running alone needs none either
    for (uint32_t t = 0; t < n_msgs; ++t) {
        const uint32_t mt = t == 0 ? m : m - (t - 1);                        // variables of the tables this round reads
        const uint64_t n_items = t == 0 ? (mt >= 1 ? 1ull << (mt - 1) : 1)  // pairs
                                        : (mt >= 2 ? 1ull << (mt - 2) : 1); // quads
        uint64_t active = (n_items + blockDim.x - 1) / blockDim.x;
        if (active > gridDim.x) active = gridDim.x;
        const bool solo = active == 1;
        if (solo && blockIdx.x != 0) return;
        if (!have_release) {
            if (threadIdx.x == 0) {
                const uint64_t t0 = globaltimer_ns();
                int bad = 0;
                for (;;) {
                    uint64_t c;
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(c) : "l"(ctl->challenge) : "memory");
                    if ((uint32_t)(c >> 32) == t) break;
                    if (globaltimer_ns() - t0 > 2000000000ull) {  // a CTA died: do not hang the device
                        bad = 1;
                        break;
                    }
                }
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                flag_sm = bad ? 2 : 0;
            }
            __syncthreads();
            if (flag_sm == 2) return;
        }
        typename A::Acc acc[3];
#pragma unroll
        for (int x = 0; x < 3; ++x) ar.acc_zero(acc[x]);
        if (blockIdx.x < active) {
            const uint64_t start = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = active * blockDim.x;
            if (t == 0) {
                for (uint64_t i = start; i < n_items; i += stride) {
                    const typename A::El p[2] = {ld_el(ar, P0, 2 * i), ld_el(ar, P0, 2 * i + 1)};
                    const typename A::El q[2] = {ld_el(ar, Q0, 2 * i), ld_el(ar, Q0, 2 * i + 1)};
                    const typename A::El s[2] = {ld_el(ar, S0, 2 * i), ld_el(ar, S0, 2 * i + 1)};
                    pqs_accumulate(ar, p, q, s, acc);
                }
            } else {
                uint64_t rw[N];
#pragma unroll
                for (int i = 0; i < N; ++i) rw[i] = __ldg(challenges + (size_t)(t - 1) * N + i);
                const typename A::El r = ar.from_words(rw);
                // t == 1 reads the inputs, then the buffers alternate: round t writes buf_a when t is odd
                const uint64_t* in[3];
                uint64_t* outp[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint64_t* src0 = k == 0 ? P0 : (k == 1 ? Q0 : S0);
                    in[k] = t == 1 ? src0 : ((t & 1) ? buf_b + k * cap_b : buf_a + k * cap_a);
                    outp[k] = (t & 1) ? buf_a + k * cap_a : buf_b + k * cap_b;
                }
                for (uint64_t i = start; i < n_items; i += stride) {
                    typename A::El v[3][2];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            v[k][h] = ar.fold(ld_el_cg(ar, in[k], 4 * i + 2 * h), ld_el_cg(ar, in[k], 4 * i + 2 * h + 1), r);
                            uint64_t o[N];
                            ar.to_words(v[k][h], o);
                            st_words<N>(outp[k] + (2 * i + h) * N, o);
                        }
                    }
                    pqs_accumulate(ar, v[0], v[1], v[2], acc);
                }
            }
            __threadfence();
            block_reduce<A, 3>(ar, acc, sm);
            bool finisher = solo;
            if (!solo) {
                if (threadIdx.x == 0) {
#pragma unroll
                    for (int x = 0; x < 3; ++x) {
                        uint64_t w[AW];
                        ar.acc_to_words(acc[x], w);
#pragma unroll
                        for (int i = 0; i < AW; ++i) __stcg(&partials[((size_t)blockIdx.x * 3 + x) * AW + i], w[i]);
                    }
                    __threadfence();
                    const unsigned int tk = atomicAdd(&ctl->ticket[t], 1u);
                    flag_sm = (tk == (unsigned int)active - 1) ? 1 : 0;
                }
                __syncthreads();
                finisher = flag_sm == 1;
                if (finisher) {
                    __threadfence();
#pragma unroll
                    for (int x = 0; x < 3; ++x) ar.acc_zero(acc[x]);
                    for (unsigned int b = threadIdx.x; b < (unsigned int)active; b += blockDim.x) {
#pragma unroll
                        for (int x = 0; x < 3; ++x) {
                            uint64_t w[AW];
#pragma unroll
                            for (int i = 0; i < AW; ++i) w[i] = __ldcg(&partials[((size_t)b * 3 + x) * AW + i]);
                            typename A::Acc o;
                            ar.acc_from_words(o, w);
                            ar.acc_merge(acc[x], o);
                        }
                    }
                    block_reduce<A, 3>(ar, acc, sm);
                }
            }
            if (finisher && threadIdx.x == 0) {
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    uint64_t w[N];
                    ar.to_words(ar.acc_final(acc[x]), w);
#pragma unroll
                    for (int i = 0; i < N; ++i) out_slots[((size_t)t * 3 + x) * N + i] = w[i];
                }
                if (t + 1 < n_msgs) {
                    if (!solo) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ctl->challenge), "l"((uint64_t)(t + 1) << 32) : "memory");
                } else {
                    if (final_fold) {  // the tables have one variable left: P~ = P[0] + r (P[1] - P[0])
                        uint64_t rw[N], o[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) rw[i] = __ldg(challenges + (size_t)t * N + i);
                        const uint64_t* pl = t == 0 ? P0 : ((t & 1) ? buf_a : buf_b);
                        ar.to_words(ar.fold(ld_el_cg(ar, pl, 0), ld_el_cg(ar, pl, 1), ar.from_words(rw)), o);
#pragma unroll
                        for (int i = 0; i < N; ++i) out_final[i] = o[i];
                    }
                    __threadfence_system();
                }
            }
            if (solo) __syncthreads();  // this CTA's stores are ordered before its next loads
        }
        have_release = solo;
    }
}

}  // namespace scb
