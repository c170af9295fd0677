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

namespace scb {

struct EqPair {  // eq(point; idx) = lo[idx & (2^lb - 1)] * hi[idx >> lb]
    const uint64_t* lo;
    const uint64_t* hi;
    uint32_t lb;
};
template <class A>
__device__ __forceinline__ typename A::El eq_at(const A& ar, const EqPair& e, uint64_t idx) {
    return ar.mul(ld_el(ar, e.lo, idx & ((1ull << e.lb) - 1)), ld_el(ar, e.hi, idx >> e.lb));
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

// phase-1 tables h1, h2 over b (one thread per b, its gates through the CSR by in0)
template <class A>
__global__ void __launch_bounds__(kThreads) k_gkr_phase1(FieldDesc f, EqPair eq_r, const uint32_t* __restrict__ off0, const uint32_t* __restrict__ idx0,
                                                         const uint8_t* __restrict__ types, const uint32_t* __restrict__ in1,
                                                         const uint64_t* __restrict__ w, uint64_t* __restrict__ h1, uint64_t* __restrict__ h2,
                                                         uint64_t n_b) {
    const A ar(f);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_b; b += stride) {
        typename A::El s1 = ar.zero(), s2 = ar.zero();
        for (uint32_t t = off0[b]; t < off0[b + 1]; ++t) {
            const uint32_t a = idx0[t];
            const typename A::El e = eq_at(ar, eq_r, a);
            const typename A::El ew = ar.mul(e, ld_el(ar, w, in1[a]));
            if (types[a]) {
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
                                                         const uint32_t* __restrict__ idx1, const uint8_t* __restrict__ types,
                                                         const uint32_t* __restrict__ in0, ElemArg wu_arg, uint64_t* __restrict__ q,
                                                         uint64_t* __restrict__ s, uint64_t n_c) {
    const A ar(f);
    const typename A::El wu = ar.from_words(wu_arg.w);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_c; c += stride) {
        typename A::El sa = ar.zero(), sm = ar.zero();
        for (uint32_t t = off1[c]; t < off1[c + 1]; ++t) {
            const uint32_t a = idx1[t];
            const typename A::El e = ar.mul(eq_at(ar, eq_r, a), eq_at(ar, eq_u, in0[a]));
            if (types[a]) sm = ar.add(sm, e);
            else sa = ar.add(sa, e);
        }
        uint64_t o[A::N];
        ar.to_words(ar.add(sa, ar.mul(wu, sm)), o);
        st_words<A::N>(q + c * A::N, o);
        ar.to_words(ar.mul(wu, sa), o);
        st_words<A::N>(s + c * A::N, o);
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

}  // namespace scb
