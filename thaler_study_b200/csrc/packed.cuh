// packed.cuh -- 32-bit intermediate tables for the small-prime policy (p < 2^28).
//
// The prover's folded tables are internal: nothing but the next round ever reads them.  For the reference's own
// fields every canonical Montgomery value fits 32 bits, so the fused fold+message kernel can store its output as
// uint32 (and read it back as uint32 next round).  Inputs and anything handed to a caller stay in ark's 8-byte
// format.  Traffic per table over a whole proof: 8N (round 0) + 8N + 2N (round 1) + 3N(1 + 1/2 + ...) = 24N bytes
// instead of 32N -- the same field elements, a quarter less HBM traffic.
#pragma once
#include <cstdint>

#include "kernels.cuh"

namespace scb {

// QP adjacent quads per thread-iteration (IN32: QP = 2 makes the load 256-bit and the store 128-bit).
template <int K, bool IN32, bool OUT32, int QP>
__global__ void __launch_bounds__(kThreads, (QP >= 4 ? 4 : (K <= 3 ? 8 : 6)))
    k_fold_round_sp(FieldDesc f, TabsIn<K> in, TabsOut<K> outp, ElemArg rarg, uint64_t n_groups, uint64_t* partials,
                    unsigned int* ticket, uint64_t* out, PeerArg peer) {
    using A = PolSP;
    constexpr int NP = K + 1;
    static_assert(IN32 || QP == 1, "QP > 1 only for packed input");
    const A ar(f);
    const A::FoldC r = ar.fold_const(ar.from_words(rarg.w));
    A::Acc acc[NP];
#pragma unroll
    for (int x = 0; x < NP; ++x) ar.acc_zero(acc[x]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        uint32_t t[K][4 * QP];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if constexpr (IN32) {
                uint64_t w[2 * QP];  // 4*QP packed entries
                ld_words<2 * QP>(in.p[k] + g * 2 * QP, w);
#pragma unroll
                for (int q = 0; q < 2 * QP; ++q) {
                    t[k][2 * q] = (uint32_t)w[q];
                    t[k][2 * q + 1] = (uint32_t)(w[q] >> 32);
                }
            } else {
                uint64_t w[4];
                ld_words<4>(in.p[k] + g * 4, w);
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
            if constexpr (OUT32) {
                uint64_t o[QP];
#pragma unroll
                for (int q = 0; q < QP; ++q) o[q] = (uint64_t)u[k][2 * q] | ((uint64_t)u[k][2 * q + 1] << 32);
                st_words<QP>(outp.p[k] + g * QP, o);
            } else {
                uint64_t o[2 * QP];
#pragma unroll
                for (int q = 0; q < 2 * QP; ++q) o[q] = u[k][q];
                st_words<2 * QP>(outp.p[k] + g * 2 * QP, o);
            }
        }
    }
    grid_reduce_finish<A, NP>(ar, acc, partials, ticket, out, K, &peer);
}

// packed uint32 table -> ark's 8-byte elements (only when a caller asks for an intermediate table)
__global__ void __launch_bounds__(kThreads) k_unpack32(const uint32_t* __restrict__ in, uint64_t* __restrict__ outp, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) outp[i] = in[i];
}

}  // namespace scb
