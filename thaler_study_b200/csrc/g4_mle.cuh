// g4_mle.cuh -- MLE evaluation of a 4-limb table in one launch with unreduced products.
//
// Replaces vsbw_multilinear_from_evaluations (multilinear-extensions/src/lib.rs:6-24: chi table by doubling, then the dot
// product with the evaluations) and [ARK] DenseMultilinearExtension::evaluate, like k_mle_eval_fused (eqfix.cuh), for
// fields of four limbs.  The structure is k_mle_eval_fused's -- every CTA builds the eq sub-tables in shared memory,
// eq(i) = lo[i mod 2^LB] * h(i >> LB) with h a product of up to three sub-table entries, a warp takes a row of 2^LB
// entries -- but the arithmetic follows g4.cuh's fourth generation: a dot product is a SUM of products, REDC is linear,
// so no product of the row is reduced: each lane adds the plain 512-bit products evals[j] * lo[j] of its 32 entries
// into one 544-bit register accumulator (64 wide multiply-adds per entry instead of 128), reduces it once per row,
// and adds row_sum * h -- again unreduced -- into a second 544-bit accumulator that is reduced once per thread.
// Per row and lane: 32 x 64 + 512 (row reduction) + 256 (h) + 64 wide multiply-adds, against 32 x 128 + 384 before.
#pragma once
#include <cstdint>

#include "eqfix.cuh"
#include "g4.cuh"

namespace scb {
namespace g4 {

struct Wide17 {  // 544-bit integer in registers
    uint32_t w[17];
};
__device__ __forceinline__ void wide_zero(Wide17& a) {
#pragma unroll
    for (int i = 0; i < 17; ++i) a.w[i] = 0;
}
__device__ __forceinline__ void wide_add(Wide17& a, const uint32_t (&t)[16]) {
    asm("add.cc.u32  %0, %0, %17;\n\t"
        "addc.cc.u32 %1, %1, %18;\n\t"
        "addc.cc.u32 %2, %2, %19;\n\t"
        "addc.cc.u32 %3, %3, %20;\n\t"
        "addc.cc.u32 %4, %4, %21;\n\t"
        "addc.cc.u32 %5, %5, %22;\n\t"
        "addc.cc.u32 %6, %6, %23;\n\t"
        "addc.cc.u32 %7, %7, %24;\n\t"
        "addc.cc.u32 %8, %8, %25;\n\t"
        "addc.cc.u32 %9, %9, %26;\n\t"
        "addc.cc.u32 %10, %10, %27;\n\t"
        "addc.cc.u32 %11, %11, %28;\n\t"
        "addc.cc.u32 %12, %12, %29;\n\t"
        "addc.cc.u32 %13, %13, %30;\n\t"
        "addc.cc.u32 %14, %14, %31;\n\t"
        "addc.cc.u32 %15, %15, %32;\n\t"
        "addc.u32    %16, %16, 0;\n\t"
        : "+r"(a.w[0]), "+r"(a.w[1]), "+r"(a.w[2]), "+r"(a.w[3]), "+r"(a.w[4]), "+r"(a.w[5]), "+r"(a.w[6]), "+r"(a.w[7]), "+r"(a.w[8]), "+r"(a.w[9]),
          "+r"(a.w[10]), "+r"(a.w[11]), "+r"(a.w[12]), "+r"(a.w[13]), "+r"(a.w[14]), "+r"(a.w[15]), "+r"(a.w[16])
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]),
          "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]));
}
// REDC of a 544-bit sum of products, canonical (the register form of wacc_reduce)
template <bool P0ONE>
__device__ __forceinline__ W8 wide_reduce(const ArithT<P0ONE>& ar, const W8& r2, const Wide17& a) {
    W8 c0, c1, c2, one_int;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        c0.w[q] = a.w[q];
        c1.w[q] = a.w[8 + q];
        c2.w[q] = one_int.w[q] = 0;
    }
    c2.w[0] = a.w[16];
    one_int.w[0] = 1;
    const W8 x0 = ar.mul(c0, one_int);
    const W8 x1 = ar.mul(ar.mul(c1, r2), one_int);
    const W8 x2 = ar.mul(c2, r2);
    return ar.add(ar.add(x0, x1), x2);
}

template <bool P0ONE>
__global__ void __launch_bounds__(kThreads, 2)
    k_mle_eval_fused_g4(FieldDesc f, PointArg pt, const uint64_t* __restrict__ evals, uint32_t v_local, uint32_t v_total, uint64_t row0,
                        uint64_t* partials, unsigned int* ticket, uint64_t* out, PeerArg peer) {
    using A = PolGN<4>;
    constexpr int N = 4, LB = MleFusedCfg<A>::LB, SUB = MleFusedCfg<A>::SUB;
    extern __shared__ uint64_t eq_sm[];
    const A pa(f);
    const ArithT<P0ONE> ar(f);
    const uint32_t hb = v_total - LB;
    const uint32_t nsub = hb == 0 ? 0 : (hb + SUB - 1) / SUB;  // <= 3
    const uint32_t q = nsub ? hb / nsub : 0, rem = nsub ? hb % nsub : 0;
    const uint32_t b1 = nsub > 0 ? q + (0 < rem ? 1 : 0) : 0, b2 = nsub > 1 ? q + (1 < rem ? 1 : 0) : 0, b3 = nsub > 2 ? q + (2 < rem ? 1 : 0) : 0;
    uint64_t* const t1 = eq_sm + ((size_t)N << LB);
    uint64_t* const t2 = t1 + ((size_t)N << SUB);
    uint64_t* const t3 = t2 + ((size_t)N << SUB);
    {  // thread group g (64 threads) doubles table g: g = 0 the low table, 1..3 the high sub-tables (as k_mle_eval_fused)
        const uint32_t g = threadIdx.x >> 6, tid = threadIdx.x & 63;
        uint64_t* const tg = g == 0 ? eq_sm : (g == 1 ? t1 : (g == 2 ? t2 : t3));
        const uint32_t gbits = g == 0 ? (uint32_t)LB : (g == 1 ? b1 : (g == 2 ? b2 : b3));
        const uint32_t gfirst = g == 0 ? 0u : (g == 1 ? (uint32_t)LB : (g == 2 ? LB + b1 : LB + b1 + b2));
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) tg[i] = f.one[i];
        }
        __syncthreads();
        for (uint32_t l = 0; l < (uint32_t)(LB > SUB ? LB : SUB); ++l) {
            if (l < gbits) {
                const W8 c = load8(pt.w + (size_t)(gfirst + l) * N);
                const uint32_t half = 1u << l;
                for (uint32_t i = tid; i < half; i += 64) {
                    const W8 cur = load8(tg + (size_t)i * N);
                    const W8 hi = ar.mul(cur, c);
                    store8(hi, tg + (size_t)(i + half) * N);
                    store8(ar.sub(cur, hi), tg + (size_t)i * N);
                }
            }
            __syncthreads();
        }
    }
    const W8 r2 = load8(f.r2);
    Wide17 total;
    wide_zero(total);
    const int lane = threadIdx.x & 31;
    const uint64_t n_rows = 1ull << (v_local - LB);
    const uint64_t n_warps = (uint64_t)gridDim.x * (kThreads / 32);
    const uint64_t m1 = (1ull << b1) - 1, m2 = (1ull << b2) - 1;
    for (uint64_t row = (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); row < n_rows; row += n_warps) {
        const uint64_t* src = evals + ((row << LB) * N);
        Wide17 rsum;
        wide_zero(rsum);
        constexpr int JL = (1 << LB) / 32;  // one element (256-bit load) per lane per step
#pragma unroll 1
        for (int j0 = 0; j0 < JL; j0 += 4) {
            uint64_t w[4][N];
#pragma unroll
            for (int j = 0; j < 4; ++j) ld_words<N>(src + (size_t)((j0 + j) * 32 + lane) * N, w[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t t[16];
                ar.mul_wide(t, load8(w[j]), load8(eq_sm + (size_t)((j0 + j) * 32 + lane) * N));
                wide_add(rsum, t);  // 32 products below p^2 < 2^510: the sum stays below 2^515
            }
        }
        W8 s = wide_reduce<P0ONE>(ar, r2, rsum);
        uint32_t t[16];
        if (nsub > 0) {
            const uint64_t ih = row0 + row;
            W8 h = load8(t1 + (size_t)(ih & m1) * N);
            if (nsub > 1) h = ar.mul(h, load8(t2 + (size_t)((ih >> b1) & m2) * N));
            if (nsub > 2) h = ar.mul(h, load8(t3 + (size_t)(ih >> (b1 + b2)) * N));
            ar.mul_wide(t, s, h);
        } else {  // no high bits: the row sum itself, as the product with Montgomery one
            ar.mul_wide(t, s, load8(f.one));
        }
        wide_add(total, t);  // at most 2^24 rows per thread would still fit; a thread sees a handful
    }
    typename A::Acc fin[1];
    {
        const W8 sres = wide_reduce<P0ONE>(ar, r2, total);
        uint64_t l[4];
        store8(sres, l);
        fin[0] = pa.from_words(l);
    }
    grid_reduce_finish<A, 1>(pa, fin, partials, ticket, out, 0, &peer);
}

}  // namespace g4
}  // namespace scb
