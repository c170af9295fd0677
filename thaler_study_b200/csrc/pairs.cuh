// pairs.cuh -- TWO sum-check rounds per pass over the tables (small-prime policy, product polynomials).
//
// Round j's message is g_j(X) = sum_{x'} prod_k f_k(X, x'), round j+1's is g_{j+1}(Y) = sum_{x''} prod_k f_k(r_j, Y, x'').
// Both are slices of ONE bivariate polynomial of degree K in each variable,
//     H(a, b) = sum_{x''} prod_k f_k(a, b, x''),        g_j(X) = H(X, 0) + H(X, 1),     g_{j+1}(Y) = H(r_j, Y),
// so a single pass that accumulates H on the (K+1) x (K+1) grid a, b in {0..K} yields both messages: the host takes
// g_j from the grid, hashes it into r_j, interpolates the grid rows at r_j (exact field arithmetic: the same field
// elements the reference's second pass would produce), hashes g_{j+1} into r_{j+1} -- and only then does the device
// touch the tables again, folding BOTH variables in one pass while accumulating the next grid.  The tables are
// streamed once per two rounds: with packed uint32 intermediates a proof moves 8 + 9 + 1.25 (1 + 1/4 + ...) = 18.7
// bytes per table entry instead of 24 (one round per pass) or 32, and half the rounds need no device
// pass -- hence no barrier, launch or host<->device turn-around -- at all.
//
// Multiplication count per entry is unchanged (the grid has (K+1)^2 products per 4 folded entries, the two line
// messages it replaces have 2 (K+1) per 2 + 2 (K+1) per ... entries); for p < 2^28 the passes stay HBM-bound.
#pragma once
#include <cstdint>

#include "kernels.cuh"
#include "persist.cuh"

namespace scb {

constexpr int kMaxGridPts = 25;  // (K+1)^2 for K <= 4

// Accumulates prod_k v_k(a, b) into acc[a * NP + b] for one 2x2 block of every table: c[k][y1 + 2 y2] canonical.
//
// The LAST factor is multiplied in without a Montgomery step and straight into the 64-bit accumulator (one
// IMAD.WIDE with the accumulator as addend): per grid point K-2 reduced products + 1 fused multiply-accumulate
// instead of K-1 reduced products + a 64-bit add.  Terms are < 13 p^2 (K <= 3) / 20 p^2 (K = 4), so an accumulator
// takes GridConsts<K>::fold_every terms before grid_fold() has to bring it back under 2^60.1; the sums carry K-2 factors
// of 2^-32, i.e. GridConsts<K>::msg_k restores the Montgomery form in msg_final.
template <int K>
struct GridConsts {
    static constexpr int fold_every = K <= 3 ? 16 : 8;
    // msg_final(acc, k) applies k-1 single REDC steps: 2(K-1) are due, K-2 were done inline and 2 by grid_canon
    static constexpr int msg_k = K >= 2 ? K - 1 : 1;
};
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
template <int NG>
__device__ __forceinline__ void grid_fold(uint32_t c32, uint64_t (&acc)[NG]) {  // c32 = 2^32 mod p; value mod p unchanged
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = mad_wide((uint32_t)(acc[i] >> 32), c32, (uint64_t)(uint32_t)acc[i]);
}
// Before sums across threads: bring every accumulator down to <= p.  A 64-bit `%` costs ~50 instructions; one fold
// (< 2^60.1) and one two-step REDC (x 2^-64, accounted for in GridConsts::msg_k) cost 8.  K = 1 sums are small.
template <int K, int NG>
__device__ __forceinline__ void grid_canon(const PolSP& ar, uint32_t c32, uint64_t (&acc)[NG]) {
    if constexpr (K >= 2) {
        grid_fold(c32, acc);
#pragma unroll
        for (int i = 0; i < NG; ++i) acc[i] = ar.redc(acc[i]);
    }
}
template <int K>
__device__ __forceinline__ void grid_accumulate(const PolSP& ar, const uint32_t (&c)[K][4], uint64_t (&acc)[(K + 1) * (K + 1)]) {
    constexpr int NP = K + 1;
    uint32_t P[NP * NP];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        // along y1, exact: va[y2][a] = c[y2][0] + a (c[y2][1] - c[y2][0])
        uint32_t va0[NP], va1[NP];
        const uint32_t d0 = ar.sub(c[k][1], c[k][0]), d1 = ar.sub(c[k][3], c[k][2]);
        va0[0] = c[k][0];
        va1[0] = c[k][2];
        va0[1] = c[k][1];
        va1[1] = c[k][3];
#pragma unroll
        for (int a = 2; a < NP; ++a) {
            va0[a] = ar.add(va0[a - 1], d0);
            va1[a] = ar.add(va1[a - 1], d1);
        }
        // along y2, lazy: v(a, b) = va0 + b (va1 - va0) < (2 NP - 1) p < 2^31
#pragma unroll
        for (int a = 0; a < NP; ++a) {
            const uint32_t e = ar.lz_diff(va1[a], va0[a]);
            uint32_t v = va0[a];
#pragma unroll
            for (int b = 0; b < NP; ++b) {
                if (b == 1) v = va1[a];
                else if (b > 1) v = ar.lz_add(v, e);
                const int i = a * NP + b;
                if (K == 1) acc[i] += v;
                else if (k == K - 1) acc[i] = mad_wide(P[i], v, acc[i]);
                else P[i] = k == 0 ? v : ar.msg_mul(P[i], v);
            }
        }
    }
}

// Pass without a fold (the first pass of a proof, or the first after a consolidation): grid of the table as it is.
// One 2x2 block = 4 adjacent entries per table per thread-iteration.
template <int K, bool IN32, bool NC>
__device__ __forceinline__ void grid_pass_sp(const PolSP& ar, const uint64_t* const (&src)[K], uint64_t n_groups, uint64_t start, uint64_t stride,
                                             uint64_t (&acc)[(K + 1) * (K + 1)]) {
    const uint32_t c32 = (uint32_t)((1ull << 32) % ar.p);
    uint32_t it = 0;
    for (uint64_t g = start; g < n_groups; g += stride) {
        if ((++it % GridConsts<K>::fold_every) == 0) grid_fold(c32, acc);
        uint32_t c[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if constexpr (IN32) {
                uint64_t w[2];
                ld_words_sel<2, NC>(src[k] + g * 2, w);
                c[k][0] = (uint32_t)w[0];
                c[k][1] = (uint32_t)(w[0] >> 32);
                c[k][2] = (uint32_t)w[1];
                c[k][3] = (uint32_t)(w[1] >> 32);
            } else {
                uint64_t w[4];
                ld_words_sel<4, NC>(src[k] + g * 4, w);
#pragma unroll
                for (int q = 0; q < 4; ++q) c[k][q] = (uint32_t)w[q];
            }
        }
        grid_accumulate<K>(ar, c, acc);
    }
    grid_canon<K>(ar, c32, acc);
}

// Fold two variables (challenges ra, rb as fold constants) and accumulate the grid of the folded table: 16 adjacent
// entries per table per thread-iteration -> one 2x2 block of the folded table, stored as 4 packed uint32.
template <int K, bool IN32, bool NC, bool GRID>
__device__ __forceinline__ void pair_pass_sp(const PolSP& ar, const PolSP::FoldC ra, const PolSP::FoldC rb, const uint64_t* const (&src)[K],
                                             uint64_t* const (&dst)[K], uint64_t n_groups, uint64_t start, uint64_t stride,
                                             uint64_t (&acc)[(K + 1) * (K + 1)]) {
    const uint32_t c32 = (uint32_t)((1ull << 32) % ar.p);
    uint32_t it = 0;
    for (uint64_t g = start; g < n_groups; g += stride) {
        if (GRID && (++it % GridConsts<K>::fold_every) == 0) grid_fold(c32, acc);
        uint32_t c[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint32_t t[16];
            if constexpr (IN32) {
                uint64_t w[8];
                ld_words_sel<8, NC>(src[k] + g * 8, w);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    t[2 * q] = (uint32_t)w[q];
                    t[2 * q + 1] = (uint32_t)(w[q] >> 32);
                }
            } else {
                uint64_t w[16];
                ld_words_sel<16, NC>(src[k] + g * 16, w);
#pragma unroll
                for (int q = 0; q < 16; ++q) t[q] = (uint32_t)w[q];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t lo = ar.fold_c(t[4 * q], t[4 * q + 1], ra), hi = ar.fold_c(t[4 * q + 2], t[4 * q + 3], ra);
                c[k][q] = ar.fold_c(lo, hi, rb);
            }
            uint64_t o[2] = {(uint64_t)c[k][0] | ((uint64_t)c[k][1] << 32), (uint64_t)c[k][2] | ((uint64_t)c[k][3] << 32)};
            st_words<2>(dst[k] + g * 2, o);
        }
        if constexpr (GRID) grid_accumulate<K>(ar, c, acc);
    }
    grid_canon<K>(ar, c32, acc);
}

// Pair pass over 8-byte tables with the loads software-pipelined across tables: while table k's 16 entries are folded,
// the 128 bytes of the next table (or of table 0 of the next thread-iteration) are already in flight.  Two raw
// buffers alternate; PAR says which one holds table 0 (it flips every iteration when K is odd), so all register-array
// indices are compile-time.
template <int K, int k, int CUR, bool IN32, bool NC>
__device__ __forceinline__ void pair_pipe_table(const PolSP& ar, const PolSP::FoldC ra, const PolSP::FoldC rb, const uint64_t* const (&src)[K],
                                                uint64_t* const (&dst)[K], uint64_t g, uint64_t g_next, bool has_next,
                                                uint64_t (&wa)[IN32 ? 8 : 16], uint64_t (&wb)[IN32 ? 8 : 16], uint32_t (&c)[K][4]) {
    constexpr int WPT = IN32 ? 8 : 16;
    uint64_t(&cur)[WPT] = CUR ? wb : wa;
    uint64_t(&nxt)[WPT] = CUR ? wa : wb;
    if constexpr (k + 1 < K) ld_words_sel<WPT, NC>(src[k + 1 < K ? k + 1 : 0] + g * WPT, nxt);
    else if (has_next) ld_words_sel<WPT, NC>(src[0] + g_next * WPT, nxt);
    uint32_t t[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        if constexpr (IN32) t[q] = (q & 1) ? (uint32_t)(cur[q / 2] >> 32) : (uint32_t)cur[q / 2];
        else t[q] = (uint32_t)cur[q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t lo = ar.fold_c(t[4 * q], t[4 * q + 1], ra), hi = ar.fold_c(t[4 * q + 2], t[4 * q + 3], ra);
        c[k][q] = ar.fold_c(lo, hi, rb);
    }
    uint64_t o[2] = {(uint64_t)c[k][0] | ((uint64_t)c[k][1] << 32), (uint64_t)c[k][2] | ((uint64_t)c[k][3] << 32)};
    st_words<2>(dst[k] + g * 2, o);
    if constexpr (k + 1 < K) pair_pipe_table<K, k + 1, CUR ^ 1, IN32, NC>(ar, ra, rb, src, dst, g, g_next, has_next, wa, wb, c);
}
// IN32: packed uint32 input (64 bytes per table and thread-iteration) instead of the caller's 8-byte entries (128)
template <int K, bool IN32, bool NC>
__device__ __forceinline__ void pair_pass_sp_pipe(const PolSP& ar, const PolSP::FoldC ra, const PolSP::FoldC rb, const uint64_t* const (&src)[K],
                                                  uint64_t* const (&dst)[K], uint64_t n_groups, uint64_t start, uint64_t stride,
                                                  uint64_t (&acc)[(K + 1) * (K + 1)]) {
    constexpr int WPT = IN32 ? 8 : 16;
    const uint32_t c32 = (uint32_t)((1ull << 32) % ar.p);
    uint32_t it = 0;
    uint64_t wa[WPT], wb[WPT];
    uint64_t g = start;
    if (g < n_groups) ld_words_sel<WPT, NC>(src[0] + g * WPT, wa);
    while (g < n_groups) {
        if ((++it % (GridConsts<K>::fold_every / 2)) == 0) grid_fold(c32, acc);
        uint64_t gn = g + stride;
        {
            uint32_t c[K][4];
            pair_pipe_table<K, 0, 0, IN32, NC>(ar, ra, rb, src, dst, g, gn, gn < n_groups, wa, wb, c);
            grid_accumulate<K>(ar, c, acc);
        }
        g = gn;
        if constexpr (K & 1) {  // table 0 of this iteration sits in the other buffer
            if (g >= n_groups) break;
            gn = g + stride;
            uint32_t c[K][4];
            pair_pipe_table<K, 0, 1, IN32, NC>(ar, ra, rb, src, dst, g, gn, gn < n_groups, wa, wb, c);
            grid_accumulate<K>(ar, c, acc);
            g = gn;
        }
    }
    grid_canon<K>(ar, c32, acc);
}

// ------------------------------------------------------------------------------------------ staged loads
// The grid passes spend ~300 issue cycles per thread-iteration at 2-3 CTAs per SM, so loads issued at the top of an
// iteration leave HBM idle while the warp computes (ncu: long_scoreboard dominates, issue active 54 %).  cp.async
// keeps the NEXT iteration's bytes in flight while the current one is being multiplied out, at no register cost:
// every thread owns WPT/2 16-byte slots per table and stage in shared memory (consecutive threads -> consecutive
// slots: conflict-free for the asynchronous writes and for the LDS.128 reads), so no CTA barrier is involved, only
// cp.async.wait_group.  `.cg` copies go through L2, which keeps them coherent with the buffers the resident kernel
// rewrites every other pass.
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int K, int WPT>  // WPT = u64 words per table per thread-iteration
struct Stager {
    // A warp's 32 consecutive groups are one contiguous tile of 32 * WPT words per table: chunk c of the tile is
    // copied by ONE warp-wide instruction (lane i moves bytes [512 c + 16 i, +16)), i.e. whole 128-byte lines per
    // request, and lane j then reads its own WPT words back from the tile.
    static constexpr int CH = WPT / 2;          // 512-byte chunks per warp tile
    static constexpr int TILE = 32 * WPT * 8;   // bytes per warp, table and stage
    uint32_t wbase;                             // shared-space address of this warp's stage-0 table-0 tile
    uint32_t lane;
    __device__ __forceinline__ Stager(void* smem)
        : wbase((uint32_t)__cvta_generic_to_shared(smem) + (threadIdx.x >> 5) * (2u * K * TILE)), lane(threadIdx.x & 31) {}
    static constexpr size_t bytes(int threads) { return (size_t)2 * K * TILE * (threads / 32); }
    // g0 = group of lane 0 (warp-uniform); chunks that lie entirely beyond the table are skipped
    __device__ __forceinline__ void issue(int s, const uint64_t* const (&src)[K], uint64_t g0, uint64_t n_groups) const {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const uint32_t off = c * 512u + lane * 16u;  // byte offset inside the tile
                if (g0 + off / (WPT * 8) < n_groups) cp_async16(wbase + (s * K + k) * TILE + off, src[k] + g0 * WPT + off / 8);
            }
        cp_async_commit();
    }
    __device__ __forceinline__ void read(int s, int k, uint64_t (&w)[WPT]) const {
#pragma unroll
        for (int c = 0; c < CH; ++c)
            asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];"
                         : "=l"(w[2 * c]), "=l"(w[2 * c + 1])
                         : "r"(wbase + (s * K + k) * TILE + lane * (WPT * 8) + c * 16)
                         : "memory");
    }
};

template <int K, bool IN32>
__device__ __forceinline__ void grid_pass_sp_staged(const PolSP& ar, void* smem, const uint64_t* const (&src)[K], uint64_t n_groups, uint64_t start,
                                                    uint64_t stride, uint64_t (&acc)[(K + 1) * (K + 1)]) {
    constexpr int WPT = IN32 ? 2 : 4;
    const Stager<K, WPT> st(smem);
    int s = 0;
    const uint32_t c32 = (uint32_t)((1ull << 32) % ar.p);
    uint32_t it = 0;
    uint64_t g0 = start - st.lane;  // warp-uniform loop: start and stride are multiples of 32 plus the lane
    if (g0 < n_groups) st.issue(0, src, g0, n_groups);
    while (g0 < n_groups) {
        if ((++it % GridConsts<K>::fold_every) == 0) grid_fold(c32, acc);
        const uint64_t gn = g0 + stride;
        __syncwarp();  // every lane is done reading the stage that is refilled next
        if (gn < n_groups) st.issue(s ^ 1, src, gn, n_groups);
        else cp_async_commit();  // empty group: keeps "all but the newest group" == "this iteration's data"
        cp_async_wait<1>();
        __syncwarp();  // ... and every lane's share of this iteration's tile has landed
        if (g0 + st.lane < n_groups) {
            uint32_t c[K][4];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint64_t w[WPT];
                st.read(s, k, w);
                if constexpr (IN32) {
                    c[k][0] = (uint32_t)w[0];
                    c[k][1] = (uint32_t)(w[0] >> 32);
                    c[k][2] = (uint32_t)w[1];
                    c[k][3] = (uint32_t)(w[1] >> 32);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) c[k][q] = (uint32_t)w[q];
                }
            }
            grid_accumulate<K>(ar, c, acc);
        }
        g0 = gn;
        s ^= 1;
    }
    cp_async_wait<0>();
    grid_canon<K>(ar, c32, acc);
}

// packed uint32 input: 16 entries = 64 bytes per table per thread-iteration
template <int K>
__device__ __forceinline__ void pair_pass_sp_staged(const PolSP& ar, void* smem, const PolSP::FoldC ra, const PolSP::FoldC rb,
                                                    const uint64_t* const (&src)[K], uint64_t* const (&dst)[K], uint64_t n_groups, uint64_t start,
                                                    uint64_t stride, uint64_t (&acc)[(K + 1) * (K + 1)]) {
    const Stager<K, 8> st(smem);
    int s = 0;
    const uint32_t c32 = (uint32_t)((1ull << 32) % ar.p);
    uint32_t it = 0;
    uint64_t g0 = start - st.lane;
    if (g0 < n_groups) st.issue(0, src, g0, n_groups);
    while (g0 < n_groups) {
        if ((++it % GridConsts<K>::fold_every) == 0) grid_fold(c32, acc);
        const uint64_t gn = g0 + stride;
        __syncwarp();
        if (gn < n_groups) st.issue(s ^ 1, src, gn, n_groups);
        else cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const uint64_t g = g0 + st.lane;
        if (g < n_groups) {
            uint32_t c[K][4];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint64_t w[8];
                st.read(s, k, w);
                uint32_t t[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    t[2 * q] = (uint32_t)w[q];
                    t[2 * q + 1] = (uint32_t)(w[q] >> 32);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t lo = ar.fold_c(t[4 * q], t[4 * q + 1], ra), hi = ar.fold_c(t[4 * q + 2], t[4 * q + 3], ra);
                    c[k][q] = ar.fold_c(lo, hi, rb);
                }
                uint64_t o[2] = {(uint64_t)c[k][0] | ((uint64_t)c[k][1] << 32), (uint64_t)c[k][2] | ((uint64_t)c[k][3] << 32)};
                st_words<2>(dst[k] + g * 2, o);
            }
            grid_accumulate<K>(ar, c, acc);
        }
        g0 = gn;
        s ^= 1;
    }
    cp_async_wait<0>();
    grid_canon<K>(ar, c32, acc);
}
// dynamic shared memory the staged passes of a K-table kernel need (threads = kThreads)
template <int K>
constexpr size_t pair_stage_bytes() {
    return Stager<K, 8>::bytes(kThreads);  // the packed pair pass is the largest: 2 stages x K x 64 B x threads
}

// ------------------------------------------------------------------------------------------ TMA-staged tiles
// Producer/consumer ring in shared memory: one elected thread of a dedicated producer warp streams whole CTA tiles
// (kThreads groups per table, contiguous in HBM) with cp.async.bulk (TMA 1-D bulk copies) that complete on an
// mbarrier; the compute warps wait on that "full" barrier, read their own groups back with LDS.128 and release the
// stage through an "empty" barrier.  Bytes in flight are set by the ring depth, not by registers or occupancy.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

constexpr int kTmaStages = 3;
template <int K, int WPT>
constexpr size_t tma_ring_bytes() {
    return (size_t)kTmaStages * K * kThreads * WPT * 8;
}

// K6a'  grid of a table as it is, TMA-staged.  blockDim = kThreads compute threads + one producer warp.
template <int K, bool IN32>
__global__ void __launch_bounds__(kThreads + 32, 3)
    k_grid_sp_tma(FieldDesc f, TabsIn<K> in, uint64_t n_groups, uint64_t* partials, unsigned int* ticket, uint64_t* out, PeerArg peer) {
    constexpr int NG = (K + 1) * (K + 1), WPT = IN32 ? 2 : 4, S = kTmaStages;
    constexpr uint32_t TILE_B = kThreads * WPT * 8;  // bytes per table and stage
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ __align__(8) uint64_t full_bar[S], empty_bar[S];
    const PolSP ar(f);
    uint64_t acc[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = 0;
    const uint64_t n_tiles = (n_groups + kThreads - 1) / kThreads;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), kThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == kThreads / 32) {
        if (lane == 0) {
            uint32_t it = 0;
            for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t s = it % S, ph = (it / S) & 1;
                if (it >= (uint32_t)S) mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                const uint64_t left = n_groups - tile * kThreads;
                const uint32_t bytes = (uint32_t)(left < kThreads ? left : kThreads) * WPT * 8;
                mbar_expect_tx(smem_u32(&full_bar[s]), K * bytes);
#pragma unroll
                for (int k = 0; k < K; ++k)
                    bulk_g2s(smem_u32(ring + (size_t)(s * K + k) * TILE_B), in.p[k] + tile * kThreads * WPT, bytes, smem_u32(&full_bar[s]));
            }
        }
    } else {
        const uint32_t c32 = (uint32_t)((1ull << 32) % ar.p);
        uint32_t it = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t s = it % S, ph = (it / S) & 1;
            if (((it + 1) % GridConsts<K>::fold_every) == 0) grid_fold(c32, acc);
            mbar_wait(smem_u32(&full_bar[s]), ph);
            if (tile * kThreads + threadIdx.x < n_groups) {
                uint32_t c[K][4];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    uint64_t w[WPT];
                    const uint32_t a = smem_u32(ring + (size_t)(s * K + k) * TILE_B) + threadIdx.x * (WPT * 8);
#pragma unroll
                    for (int q = 0; q < WPT; q += 2)
                        asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(w[q]), "=l"(w[q + 1]) : "r"(a + q * 8) : "memory");
                    if constexpr (IN32) {
                        c[k][0] = (uint32_t)w[0];
                        c[k][1] = (uint32_t)(w[0] >> 32);
                        c[k][2] = (uint32_t)w[1];
                        c[k][3] = (uint32_t)(w[1] >> 32);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) c[k][q] = (uint32_t)w[q];
                    }
                }
                grid_accumulate<K>(ar, c, acc);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty_bar[s]));
        }
        grid_canon<K>(ar, c32, acc);
    }
    grid_reduce_finish<PolSP, NG>(ar, acc, partials, ticket, out, GridConsts<K>::msg_k, &peer);
}

// ------------------------------------------------------------------------------------------ stand-alone kernels
// K6a  grid of a table as it is (Prover::new for the small-prime policy: c_1, g_1 and g_2 from one pass).
template <int K, bool IN32, bool STAGED>
__global__ void __launch_bounds__(kThreads, 3)
    k_grid_sp(FieldDesc f, TabsIn<K> in, uint64_t n_groups, uint64_t* partials, unsigned int* ticket, uint64_t* out, PeerArg peer) {
    constexpr int NG = (K + 1) * (K + 1);
    const PolSP ar(f);
    uint64_t acc[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = 0;
    const uint64_t* src[K];
#pragma unroll
    for (int k = 0; k < K; ++k) src[k] = in.p[k];
    extern __shared__ uint4 stage_smem[];
    const uint64_t start = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    if constexpr (STAGED) grid_pass_sp_staged<K, IN32>(ar, stage_smem, src, n_groups, start, stride, acc);
    else grid_pass_sp<K, IN32, true>(ar, src, n_groups, start, stride, acc);
    grid_reduce_finish<PolSP, NG>(ar, acc, partials, ticket, out, GridConsts<K>::msg_k, &peer);
}
// K6a''  grid kernel with the NEXT thread-iteration's loads issued before the current one is multiplied out (register
// double buffer: 115 registers, 2 CTAs per SM).  The pass spends ~270 issue slots per thread-iteration, so loads
// issued at the top of an iteration leave HBM idle while the warp computes; with the next iteration's 96 bytes per
// thread already in flight the kernel runs at 6.33 TB/s instead of 6.05 (same-box A/B).  Default for 8-byte tables.
template <int K>
__global__ void __launch_bounds__(kThreads, 2)
    k_grid_sp_pf(FieldDesc f, TabsIn<K> in, uint64_t n_groups, uint64_t* partials, unsigned int* ticket, uint64_t* out, PeerArg peer) {
    constexpr int NG = (K + 1) * (K + 1);
    const PolSP ar(f);
    uint64_t acc[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = 0;
    const uint32_t c32 = (uint32_t)((1ull << 32) % ar.p);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t w[K][4];
    if (g < n_groups) {
#pragma unroll
        for (int k = 0; k < K; ++k) ld_words<4>(in.p[k] + g * 4, w[k]);
    }
    uint32_t it = 0;
    while (g < n_groups) {
        if ((++it % GridConsts<K>::fold_every) == 0) grid_fold(c32, acc);
        uint32_t c[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) c[k][q] = (uint32_t)w[k][q];
        const uint64_t gn = g + stride;
        if (gn < n_groups) {
#pragma unroll
            for (int k = 0; k < K; ++k) ld_words<4>(in.p[k] + gn * 4, w[k]);
        }
        grid_accumulate<K>(ar, c, acc);
        g = gn;
    }
    grid_canon<K>(ar, c32, acc);
    grid_reduce_finish<PolSP, NG>(ar, acc, partials, ticket, out, GridConsts<K>::msg_k, &peer);
}
// ------------------------------------------------------------------------------------------ 21-bit triples
// K = 3 tables over a field of at most 21 bits (the reference's F_1572869): entry i of all three tables fits ONE
// 64-bit word, A[i] | B[i] << 21 | C[i] << 42.  Prover::new's grid pass has every entry in registers anyway, so it
// writes that word (8 bytes per index instead of the 24 it read); the first pair pass -- the one that would read the
// caller's 8-byte tables a second time -- then streams 8 bytes per index instead of 24.  Per index a proof moves
// 24 + 8 (grid pass) + 8 + 3 (pair pass) + 3.75 (1 + 1/4 + ...) = 48 bytes instead of 24 + 24 + 3 + 3.75 + ... = 56,
// and the index structure (adjacent pairs, power-of-two groups) is untouched.  Same field elements: the packed word
// holds the canonical values the second read would have fetched.
constexpr uint32_t kW21Mask = 0x1fffffu;
__global__ void __launch_bounds__(kThreads, 2)
    k_grid_sp_pf_w21(FieldDesc f, TabsIn<3> in, uint64_t* __restrict__ w21, uint64_t n_groups, uint64_t* partials, unsigned int* ticket,
                     uint64_t* out, PeerArg peer) {
    constexpr int K = 3, NG = (K + 1) * (K + 1);
    const PolSP ar(f);
    uint64_t acc[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = 0;
    const uint32_t c32 = (uint32_t)((1ull << 32) % ar.p);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t w[K][4];
    if (g < n_groups) {
#pragma unroll
        for (int k = 0; k < K; ++k) ld_words<4>(in.p[k] + g * 4, w[k]);
    }
    uint32_t it = 0;
    while (g < n_groups) {
        if ((++it % GridConsts<K>::fold_every) == 0) grid_fold(c32, acc);
        uint32_t c[K][4];
        uint64_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int k = 0; k < K; ++k) c[k][q] = (uint32_t)w[k][q];
            o[q] = w[0][q] | (w[1][q] << 21) | (w[2][q] << 42);  // canonical entries are below p < 2^21
        }
        st_words<4>(w21 + g * 4, o);
        const uint64_t gn = g + stride;
        if (gn < n_groups) {
#pragma unroll
            for (int k = 0; k < K; ++k) ld_words<4>(in.p[k] + gn * 4, w[k]);
        }
        grid_accumulate<K>(ar, c, acc);
        g = gn;
    }
    grid_canon<K>(ar, c32, acc);
    grid_reduce_finish<PolSP, NG>(ar, acc, partials, ticket, out, GridConsts<K>::msg_k, &peer);
}
// Two variables folded at once: the nested folds lo = t0 + ra (t1 - t0), hi = t2 + ra (t3 - t2), c = lo + rb (hi - lo) are the
// bilinear form c = w00 t0 + w10 t1 + w01 t2 + w11 t3 with w00 = (1 - ra)(1 - rb), w10 = ra (1 - rb), w01 = (1 - ra) rb,
// w11 = ra rb -- the same field element, so the same canonical value.  With the weights as per-pass constants W = w 2^32
// (fold_const's form) the four products go unreduced into one 64-bit sum (< 4 p^2 < 2^58) and ONE Montgomery step brings it
// back: 4 wide multiply-adds + mul.lo + mad.wide + one conditional subtraction (sum / 2^32 + p < 1.25 p for p < 2^28)
// instead of 3 x (wide multiply + mul.lo + mad.wide + two adds + two conditional subtractions).
struct Fold2C {
    uint32_t w00, w10, w01, w11;
};
__device__ __forceinline__ Fold2C fold2_const(const PolSP& ar, const FieldDesc& f, uint32_t ra_m, uint32_t rb_m) {  // Montgomery challenges
    const uint32_t one_m = (uint32_t)f.one[0];
    const uint32_t na = ar.sub(one_m, ra_m), nb = ar.sub(one_m, rb_m);
    Fold2C c;
    c.w00 = ar.fold_const(ar.mul(na, nb));
    c.w10 = ar.fold_const(ar.mul(ra_m, nb));
    c.w01 = ar.fold_const(ar.mul(na, rb_m));
    c.w11 = ar.fold_const(ar.mul(ra_m, rb_m));
    return c;
}
__device__ __forceinline__ uint32_t fold2(const PolSP& ar, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, const Fold2C& c) {
    uint64_t s = (uint64_t)t0 * c.w00;
    s = mad_wide(t1, c.w10, s);
    s = mad_wide(t2, c.w01, s);
    s = mad_wide(t3, c.w11, s);
    return ar.reduce_once(ar.redc32(s));
}
// the pair pass over the triple words: 16 words in, 3 x 4 packed uint32 out per thread-iteration.
// PF: the next thread-iteration's 128 bytes are in flight while this one is folded (two CTAs per SM instead of three).
// F2: the two folds as one bilinear form (fold2) instead of three nested folds.
template <bool PF, bool F2>
__global__ void __launch_bounds__(kThreads, (PF ? 2 : 3))
    k_pair_pass_sp_w21(FieldDesc f, const uint64_t* __restrict__ w21, TabsOut<3> outp, ElemArg ra_arg, ElemArg rb_arg, uint64_t n_groups,
                       uint64_t* partials, unsigned int* ticket, uint64_t* out, PeerArg peer) {
    constexpr int K = 3, NG = (K + 1) * (K + 1);
    const PolSP ar(f);
    const PolSP::FoldC ra = ar.fold_const(ar.from_words(ra_arg.w)), rb = ar.fold_const(ar.from_words(rb_arg.w));
    const Fold2C w2 = fold2_const(ar, f, ar.from_words(ra_arg.w), ar.from_words(rb_arg.w));
    uint64_t acc[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = 0;
    const uint32_t c32 = (uint32_t)((1ull << 32) % ar.p);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t w[16], wn[PF ? 16 : 1];
    if (PF && g < n_groups) ld_words<16>(w21 + g * 16, w);
    uint32_t it = 0;
    while (g < n_groups) {
        if ((++it % GridConsts<K>::fold_every) == 0) grid_fold(c32, acc);
        const uint64_t gn = g + stride;
        if constexpr (PF) {
            if (gn < n_groups) ld_words<16>(w21 + gn * 16, wn);
        } else {
            ld_words<16>(w21 + g * 16, w);
        }
        uint32_t c[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint32_t t[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) t[q] = (uint32_t)(w[q] >> (21 * k)) & kW21Mask;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if constexpr (F2) {
                    c[k][q] = fold2(ar, t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3], w2);
                } else {
                    const uint32_t lo = ar.fold_c(t[4 * q], t[4 * q + 1], ra), hi = ar.fold_c(t[4 * q + 2], t[4 * q + 3], ra);
                    c[k][q] = ar.fold_c(lo, hi, rb);
                }
            }
            uint64_t o[2] = {(uint64_t)c[k][0] | ((uint64_t)c[k][1] << 32), (uint64_t)c[k][2] | ((uint64_t)c[k][3] << 32)};
            st_words<2>(outp.p[k] + g * 2, o);
        }
        grid_accumulate<K>(ar, c, acc);
        if constexpr (PF) {
#pragma unroll
            for (int q = 0; q < 16; ++q) w[q] = wn[q];
        }
        g = gn;
    }
    grid_canon<K>(ar, c32, acc);
    grid_reduce_finish<PolSP, NG>(ar, acc, partials, ticket, out, GridConsts<K>::msg_k, &peer);
}

// K6b  one pair pass as its own launch (what the resident kernel runs per pass; used for profiling and as the
// non-resident path).
template <int K, bool IN32, bool STAGED>
__global__ void __launch_bounds__(kThreads, 3)
    k_pair_pass_sp(FieldDesc f, TabsIn<K> in, TabsOut<K> outp, ElemArg ra_arg, ElemArg rb_arg, uint64_t n_groups, uint64_t* partials,
                   unsigned int* ticket, uint64_t* out, PeerArg peer) {
    constexpr int NG = (K + 1) * (K + 1);
    const PolSP ar(f);
    const PolSP::FoldC ra = ar.fold_const(ar.from_words(ra_arg.w)), rb = ar.fold_const(ar.from_words(rb_arg.w));
    uint64_t acc[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = 0;
    const uint64_t* src[K];
    uint64_t* dst[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        src[k] = in.p[k];
        dst[k] = outp.p[k];
    }
    extern __shared__ uint4 stage_smem[];
    const uint64_t start = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    if constexpr (IN32 && STAGED) pair_pass_sp_staged<K>(ar, stage_smem, ra, rb, src, dst, n_groups, start, stride, acc);
    else pair_pass_sp<K, IN32, true, true>(ar, ra, rb, src, dst, n_groups, start, stride, acc);
    grid_reduce_finish<PolSP, NG>(ar, acc, partials, ticket, out, GridConsts<K>::msg_k, &peer);
}

// ------------------------------------------------------------------------------------------ resident kernel
// All pair passes of a proof in one cooperative launch.  Pass t folds the two lowest variables of a table of m
// entries-bits (m >= 3) by the challenge pair number t and posts: the (K+1)^2 grid of the folded table when it still
// has >= 2 variables, its (K+1) line sums when it has exactly one.  n_passes = ceil((m - 2) / 2).  Barrier, mailbox
// and solo-CTA endgame as in k_persist_rounds.
// U64IN: the first pass may read the caller's 8-byte tables.  The packed-only instantiation (what a proof runs after
// its first pass was a launch of its own) has the registers to pipeline the packed loads across tables as well.
template <int K, bool U64IN>
__global__ void __launch_bounds__(kThreads, (K <= 2 ? 3 : 2))
    k_persist_pairs_sp(FieldDesc f, TabsIn<K> in0, TabsOut<K> buf_a, TabsOut<K> buf_b, ElemArg ra0, ElemArg rb0, uint32_t m, uint32_t n_passes,
                       int in0_w32, TailMailbox* mb, PersistCtl* ctl, uint64_t* partials, uint64_t timeout_ns, int use_stage, PeerArg peer) {
    using A = PolSP;
    constexpr int NP = K + 1, NG = NP * NP;
    bool src_w32 = in0_w32 != 0;
    const A ar(f);
    extern __shared__ uint4 stage_smem[];  // pair_stage_bytes<K>()
    __shared__ uint64_t sm[32 * NG];
    __shared__ uint64_t r_sm[2];
    __shared__ int flag_sm;  // 1: this CTA took the last ticket, 2: abort
    if (threadIdx.x == 0) {
        r_sm[0] = ra0.w[0];
        r_sm[1] = rb0.w[0];
        if (blockIdx.x == 0) st_sys(&mb->stamp[2 * kTailMaxRounds + 1], globaltimer_ns());  // kernel start
    }
    __syncthreads();
    bool have_r = true;
    for (uint32_t t = 0; t < n_passes; ++t) {
        // m >= 4: 2^(m-4) thread-iterations of 16 entries; m == 3: one thread folds 8 entries to a line
        const uint64_t n_groups = m >= 4 ? 1ull << (m - 4) : 1;
        uint64_t active = (n_groups + blockDim.x - 1) / blockDim.x;
        if (active > gridDim.x) active = gridDim.x;
        const bool solo = active == 1;
        if (solo && blockIdx.x != 0) return;
        if (!have_r) {  // barrier + challenge pair: released by the CTA that finished pass t-1
            if (threadIdx.x == 0) {
                uint64_t c[2];
                const uint64_t t0 = globaltimer_ns();
                int bad = 0;
                for (;;) {
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(c[0]), "=l"(c[1]) : "l"(ctl->challenge) : "memory");
                    if ((uint32_t)(c[0] >> 32) == t && (uint32_t)(c[1] >> 32) == t) break;
                    if ((uint32_t)(c[0] >> 32) == kMbAbortTag || globaltimer_ns() - t0 > 4 * timeout_ns) {
                        bad = 1;
                        break;
                    }
                }
                asm volatile("fence.acq_rel.gpu;" ::: "memory");  // pass t-1's folded entries before this pass's loads
                r_sm[0] = (uint32_t)c[0];
                r_sm[1] = (uint32_t)c[1];
                flag_sm = bad ? 2 : 0;
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
        const A::FoldC ra = ar.fold_const((uint32_t)r_sm[0]), rb = ar.fold_const((uint32_t)r_sm[1]);
        uint64_t acc[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) acc[i] = 0;
        const bool grid_out = m >= 4;  // else (m == 3) a line of NP sums in acc[0..NP)
        if (blockIdx.x < active) {
            const uint64_t start = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = active * blockDim.x;
            if (m >= 4) {
                if constexpr (U64IN) {
                    if (!src_w32 && (use_stage & 2)) pair_pass_sp_pipe<K, false, true>(ar, ra, rb, src, dst, n_groups, start, stride, acc);
                    else if (!src_w32) pair_pass_sp<K, false, true, true>(ar, ra, rb, src, dst, n_groups, start, stride, acc);  // only the caller's tables
                    else if (use_stage & 1) pair_pass_sp_staged<K>(ar, stage_smem, ra, rb, src, dst, n_groups, start, stride, acc);
                    else if (t == 0) pair_pass_sp<K, true, true, true>(ar, ra, rb, src, dst, n_groups, start, stride, acc);
                    else pair_pass_sp<K, true, false, true>(ar, ra, rb, src, dst, n_groups, start, stride, acc);
                } else {
                    pair_pass_sp_pipe<K, true, false>(ar, ra, rb, src, dst, n_groups, start, stride, acc);  // L2-coherent loads throughout
                }
            } else if (start == 0) {  // 8 entries per table -> 2 -> line sums
                uint32_t prod[NP];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    uint32_t e[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (src_w32) {
                            const uint64_t w = __ldcg(src[k] + q / 2);
                            e[q] = (q & 1) ? (uint32_t)(w >> 32) : (uint32_t)w;
                        } else {
                            e[q] = (uint32_t)__ldcg(src[k] + q);
                        }
                    }
                    const uint32_t u0 = ar.fold_c(ar.fold_c(e[0], e[1], ra), ar.fold_c(e[2], e[3], ra), rb);
                    const uint32_t u1 = ar.fold_c(ar.fold_c(e[4], e[5], ra), ar.fold_c(e[6], e[7], ra), rb);
                    __stcg(dst[k], (uint64_t)u0 | ((uint64_t)u1 << 32));  // the folded pair, for out_folded
                    pair_into_prod<A, NP>(ar, k == 0, u0, u1, prod);
                }
#pragma unroll
                for (int x = 0; x < NP; ++x) acc[x] += prod[x];
            }
            __threadfence();  // folded entries are visible device-wide before this CTA's ticket / next pass's loads
            block_reduce<A, NG>(ar, acc, sm);
            bool finisher = solo;
            if (!solo) {
                if (threadIdx.x == 0) {
#pragma unroll
                    for (int i = 0; i < NG; ++i) __stcg(&partials[(size_t)blockIdx.x * NG + i], acc[i]);
                    __threadfence();
                    const unsigned int tk = atomicAdd(&ctl->ticket[t], 1u);
                    flag_sm = (tk == (unsigned int)active - 1) ? 1 : 0;
                }
                __syncthreads();
                finisher = flag_sm == 1;
                if (finisher) {
                    __threadfence();
#pragma unroll
                    for (int i = 0; i < NG; ++i) acc[i] = 0;
                    for (unsigned int b = threadIdx.x; b < (unsigned int)active; b += blockDim.x) {
#pragma unroll
                        for (int i = 0; i < NG; ++i) acc[i] += __ldcg(&partials[(size_t)b * NG + i]);
                    }
                    block_reduce<A, NG>(ar, acc, sm);
                }
            }
            const uint64_t hi = (uint64_t)(t + 1) << 32;
            if (finisher) {  // CTA-uniform.  The (K+1)^2 final reductions (64-bit modulo + REDC steps) run one per thread
                const int n_out = grid_out ? NG : NP;
                if (threadIdx.x == 0) {
#pragma unroll
                    for (int i = 0; i < NG; ++i) sm[i] = acc[i];
                }
                __syncthreads();
                if (threadIdx.x < NG) {
                    const uint64_t v = ar.msg_final(sm[threadIdx.x], grid_out ? GridConsts<K>::msg_k : K);
                    if (peer.world > 1) sm[NG + threadIdx.x] = v;
                    else if ((int)threadIdx.x < n_out) st_sys(&mb->evals[threadIdx.x], hi | v);
                }
                if (peer.world > 1) {  // sharded prover: add the peer GPUs' sums of this pass (NVLink peer windows)
                    __syncthreads();
                    if (threadIdx.x < 32) {  // warp 0: stores to all peers in parallel, lane g polls rank g, lane i adds sum i
                        PeerArg pa = peer;
                        pa.seq += t;
                        peer_exchange_warp<A, NG>(ar, pa, sm + NG);
                        if ((int)threadIdx.x < n_out) st_sys(&mb->evals[threadIdx.x], hi | sm[NG + threadIdx.x]);
                    }
                    __syncthreads();
                }
            }
            if (finisher && threadIdx.x == 0) {
                st_sys(&mb->stamp[2 * t], globaltimer_ns());
                int bad = 0;
                if (t + 1 < n_passes) {
                    uint64_t c[2];
                    const uint64_t t0 = globaltimer_ns();
                    for (;;) {
                        asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(c[0]), "=l"(c[1]) : "l"(mb->challenge) : "memory");
                        if ((uint32_t)(c[0] >> 32) == t + 1 && (uint32_t)(c[1] >> 32) == t + 1) break;
                        if ((uint32_t)(c[0] >> 32) == kMbAbortTag || globaltimer_ns() - t0 > timeout_ns) {
                            bad = 1;
                            break;
                        }
                    }
                    if (bad) {
                        st_gpu(&ctl->challenge[0], (uint64_t)kMbAbortTag << 32);
                        st_sys(&mb->dev_status, 2);
                    } else {
                        r_sm[0] = (uint32_t)c[0];
                        r_sm[1] = (uint32_t)c[1];
                        if (!solo) {
                            st_gpu(&ctl->challenge[0], hi | (uint32_t)c[0]);
                            st_gpu(&ctl->challenge[1], hi | (uint32_t)c[1]);
                        }
                        st_sys(&mb->stamp[2 * t + 1], globaltimer_ns());
                    }
                } else {
                    st_sys(&mb->dev_status, 1);
                }
                if (solo) flag_sm = bad ? 2 : 0;
            }
            if (solo) {
                __syncthreads();
                if (flag_sm == 2) return;
            }
        }
        have_r = solo;
        m -= 2;
        src_w32 = true;
    }
}

}  // namespace scb
