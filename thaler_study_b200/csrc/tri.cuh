// tri.cuh -- triangle_counting::G (triangle-counting/src/lib.rs:70-166) in matrix form (hard part 6).
//
// g(x, y, z) = f1(x, y) f2(y, z) f3(x, z).  While an x variable is being eliminated (the first xn rounds and
// Prover::new's sum) every value the protocol needs has the shape  sum_{x, z} f3_X(x, z) * sum_y f1_X(x, y) f2(y, z).
// The inner sum is linear in f1, and f1 only ever changes by FOLDING an x variable (triangle-counting/src/lib.rs:89-118),
// which is linear too -- so the matrix
//     M[z][x] = sum_y f2[z][y] * f1[y][x]            (tables at index (row << bits) | column, :150-157,170-172)
// is computed ONCE per proof (one field matmul, n^3 products) and then folds along x exactly like a table:
// M_{j+1}[z][x'] = M_j[z][2x'] + r (M_j[z][2x'+1] - M_j[z][2x']).  The x phase is then the sum-check of the PRODUCT of
// the two tables M and f3 over the index (z << xn) | x -- adjacent pairs, streamed by the product kernels of kernels.cuh
// -- instead of n^3 products per message point per round (round 1's k_triangle_round: 3 n^3 / 2 products in round 0
// alone).  Exact field arithmetic: the same sums, the same messages.
//
// k_field_matmul: shared-memory tiled, register-blocked.  Small-prime policy (p < 2^28, all of the reference's moduli):
// products of 32-bit residues are accumulated unreduced in 64 bits -- ONE IMAD.WIDE per multiply-add, no reduction in
// the inner loop; the accumulator is folded with 2^32 mod p every 128 terms and reduced once per output (redc:
// sum aR bR 2^-64 = (sum ab) R, the Montgomery form ark stores).  Other policies: Montgomery product + modular add.
#pragma once
#include <cstdint>

#include "kernels.cuh"

namespace scb {

constexpr int kMmTile = 64;   // outputs per CTA: kMmTile (z) x kMmTile (x)
constexpr int kMmK = 32;      // y values per shared-memory step

// M[z][x] = sum_y B[z][y] * A[y][x];  A: Y x X (f1), B: Z x Y (f2), M: Z x X.  Dimensions are powers of two.
// grid = (ceil(X/64), ceil(Z/64)); 256 threads, each a 4 x 4 block of outputs: four adjacent x (one 128-bit
// shared-memory read per step) by four z strided by 16 (broadcast reads).  Tiles smaller than 64 (small tables) are handled by bounds checks.
__global__ void __launch_bounds__(256) k_field_matmul_sp(FieldDesc f, const uint64_t* __restrict__ A, const uint64_t* __restrict__ B,
                                                        uint64_t* __restrict__ M, uint32_t X, uint32_t Y, uint32_t Z) {
    __shared__ __align__(16) uint32_t sa[kMmK][kMmTile + 4];   // [y][x]; rows are 272 B apart: 16-byte aligned
    __shared__ uint32_t sb[kMmTile][kMmK + 1];   // [z][y]
    const PolSP ar(f);
    const uint32_t c32 = (uint32_t)((1ull << 32) % ar.p);
    const uint32_t x0 = blockIdx.x * kMmTile, z0 = blockIdx.y * kMmTile;
    const uint32_t tx = threadIdx.x & 15, tz = threadIdx.x >> 4;  // outputs x0 + 4 tx + i, z0 + tz + 16 j
    uint64_t acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0;
    uint32_t since_fold = 0;
    for (uint32_t y0 = 0; y0 < Y; y0 += kMmK) {
        // stage A[y0 .. y0+32)[x0 .. x0+64) and B[z0 .. z0+64)[y0 .. y0+32): 2048 entries each, 8 per thread, the low
        // 32 bits of every 8-byte element (values below 2^28)
        for (uint32_t e = threadIdx.x; e < kMmK * kMmTile; e += 256) {
            const uint32_t yy = e / kMmTile, xx = e % kMmTile;
            sa[yy][xx] = (y0 + yy < Y && x0 + xx < X) ? (uint32_t)A[(size_t)(y0 + yy) * X + x0 + xx] : 0u;
            const uint32_t zz = e / kMmK, y2 = e % kMmK;
            sb[zz][y2] = (z0 + zz < Z && y0 + y2 < Y) ? (uint32_t)B[(size_t)(z0 + zz) * Y + y0 + y2] : 0u;
        }
        __syncthreads();
#pragma unroll 8
        for (uint32_t k = 0; k < kMmK; ++k) {
            uint32_t a[4], b[4];
            const uint4 av = *reinterpret_cast<const uint4*>(&sa[k][4 * tx]);  // one 128-bit read: 16 lanes x 16 B contiguous
            a[0] = av.x, a[1] = av.y, a[2] = av.z, a[3] = av.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sb[tz + 16 * j][k];  // two distinct words per warp: broadcasts
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] += (uint64_t)a[i] * b[j];  // mad.wide.u32 with a 64-bit addend
        }
        __syncthreads();
        since_fold += kMmK;
        if (since_fold >= 128) {  // products < 2^56: 128 of them on top of a folded value (< 2^61) stay below 2^64
            since_fold = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] = (acc[j][i] & 0xffffffffull) + (acc[j][i] >> 32) * c32;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t x = x0 + 4 * tx + i, z = z0 + tz + 16 * j;
            if (x < X && z < Z) {
                const uint64_t t = (acc[j][i] & 0xffffffffull) + (acc[j][i] >> 32) * c32;  // < 2^61
                M[(size_t)z * X + x] = ar.reduce_once(ar.redc(t));
            }
        }
}

// Generic policies: 16 x 16 outputs per CTA, one per thread, operands staged in shared memory as full elements.
template <class A>
__global__ void __launch_bounds__(256) k_field_matmul_gen(FieldDesc f, const uint64_t* __restrict__ Am, const uint64_t* __restrict__ Bm,
                                                         uint64_t* __restrict__ M, uint32_t X, uint32_t Y, uint32_t Z) {
    constexpr int N = A::N, T = 16;
    __shared__ uint64_t sa[T][T][N];  // [y][x]
    __shared__ uint64_t sb[T][T][N];  // [z][y]
    const A ar(f);
    const uint32_t tx = threadIdx.x & 15, tz = threadIdx.x >> 4;
    const uint32_t x = blockIdx.x * T + tx, z = blockIdx.y * T + tz;
    typename A::Acc acc;
    ar.acc_zero(acc);
    for (uint32_t y0 = 0; y0 < Y; y0 += T) {
        {
            const uint32_t yy = tz, xx = blockIdx.x * T + tx;  // A[y0 + tz][x]
#pragma unroll
            for (int q = 0; q < N; ++q) sa[tz][tx][q] = (y0 + yy < Y && xx < X) ? Am[((size_t)(y0 + yy) * X + xx) * N + q] : 0ull;
            const uint32_t zz = blockIdx.y * T + tz, y2 = y0 + tx;  // B[z][y0 + tx]
#pragma unroll
            for (int q = 0; q < N; ++q) sb[tz][tx][q] = (zz < Z && y2 < Y) ? Bm[((size_t)zz * Y + y2) * N + q] : 0ull;
        }
        __syncthreads();
#pragma unroll 1
        for (uint32_t k = 0; k < T; ++k) ar.acc_add(acc, ar.lz_mul(ar.lz(ar.from_words(sa[k][tx])), ar.lz(ar.from_words(sb[tz][k]))));
        __syncthreads();
    }
    if (x < X && z < Z) {
        uint64_t o[N];
        ar.to_words(ar.acc_final(acc), o);
#pragma unroll
        for (int q = 0; q < N; ++q) M[((size_t)z * X + x) * N + q] = o[q];
    }
}

}  // namespace scb
