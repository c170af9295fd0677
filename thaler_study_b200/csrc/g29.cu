// g29.cu -- translation unit of the radix-2^29 lazy-carry 4-limb kernels (g29.cuh, lazy29.hpp; option g4_kernel = 2, kept as
// a measured negative result): its own object file so that it compiles next to g4.cu and engine.cu.
#include <cuda_runtime.h>

#include "g29.cuh"
#include "g4_launch.hpp"

namespace scb {

template <int K, int MINB>
static cudaError_t launch_k29(int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const g29::Desc29x& dx, const uint64_t* const* in,
                              uint64_t* const* outp, const ElemArg& r5, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res,
                              const PeerArg& pa, int max_grid) {
    auto kern = g29::k_fold_round_g29<K, MINB>;
    static int nb_cached = 0;
    if (nb_cached == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, 0) != cudaSuccess || nb < 1) nb = 1;
        nb_cached = nb;
    }
    int nb = nb_cached;
    if (blocks_per_sm_cap > 0 && blocks_per_sm_cap < nb) nb = blocks_per_sm_cap;
    uint64_t want = (n_quads + kThreads - 1) / kThreads, cap = (uint64_t)sms * nb;
    if (cap > (uint64_t)max_grid) cap = max_grid;
    if (want < 1) want = 1;
    const int grid = (int)(want < cap ? want : cap);
    TabsIn<K> ti;
    TabsOut<K> to;
    for (int k = 0; k < K; ++k) {
        ti.p[k] = in[k];
        to.p[k] = outp[k];
    }
    kern<<<grid, kThreads, 0, stream>>>(f, dx, ti, to, r5, n_quads, partials, ticket, res, pa);
    return cudaGetLastError();
}

// 2^e mod p by e doublings (four 64-bit words), as 29-bit limbs
static void pow2_mod_limbs(const FieldDesc& f, int e, uint32_t (&out)[9]) {
    uint64_t x[4] = {1, 0, 0, 0};
    auto geq = [&](const uint64_t* a, uint64_t top) {
        if (top) return true;
        for (int i = 3; i >= 0; --i) {
            if (a[i] > f.p[i]) return true;
            if (a[i] < f.p[i]) return false;
        }
        return true;
    };
    for (int s = 0; s < e; ++s) {
        uint64_t top = x[3] >> 63;
        for (int i = 3; i > 0; --i) x[i] = (x[i] << 1) | (x[i - 1] >> 63);
        x[0] <<= 1;
        if (geq(x, top)) {
            uint64_t borrow = 0;
            for (int i = 0; i < 4; ++i) {
                const uint64_t d = x[i] - f.p[i], b1 = x[i] < f.p[i], d2 = d - borrow, b2 = d < borrow;
                x[i] = d2;
                borrow = b1 | b2;
            }
        }
    }
    uint32_t w[8];
    for (int i = 0; i < 4; ++i) {
        w[2 * i] = (uint32_t)x[i];
        w[2 * i + 1] = (uint32_t)(x[i] >> 32);
    }
    const l29::L9 r = l29::from_words(w);
    for (int j = 0; j < 9; ++j) out[j] = r.l[j];
}

template <int K>
static cudaError_t launch_r29(int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const g29::Desc29x& dx, const uint64_t* const* in,
                              uint64_t n_pairs, uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid) {
    auto kern = g29::k_round_evals_g29<K>;
    static int nb_cached = 0;
    if (nb_cached == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, 0) != cudaSuccess || nb < 1) nb = 1;
        nb_cached = nb;
    }
    int nb = nb_cached;
    if (blocks_per_sm_cap > 0 && blocks_per_sm_cap < nb) nb = blocks_per_sm_cap;
    uint64_t want = (n_pairs + kThreads - 1) / kThreads, cap = (uint64_t)sms * nb;
    if (cap > (uint64_t)max_grid) cap = max_grid;
    if (want < 1) want = 1;
    TabsIn<K> ti;
    for (int k = 0; k < K; ++k) ti.p[k] = in[k];
    kern<<<(int)(want < cap ? want : cap), kThreads, 0, stream>>>(f, dx, ti, n_pairs, partials, ticket, res, pa);
    return cudaGetLastError();
}
cudaError_t launch_round_evals_g29(int K, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in, uint64_t n_pairs,
                                   uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid) {
    g29::Desc29x dx;
    if (!l29::make_desc(f.p, f.bits, &dx.d)) return cudaErrorInvalidValue;
    pow2_mod_limbs(f, 261, dx.c1);
    pow2_mod_limbs(f, 522, dx.c2);
    switch (K) {
        case 1: return launch_r29<1>(blocks_per_sm_cap, sms, stream, f, dx, in, n_pairs, partials, ticket, res, pa, max_grid);
        case 2: return launch_r29<2>(blocks_per_sm_cap, sms, stream, f, dx, in, n_pairs, partials, ticket, res, pa, max_grid);
        case 3: return launch_r29<3>(blocks_per_sm_cap, sms, stream, f, dx, in, n_pairs, partials, ticket, res, pa, max_grid);
        case 4: return launch_r29<4>(blocks_per_sm_cap, sms, stream, f, dx, in, n_pairs, partials, ticket, res, pa, max_grid);
        default: return cudaErrorInvalidValue;
    }
}

bool g29_supported(const FieldDesc& f) {
    l29::Desc29 d;
    return f.n == 4 && l29::make_desc(f.p, f.bits, &d);
}

cudaError_t launch_fold_round_g29(int K, int minb, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in,
                                  uint64_t* const* outp, const ElemArg& r5, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res,
                                  const PeerArg& pa, int max_grid) {
    g29::Desc29x dx;
    if (!l29::make_desc(f.p, f.bits, &dx.d)) return cudaErrorInvalidValue;
    pow2_mod_limbs(f, 261, dx.c1);
    pow2_mod_limbs(f, 522, dx.c2);
    switch (K) {
        case 1: return minb >= 3 ? launch_k29<1, 3>(blocks_per_sm_cap, sms, stream, f, dx, in, outp, r5, n_quads, partials, ticket, res, pa, max_grid) : launch_k29<1, 2>(blocks_per_sm_cap, sms, stream, f, dx, in, outp, r5, n_quads, partials, ticket, res, pa, max_grid);
        case 2: return minb >= 3 ? launch_k29<2, 3>(blocks_per_sm_cap, sms, stream, f, dx, in, outp, r5, n_quads, partials, ticket, res, pa, max_grid) : launch_k29<2, 2>(blocks_per_sm_cap, sms, stream, f, dx, in, outp, r5, n_quads, partials, ticket, res, pa, max_grid);
        case 3: return minb >= 3 ? launch_k29<3, 3>(blocks_per_sm_cap, sms, stream, f, dx, in, outp, r5, n_quads, partials, ticket, res, pa, max_grid) : launch_k29<3, 2>(blocks_per_sm_cap, sms, stream, f, dx, in, outp, r5, n_quads, partials, ticket, res, pa, max_grid);
        case 4: return minb >= 3 ? launch_k29<4, 3>(blocks_per_sm_cap, sms, stream, f, dx, in, outp, r5, n_quads, partials, ticket, res, pa, max_grid) : launch_k29<4, 2>(blocks_per_sm_cap, sms, stream, f, dx, in, outp, r5, n_quads, partials, ticket, res, pa, max_grid);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace scb
