// g4.cu -- translation unit of the 4-limb fused fold + message kernel (g4.cuh); its own object file so that the
// ~50 KB unrolled Montgomery bodies compile in parallel with engine.cu.
#include <cuda_runtime.h>

#include "g4.cuh"
#include "g4_launch.hpp"

namespace scb {

template <int K>
static cudaError_t launch_k(int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in, uint64_t* const* outp,
                            const ElemArg& r, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid) {
    auto kern = g4::k_fold_round_g4<K>;
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
    kern<<<grid, kThreads, 0, stream>>>(f, ti, to, r, n_quads, partials, ticket, res, pa);
    return cudaGetLastError();
}

cudaError_t launch_fold_round_g4(int K, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in,
                                 uint64_t* const* outp, const ElemArg& r, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res,
                                 const PeerArg& pa, int max_grid) {
    switch (K) {
        case 1: return launch_k<1>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
        case 2: return launch_k<2>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
        case 3: return launch_k<3>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
        case 4: return launch_k<4>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace scb
