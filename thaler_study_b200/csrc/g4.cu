// g4.cu -- translation unit of the 4-limb kernels in 32-bit-limb arithmetic (g4.cuh, g4_mle.cuh); its own object file so
// that the ~50 KB unrolled Montgomery bodies compile in parallel with engine.cu and g29.cu.
#include <cuda_runtime.h>

#include "g4.cuh"
#include "g4_mle.cuh"
#include "g4_launch.hpp"

namespace scb {

template <int K, int MINB>
static cudaError_t launch_k(int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in, uint64_t* const* outp,
                            const ElemArg& r, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid) {
    auto kern = g4::k_fold_round_g4<K, MINB>;
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


// ---- fourth generation: wide accumulators in shared memory (k_fold_round_g4w / k_round_evals_g4w)
// nb_cached: the caller's per-kernel static (kernels that differ only in a bool template argument share a function TYPE,
// so a static in here would be shared between them and the second one would never get its shared-memory attribute)
constexpr int kWideMaxDev = 16;
struct WideCache {  // per device: the shared-memory opt-in is a per-context attribute
    int nb[kWideMaxDev] = {};
};
template <class Kern>
static cudaError_t wide_grid(Kern kern, WideCache& cache, size_t smem, int blocks_per_sm_cap, int sms, uint64_t items, int max_grid, int* grid_out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kWideMaxDev) dev = 0;
    int& nb_cached = cache.nb[dev];
    if (nb_cached == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem) != cudaSuccess || nb < 1) nb = 1;
        nb_cached = nb;
    }
    int nb = nb_cached;
    if (blocks_per_sm_cap > 0 && blocks_per_sm_cap < nb) nb = blocks_per_sm_cap;
    uint64_t want = (items + kThreads - 1) / kThreads, cap = (uint64_t)sms * nb;
    if (cap > (uint64_t)max_grid) cap = max_grid;
    if (want < 1) want = 1;
    *grid_out = (int)(want < cap ? want : cap);
    return cudaSuccess;
}
template <int K, bool P0ONE, int MINB = 2>
static cudaError_t launch_kw(int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in, uint64_t* const* outp,
                             const g4::FoldTab& r, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid) {
    auto kern = g4::k_fold_round_g4w<K, P0ONE, MINB>;
    const size_t smem = g4::wacc_smem_bytes(g4::n_sums(K));
    static WideCache nb_cached;
    int grid = 1;
    const cudaError_t e = wide_grid(kern, nb_cached, smem, blocks_per_sm_cap, sms, n_quads, max_grid, &grid);
    if (e != cudaSuccess) return e;
    TabsIn<K> ti;
    TabsOut<K> to;
    for (int k = 0; k < K; ++k) {
        ti.p[k] = in[k];
        to.p[k] = outp[k];
    }
    kern<<<grid, kThreads, smem, stream>>>(f, ti, to, r, n_quads, partials, ticket, res, pa);
    return cudaGetLastError();
}
template <int K, bool P0ONE, int MINB = 2>
static cudaError_t launch_rw(int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in, uint64_t n_pairs,
                             uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid) {
    auto kern = g4::k_round_evals_g4w<K, P0ONE, MINB>;
    const size_t smem = g4::wacc_smem_bytes(g4::n_sums(K) + 1);
    static WideCache nb_cached;
    int grid = 1;
    const cudaError_t e = wide_grid(kern, nb_cached, smem, blocks_per_sm_cap, sms, n_pairs, max_grid, &grid);
    if (e != cudaSuccess) return e;
    TabsIn<K> ti;
    for (int k = 0; k < K; ++k) ti.p[k] = in[k];
    kern<<<grid, kThreads, smem, stream>>>(f, ti, n_pairs, partials, ticket, res, pa);
    return cudaGetLastError();
}
bool g4w_supported(const FieldDesc& f, int K) { return f.n == 4 && f.bits <= 255 && K >= 2 && K <= 4; }
bool g4_p0one(const FieldDesc& f) { return (uint32_t)f.p[0] == 1u; }

int g_g4w_minb = 0;  // 0: measured defaults (K = 3 fold kernel: 3 resident CTAs per SM, 80 registers -- 10.16 against 10.66 ms at 2^28 x 3;
                     // everything else 2); 1 / 2 / 3 force the K = 3 variants (option g4_blocks; profiles/r02_mont29.md)
cudaError_t launch_fold_round_g4w(int K, bool p0one, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in,
                                  uint64_t* const* outp, const g4::FoldTab& r, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res,
                                  const PeerArg& pa, int max_grid) {
    if (K == 3) {
        const int mb = g_g4w_minb == 0 ? 3 : g_g4w_minb;
        if (mb == 1 && p0one) return launch_kw<3, true, 1>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
        if (mb == 3) return p0one ? launch_kw<3, true, 3>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid)
                                  : launch_kw<3, false, 3>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
    }
#define SCB_KW(KK)                                                                                                                          \
    case KK:                                                                                                                                \
        return p0one ? launch_kw<KK, true>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid)   \
                     : launch_kw<KK, false>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
    switch (K) {
        SCB_KW(2)
        SCB_KW(3)
        SCB_KW(4)
        default: return cudaErrorInvalidValue;
    }
#undef SCB_KW
}
cudaError_t launch_round_evals_g4w(int K, bool p0one, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in,
                                   uint64_t n_pairs, uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid) {
    if (g_g4w_minb == 1 && K == 3 && p0one) return launch_rw<3, true, 1>(blocks_per_sm_cap, sms, stream, f, in, n_pairs, partials, ticket, res, pa, max_grid);
#define SCB_RW(KK)                                                                                                                  \
    case KK:                                                                                                                        \
        return p0one ? launch_rw<KK, true>(blocks_per_sm_cap, sms, stream, f, in, n_pairs, partials, ticket, res, pa, max_grid)    \
                     : launch_rw<KK, false>(blocks_per_sm_cap, sms, stream, f, in, n_pairs, partials, ticket, res, pa, max_grid);
    switch (K) {
        SCB_RW(2)
        SCB_RW(3)
        SCB_RW(4)
        default: return cudaErrorInvalidValue;
    }
#undef SCB_RW
}

// ---- MLE evaluation of a 4-limb table with unreduced products (g4_mle.cuh)
template <bool P0ONE>
static cudaError_t launch_mle(int sms, cudaStream_t stream, const FieldDesc& f, const PointArg& pt, const uint64_t* evals, uint32_t v_local, uint32_t v_total,
                              uint64_t row0, uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid) {
    auto kern = g4::k_mle_eval_fused_g4<P0ONE>;
    constexpr size_t smem = MleFusedCfg<PolGN<4>>::smem_bytes;
    static WideCache nb_cached;
    int grid = 1;
    const uint64_t n_rows = (1ull << v_local) >> MleFusedCfg<PolGN<4>>::LB;
    const cudaError_t e = wide_grid(kern, nb_cached, smem, 0, sms, n_rows * 32, max_grid, &grid);
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, smem, stream>>>(f, pt, evals, v_local, v_total, row0, partials, ticket, res, pa);
    return cudaGetLastError();
}
cudaError_t launch_mle_eval_fused_g4(bool p0one, int sms, cudaStream_t stream, const FieldDesc& f, const PointArg& pt, const uint64_t* evals, uint32_t v_local,
                                     uint32_t v_total, uint64_t row0, uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid) {
    return p0one ? launch_mle<true>(sms, stream, f, pt, evals, v_local, v_total, row0, partials, ticket, res, pa, max_grid)
                 : launch_mle<false>(sms, stream, f, pt, evals, v_local, v_total, row0, partials, ticket, res, pa, max_grid);
}

cudaError_t launch_fold_round_g4(int K, int minb, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in,
                                 uint64_t* const* outp, const ElemArg& r, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res,
                                 const PeerArg& pa, int max_grid) {
    switch (K) {
        case 1: return minb >= 3 ? launch_k<1, 3>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid) : launch_k<1, 2>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
        case 2: return minb >= 3 ? launch_k<2, 3>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid) : launch_k<2, 2>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
        case 3: return minb >= 3 ? launch_k<3, 3>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid) : launch_k<3, 2>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
        case 4: return minb >= 3 ? launch_k<4, 3>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid) : launch_k<4, 2>(blocks_per_sm_cap, sms, stream, f, in, outp, r, n_quads, partials, ticket, res, pa, max_grid);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace scb
