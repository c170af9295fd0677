// g4_launch.hpp -- host entry of the 4-limb fused fold + message kernel (g4.cuh), shared by g4.cu and engine.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "eqfix.cuh"
#include "g4_types.hpp"
#include "kernels.cuh"

namespace scb {

// number of sums the kernel writes for K tables (S_0, S_inf, S_2, ..): see g4.cuh
inline int g4_n_sums(int K) { return K; }
// Launches k_fold_round_g4<K>: folds variable 0 of the K tables by r (outp: K tables of 2 * n_quads elements) and
// accumulates the sums of the folded tables' round message.  res: g4_n_sums(K) canonical elements (mapped host or
// device memory).  One resident wave of CTAs (occupancy calculator, optionally capped).  minb: the kernel variant compiled
// for that many resident CTAs per SM (2: 128 registers per thread; 3: 80, with a few spills to local memory).
cudaError_t launch_fold_round_g4(int K, int minb, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in,
                                 uint64_t* const* outp, const ElemArg& r, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res,
                                 const PeerArg& pa, int max_grid);

// Third generation (g29.cuh, lazy29.hpp): the same pass in radix 2^29 with lazy carries -- full-rate IMAD.WIDE, no carry
// flags.  r5 = r * 2^5 (field product, Montgomery-256 words); the sums come back multiplied by 2^(-5 (K-1)) (the caller
// multiplies 32^(K-1) back).  g29_supported: moduli of 250..255 bits.
bool g29_supported(const FieldDesc& f);
cudaError_t launch_fold_round_g29(int K, int minb, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in,
                                  uint64_t* const* outp, const ElemArg& r5, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res,
                                  const PeerArg& pa, int max_grid);

// Round-0 message in the same arithmetic (k_round_evals_g29): res receives K + 1 sums in the order S_0, S_inf, S_2 ..
// S_{K-1}, S_1, each short of 2^(5 (K-1)).
cudaError_t launch_round_evals_g29(int K, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in, uint64_t n_pairs,
                                   uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid);

// Fourth generation (g4.cuh, "wide accumulators"): the last product of every message point is accumulated unreduced in
// 544-bit shared-memory accumulators; p0one selects the variant for moduli that are 1 modulo 2^32 (g4_p0one).  K = 2..4.
// launch_fold_round_g4w: sums S_0, S_inf, S_2 .. S_{K-1} like launch_fold_round_g4, the challenge as its fold table
// (g4_types.hpp); launch_round_evals_g4w: K + 1 sums in
// the order S_0, S_inf, S_2 .. S_{K-1}, S_1 (no claim in round 0).
extern int g_g4w_minb;  // option g4_blocks4: resident CTAs per SM the K = 3 kernels are compiled for (0: measured defaults)
bool g4w_supported(const FieldDesc& f, int K);
bool g4_p0one(const FieldDesc& f);
cudaError_t launch_fold_round_g4w(int K, bool p0one, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in,
                                  uint64_t* const* outp, const g4::FoldTab& r, uint64_t n_quads, uint64_t* partials, unsigned int* ticket, uint64_t* res,
                                  const PeerArg& pa, int max_grid);
cudaError_t launch_round_evals_g4w(int K, bool p0one, int blocks_per_sm_cap, int sms, cudaStream_t stream, const FieldDesc& f, const uint64_t* const* in,
                                   uint64_t n_pairs, uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid);

// MLE evaluation of a 4-limb table in one launch with unreduced products (g4_mle.cuh); arguments as k_mle_eval_fused.
// Needs f.bits <= 255 (the unreduced sums are sized on p < 2^255).
cudaError_t launch_mle_eval_fused_g4(bool p0one, int sms, cudaStream_t stream, const FieldDesc& f, const PointArg& pt, const uint64_t* evals, uint32_t v_local,
                                     uint32_t v_total, uint64_t row0, uint64_t* partials, unsigned int* ticket, uint64_t* res, const PeerArg& pa, int max_grid);

}  // namespace scb
