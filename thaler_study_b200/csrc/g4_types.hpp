// g4_types.hpp -- plain-data kernel arguments of the 4-limb kernels, shared by g4.cuh (device) and the host code.
#pragma once
#include <cstdint>

namespace scb {
namespace g4 {

// Table of a fold pass's fixed multiplier r for ArithT::mul_fixed_raw: t[i] = r * 2^(32 i + 64) mod p as a plain integer
// (eight 32-bit words, canonical), built on the host once per round (engine.cu: g4_fold_table).  256 bytes, by value.
struct FoldTab {
    uint32_t t[8][8];
};

}  // namespace g4
}  // namespace scb
