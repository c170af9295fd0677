// fielddesc.hpp -- plain-data field descriptor shared by host and device code.
#pragma once
#include <cstdint>

namespace scb {

constexpr int kMaxLimbs = 4;

// Passed to kernels by value (lives in the constant bank).
struct FieldDesc {
    uint64_t p[kMaxLimbs];
    uint64_t one[kMaxLimbs];  // R mod p
    uint64_t r2[kMaxLimbs];   // R^2 mod p
    uint64_t inv;             // -p^{-1} mod 2^64
    uint32_t n;               // limbs
    uint32_t bits;            // bit length of p
};

enum Policy : uint32_t { POL_SP = 0, POL_G1 = 1, POL_G4 = 4 };

}  // namespace scb
