// internal.hpp -- handle layouts shared by engine.cu (device engine, level 1 of the C ABI) and
// protocol.cpp (host protocol mirror, level 2).  No CUDA types here.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <vector>

#include "fielddesc.hpp"
#include "host/hostfield.hpp"

struct scb_poly;

namespace scb {

struct FieldImpl {
    FieldDesc d;
    HostField h;
    uint32_t policy;  // Policy
};

void set_error(const char* fmt, ...);

// engine.cu: should all remaining rounds of this product polynomial run in a resident kernel (tail.cuh for small
// tables, persist.cuh for large ones)?  Decided per field policy and table size.  need_grid: only if the grid-wide
// kernel would be used (the one that can exchange partial sums with peer GPUs every round).
bool resident_rounds_ok(const struct ::scb_poly* p, bool need_grid);
// engine.cu: can this polynomial's proof run two rounds per pass over the tables (pairs.cuh)?
bool pair_passes_ok(const struct ::scb_poly* p);
// engine.cu: scb_poly_grid_evals; make_w21: also leave the 21-bit triples for a first pair pass that runs alone (pairs.cuh)
int poly_grid_evals_ex(const struct ::scb_poly* p, uint64_t* out_elems, bool make_w21);
// engine.cu: did the grid pass leave 21-bit triples of this polynomial's tables (pairs.cuh)?
bool poly_has_w21(const struct ::scb_poly* p);
// engine.cu: are this polynomial's tables stored as packed uint32 (packed.cuh)?
bool poly_is_packed(const struct ::scb_poly* p);

}  // namespace scb

struct scb_field {
    std::shared_ptr<scb::FieldImpl> impl;
};
