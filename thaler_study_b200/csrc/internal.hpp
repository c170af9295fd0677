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

namespace scb {

struct FieldImpl {
    FieldDesc d;
    HostField h;
    uint32_t policy;  // Policy
};

void set_error(const char* fmt, ...);

}  // namespace scb

struct scb_field {
    std::shared_ptr<scb::FieldImpl> impl;
};
