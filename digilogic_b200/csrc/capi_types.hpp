// capi_types.hpp — the opaque handle types of include/digilogic_b200.h (shared by capi.cpp and batch.cu).
#pragma once
#include <memory>

#include "engine.hpp"
#include "netlist.hpp"

struct b2_builder {
    b2::Builder b;
};
struct b2_sim {
    std::unique_ptr<b2::Engine> e;
};
