// Shapes of the reference's own test circuits (MSM, pairing checks), traced through the chip mirror.
#pragma once
#include "tracer.h"
namespace h2e {
inline void build_circuit(Context& ctx, int kind, const uint64_t* params, size_t n_params) {
    throw std::runtime_error("circuit kind not implemented");
}
}  // namespace h2e
