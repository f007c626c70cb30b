// ORACLE (test infrastructure, NOT product code). Whole-circuit harness shapes.
#pragma once
#include "chips.h"
namespace orc {
inline int run_circuit(int kind, const uint64_t* params, size_t n_params, const std::vector<BN>& in, std::shared_ptr<Context> ctx) {
    throw OraclePanic{"circuit kind not implemented"};
}
}  // namespace orc
