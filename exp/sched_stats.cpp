#include <cstdio>
#include <map>
#include "../halo2ecc-s_b200/csrc/circuits.h"
#include "../halo2ecc-s_b200/csrc/schedule.h"
using namespace h2e;
int main(int argc, char** argv) {
    int kind = argc > 1 ? atoi(argv[1]) : 2;
    uint64_t params[1] = {argc > 2 ? (uint64_t)atoll(argv[2]) : 0};
    Context ctx;
    build_circuit(ctx, kind, params, 1);
    Schedule sc = levelise(ctx.shape);
    size_t n = sc.program.size(), nl = sc.level_start.size() - 1;
    printf("kind %d: instrs %zu levels %zu avg width %.1f max width %u slots %zu\n", kind, n, nl, (double)n / nl, sc.max_width, ctx.shape.slot_cell.size());
    // width histogram and "time at width" assuming unit cost
    std::map<int, size_t> hist;
    size_t heavy_levels = 0;
    for (size_t l = 0; l < nl; l++) {
        uint32_t w = sc.level_start[l + 1] - sc.level_start[l];
        int b = w <= 1 ? 1 : w <= 2 ? 2 : w <= 4 ? 4 : w <= 8 ? 8 : w <= 16 ? 16 : w <= 32 ? 32 : w <= 64 ? 64 : 128;
        hist[b]++;
    }
    for (auto& kv : hist) printf("  width<=%d: %zu levels\n", kv.first, kv.second);
    // cost-weighted critical path: cost per op
    auto cost = [](const Instr& in) { switch (in.op) { case OP_INT_MUL: return 10.0; case OP_DIV_CORE: return 60.0; case OP_IS_INT_ZERO: return 120.0; case OP_REDUCE: return 4.0; default: return 1.0; } };
    double total = 0, crit = 0;
    for (size_t l = 0; l < nl; l++) { double mx = 0; for (uint32_t i = sc.level_start[l]; i < sc.level_start[l + 1]; i++) { total += cost(sc.program[i]); mx = std::max(mx, cost(sc.program[i])); } crit += mx; }
    printf("  cost-weighted: total %.0f, level-critical-path %.0f, ratio %.1f\n", total, crit, total / crit);
    return 0;
}
