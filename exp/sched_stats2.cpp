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
    Schedule sc = levelise(ctx.shape, true, argc > 3 ? atoi(argv[3]) != 0 : true);
    size_t nl = sc.level_start.size() - 1;
    std::map<int, size_t> hist;
    std::map<int, size_t> heaviest;  // op of heaviest critical instr per level
    size_t ncrit = 0, ntail = 0;
    for (size_t l = 0; l < nl; l++) {
        uint32_t w = sc.level_mid[l] - sc.level_start[l];
        ncrit += w; ntail += sc.level_start[l + 1] - sc.level_mid[l];
        int b = w == 0 ? 0 : w <= 1 ? 1 : w <= 2 ? 2 : w <= 4 ? 4 : w <= 8 ? 8 : w <= 16 ? 16 : w <= 32 ? 32 : w <= 64 ? 64 : 128;
        hist[b]++;
        if (w) heaviest[sc.program[sc.level_start[l]].op]++;
    }
    printf("levels %zu crit %zu tail %zu\n", nl, ncrit, ntail);
    { double tot = 0; size_t hl = 0; for (size_t l = 0; l < nl; l++) { uint32_t w = sc.level_mid[l] - sc.level_start[l]; double c = w ? instr_cost(sc.program[sc.level_start[l]]) : 0; if (c >= 2500) hl++; tot += c + 2300; } printf("heavy levels %zu, est. critical time %.1f Mcycles = %.1f ms\n", hl, tot / 1e6, tot / 1.9e6); }
    for (auto& kv : hist) printf("  crit width<=%d: %zu levels\n", kv.first, kv.second);
    for (auto& kv : heaviest) printf("  heaviest op %d: %zu levels\n", kv.first, kv.second);
    std::map<int, size_t> opc;
    for (auto& in : sc.program) opc[in.op * 2 + ((in.flags & 0x80) ? 1 : 0)]++;
    for (auto& kv : opc) printf("  op %d %s: %zu\n", kv.first / 2, kv.first & 1 ? "deferred" : "critical", kv.second);
    return 0;
}
