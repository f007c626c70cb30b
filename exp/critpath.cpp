// longest dependency path of the (split) program under a per-op latency model
#include <cstdio>
#include <map>
#include "../halo2ecc-s_b200/csrc/circuits.h"
#include "../halo2ecc-s_b200/csrc/schedule.h"
using namespace h2e;
int main(int argc, char** argv) {
    int kind = argc > 1 ? atoi(argv[1]) : 2;
    uint64_t params[1] = {argc > 2 ? (uint64_t)atoll(argv[2]) : 0};
    double hop = argc > 3 ? atof(argv[3]) : 1700;
    Context ctx;
    build_circuit(ctx, kind, params, 1);
    const Shape& sh = ctx.shape;
    Schedule sc = levelise(sh);
    // program in level order is a topological order; recompute producers
    const std::vector<Instr>& p = sc.program;
    size_t n = p.size();
    std::vector<uint32_t> producer(sh.slot_cell.size(), 0xffffffffu);
    // block ends: sort by out
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return p[a].out != p[b].out ? p[a].out < p[b].out : a < b; });
    // non-head instrs own [out, next distinct out)
    for (size_t k = 0; k < n; k++) {
        size_t i = order[k];
        if (p[i].op == OP_INT_MUL_HEAD || p[i].op == OP_REDUCE_HEAD) continue;
        uint32_t end = (uint32_t)sh.slot_cell.size();
        for (size_t k2 = k + 1; k2 < n; k2++) if (p[order[k2]].out != p[i].out) { end = p[order[k2]].out; break; }
        if ((p[i].op == OP_INT_MUL_TAIL && !(p[i].flags & 2))) continue;
        for (uint32_t s = p[i].out; s < end; s++) producer[s] = (uint32_t)i;
    }
    for (size_t i = 0; i < n; i++) if (p[i].op == OP_INT_MUL_HEAD || p[i].op == OP_REDUCE_HEAD) for (uint32_t s : head_cells(p[i])) producer[s] = (uint32_t)i;
    std::vector<double> fin(n, 0);
    std::vector<uint32_t> ins;
    double best = 0; size_t nh = 0;
    std::vector<uint32_t> depthH(n, 0);
    uint32_t maxH = 0;
    for (size_t i = 0; i < n; i++) {
        instr_inputs(p[i], sh, ins);
        double st = 0; uint32_t dh = 0;
        for (uint32_t s : ins) { uint32_t pr = producer[s]; if (pr == 0xffffffffu || pr >= i) continue; st = std::max(st, fin[pr] + hop); dh = std::max(dh, depthH[pr]); }
        double c = instr_cost(p[i]);
        if (p[i].flags & 0x80) c = 0;  // deferred: not on the path
        fin[i] = st + c;
        depthH[i] = dh + ((p[i].op == OP_INT_MUL_HEAD || p[i].op == OP_REDUCE_HEAD) ? 1 : 0);
        maxH = std::max(maxH, depthH[i]);
        best = std::max(best, fin[i]);
    }
    printf("kind %d: %zu instrs, dataflow critical path %.1f Mcycles (hop %.0f) = %.1f ms @1.9GHz; max HEAD depth %u; levels %zu\n", kind, n, best / 1e6, hop, best / 1.9e6, maxH, sc.level_start.size() - 1);
    return 0;
}
