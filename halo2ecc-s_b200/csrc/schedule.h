// Levelised schedule of a traced program.
//
// The program a shape traces is straight-line and every macro-op writes its own rows, so the only
// ordering that matters is data flow between macro-ops (operand slots). Instructions are grouped
// into dependency levels; all instructions of one level are independent and can run on different
// warps of the team that owns an instance tile. This is how one circuit instance (e.g. a pairing
// check: ~175k macro-ops, critical path ~10k) is spread over many warps instead of one thread.
#pragma once
#include <algorithm>
#include <vector>

#include "tracer.h"

namespace h2e {

inline unsigned limbs_of_field(uint8_t f) { return field_info((Field)f).limbs; }

// operand slots an instruction reads
inline void instr_inputs(const Instr& in, const Shape& sh, std::vector<uint32_t>& out) {
    out.clear();
    unsigned L = limbs_of_field(in.field);
    auto range = [&](unsigned from, unsigned to) {
        for (unsigned i = from; i < to; i++) out.push_back(in.a[i]);
    };
    switch (in.op) {
        case OP_INT_ADD:
        case OP_INT_SUB: range(0, 2 * L); break;
        case OP_INT_NEG:
        case OP_MUL_SMALL:
        case OP_SUM_ASSERT_ZERO: range(0, L); break;
        case OP_REDUCE:
        case OP_IS_INT_ZERO:
        case OP_CACHE_INT: range(0, L + 1); break;
        case OP_INT_MUL:
        case OP_DIV_CORE: range(0, 2 * L + 2); break;
        case OP_MASK_INT: range(0, L + 2); break;
        case OP_BISEC_INT: range(0, 2 * L + 3); break;
        case OP_LINSUM:
            for (unsigned i = 0; i < in.a[0]; i++) out.push_back(in.a[2 + 2 * i]);
            break;
        case OP_MUL:
        case OP_BOOL:
        case OP_ASSERT_EQUAL: range(0, 2); break;
        case OP_BISEC: range(0, 3); break;
        case OP_IS_ZERO:
        case OP_ASSERT_CONST:
        case OP_DECOMPOSE_NATIVE:
        case OP_DECOMPOSE_LIMB: range(0, 1); break;
        case OP_SELECT_INT:
            out.push_back(in.a[0]);
            for (unsigned i = 0; i < in.a[2] * (L + 1); i++) out.push_back(sh.tables[in.a[1] + i]);
            break;
        default: break;  // input-only ops
    }
}

struct Schedule {
    std::vector<Instr> program;        // instructions sorted by (level, opcode)
    std::vector<uint32_t> level_start; // level l = program[level_start[l] .. level_start[l+1])
    uint32_t max_width = 0;
};

inline Schedule levelise(const Shape& sh) {
    const std::vector<Instr>& p = sh.program;
    size_t n = p.size();
    std::vector<uint32_t> producer(sh.slot_cell.size(), 0);
    for (size_t i = 0; i < n; i++) {
        uint32_t end = i + 1 < n ? p[i + 1].out : (uint32_t)sh.slot_cell.size();
        for (uint32_t s = p[i].out; s < end; s++) producer[s] = (uint32_t)i;
    }
    std::vector<uint32_t> level(n, 0);
    std::vector<uint32_t> ins;
    uint32_t n_levels = 0;
    for (size_t i = 0; i < n; i++) {
        instr_inputs(p[i], sh, ins);
        uint32_t lv = 0;
        for (uint32_t s : ins) {
            if (s >= p[i].out) throw std::logic_error("instruction reads a slot it has not seen produced");
            lv = std::max(lv, level[producer[s]] + 1);
        }
        level[i] = lv;
        n_levels = std::max(n_levels, lv + 1);
    }
    Schedule sc;
    std::vector<uint32_t> count(n_levels + 1, 0);
    for (size_t i = 0; i < n; i++) count[level[i] + 1]++;
    for (uint32_t l = 0; l < n_levels; l++) {
        sc.max_width = std::max(sc.max_width, count[l + 1]);
        count[l + 1] += count[l];
    }
    sc.level_start = count;
    std::vector<uint32_t> cursor(count.begin(), count.end() - 1);
    sc.program.resize(n);
    for (size_t i = 0; i < n; i++) sc.program[cursor[level[i]]++] = p[i];
    // inside a level, heavy ops first (they bound the level's duration), equal opcodes adjacent
    auto weight = [](const Instr& in) -> int {
        switch (in.op) {
            case OP_IS_INT_ZERO: return 120;
            case OP_DIV_CORE: return 60;
            case OP_DECOMPOSE_NATIVE: return 20;
            case OP_DECOMPOSE_LIMB: return 12;
            case OP_INT_MUL: return 10;
            case OP_IS_ZERO: return 40;
            case OP_REDUCE: return 5;
            default: return 1;
        }
    };
    for (uint32_t l = 0; l < n_levels; l++)
        std::stable_sort(sc.program.begin() + sc.level_start[l], sc.program.begin() + sc.level_start[l + 1], [&](const Instr& a, const Instr& b) {
            int wa = weight(a), wb = weight(b);
            return wa != wb ? wa > wb : a.op < b.op;
        });
    return sc;
}

}  // namespace h2e
