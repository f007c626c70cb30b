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

// slots written by OP_INT_MUL_HEAD (and read back by OP_INT_MUL_TAIL): limb accumulator cells and
// native cells of the rem block and of the d block (layout: see IntBlock in vm_ops.cuh)
inline std::vector<uint32_t> head_cells(const Instr& in) {
    unsigned L = limbs_of_field(in.field);
    std::vector<uint32_t> r;
    for (unsigned blk = 0; blk < 2; blk++) {
        uint32_t base = in.out + blk * (8 * L - 1);
        for (unsigned i = 0; i < L; i++) r.push_back(base + (i < L - 1 ? 7 * i + 6 : 7 * (L - 1) + 4));
        r.push_back(base + 8 * L - 2);
    }
    return r;
}

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
        case OP_INT_MUL_HEAD:
            range(0, L);
            range(L + 1, 2 * L + 1);
            break;
        case OP_INT_MUL_TAIL:
            range(0, 2 * L + 2);
            for (uint32_t s : head_cells(in)) out.push_back(s);
            break;
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
    std::vector<uint32_t> level_mid;   // [level_start[l], level_mid[l]) critical ops, [level_mid[l], level_start[l+1]) deferred (TAIL) ops
    uint32_t max_width = 0;
};

inline Schedule levelise(const Shape& sh, bool split_int_mul = true) {
    // team-mode program: every OP_INT_MUL becomes HEAD (critical path) + TAIL (off the critical path)
    std::vector<Instr> p;
    p.reserve(sh.program.size() * 5 / 4);
    std::vector<uint32_t> block_end;  // one past the last slot of the instruction's block
    for (size_t i = 0; i < sh.program.size(); i++) {
        uint32_t end = i + 1 < sh.program.size() ? sh.program[i + 1].out : (uint32_t)sh.slot_cell.size();
        if (split_int_mul && sh.program[i].op == OP_INT_MUL) {
            Instr h = sh.program[i], ta = sh.program[i], tb = sh.program[i];
            h.op = OP_INT_MUL_HEAD;
            ta.op = OP_INT_MUL_TAIL;
            ta.flags = 1;  // assign blocks
            tb.op = OP_INT_MUL_TAIL;
            tb.flags = 2;  // constraint rows
            p.push_back(h);
            block_end.push_back(h.out);  // HEAD's slots are claimed below
            p.push_back(ta);
            block_end.push_back(h.out);  // nothing depends on TAIL cells; the block is claimed by the last piece
            p.push_back(tb);
            block_end.push_back(end);
        } else {
            p.push_back(sh.program[i]);
            block_end.push_back(end);
        }
    }
    size_t n = p.size();
    std::vector<uint32_t> producer(sh.slot_cell.size(), 0);
    for (size_t i = 0; i < n; i++)
        for (uint32_t s = p[i].out; s < block_end[i]; s++) producer[s] = (uint32_t)i;
    for (size_t i = 0; i < n; i++)
        if (p[i].op == OP_INT_MUL_HEAD)
            for (uint32_t s : head_cells(p[i])) producer[s] = (uint32_t)i;
    std::vector<uint32_t> level(n, 0);
    std::vector<uint32_t> ins;
    uint32_t n_levels = 0;
    for (size_t i = 0; i < n; i++) {
        instr_inputs(p[i], sh, ins);
        uint32_t lv = 0;
        for (uint32_t s : ins) {
            if (s >= p[i].out && p[i].op != OP_INT_MUL_TAIL) throw std::logic_error("instruction reads a slot it has not seen produced");
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
            case OP_INT_MUL_HEAD: return 11;
            case OP_INT_MUL_TAIL: return (in.flags & 2) ? 8 : 3;
            case OP_IS_ZERO: return 40;
            case OP_REDUCE: return 5;
            default: return 1;
        }
    };
    sc.level_mid.resize(n_levels);
    for (uint32_t l = 0; l < n_levels; l++) {
        auto b = sc.program.begin() + sc.level_start[l], e = sc.program.begin() + sc.level_start[l + 1];
        // deferred ops (nothing depends on their output) go last; the rest heavy-first
        auto mid = std::stable_partition(b, e, [](const Instr& in) { return in.op != OP_INT_MUL_TAIL; });
        sc.level_mid[l] = (uint32_t)(mid - sc.program.begin());
        auto by_weight = [&](const Instr& x, const Instr& y) {
            int wa = weight(x), wb = weight(y);
            return wa != wb ? wa > wb : x.op < y.op;
        };
        std::stable_sort(b, mid, by_weight);
        std::stable_sort(mid, e, by_weight);
    }
    return sc;
}

}  // namespace h2e
