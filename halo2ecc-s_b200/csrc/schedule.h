// Levelised schedule of a traced program.
//
// The program a shape traces is straight-line and every macro-op writes its own rows, so the only
// ordering that matters is data flow between macro-ops (operand slots). Instructions are grouped
// into dependency levels; all instructions of one level are independent and can run on different
// warps of the team that owns an instance tile. This is how one circuit instance (e.g. a pairing
// check: ~175k macro-ops, critical path ~10k) is spread over many warps instead of one thread.
#pragma once
#include <cstdlib>
#include <algorithm>
#include <set>
#include <vector>

#include "tracer.h"

namespace h2e {

inline unsigned limbs_of_field(uint8_t f) { return field_info((Field)f).limbs; }

// slots written by OP_INT_MUL_HEAD (and read back by OP_INT_MUL_TAIL): limb accumulator cells and
// native cells of the rem block and of the d block (layout: see IntBlock in vm_ops.cuh)
inline std::vector<uint32_t> head_cells(const Instr& in) {
    unsigned L = limbs_of_field(in.field);
    std::vector<uint32_t> r;
    if (in.op == OP_IS_INT_ZERO_HEAD) return {in.a[13]};  // the condition cell
    if (in.op == OP_DIV_HEAD_S || in.op == OP_DIV_TAIL) {
        // limb accumulators + native of the c block
        for (unsigned i = 0; i < L; i++) r.push_back(in.out + (i < L - 1 ? 7 * i + 6 : 7 * (L - 1) + 4));
        r.push_back(in.out + 8 * L - 2);
        return r;
    }
    if (in.op == OP_REDUCE_HEAD || in.op == OP_REDUCE_TAIL) {
        // rem block accumulators + native, then the quotient cell (first cell of assign_common(d))
        for (unsigned i = 0; i < L; i++) r.push_back(in.out + (i < L - 1 ? 7 * i + 6 : 7 * (L - 1) + 4));
        r.push_back(in.out + 8 * L - 2);
        r.push_back(in.out + 8 * L - 1);
        return r;
    }
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
        case OP_INT_ADD: range(0, 2 * L + 2); break;
        case OP_INT_SUB:
            range(0, 2 * L);
            range(2 * L + 1, 2 * L + 3);
            break;
        case OP_INT_NEG:
        case OP_MUL_SMALL:
            range(0, L);
            range(L + 1, L + 2);
            break;
        case OP_SUM_ASSERT_ZERO: range(0, L); break;
        case OP_REDUCE:
        case OP_IS_INT_ZERO:
        case OP_IS_INT_ZERO_HEAD:
        case OP_CACHE_INT: range(0, L + 1); break;
        case OP_IS_INT_ZERO_TAIL:  // up to three blocks share one inversion: flags & 3 = number of blocks
            for (unsigned j = 0; j < (in.flags & 3u); j++) range(j * (L + 1), (j + 1) * (L + 1));
            break;
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
        case OP_REDUCE_HEAD: range(0, L); break;
        case OP_DIV_INV:
            for (unsigned j = 0; j < std::max(1u, in.flags & 3u); j++) range(j * (L + 1), j * (L + 1) + L);
            break;
        case OP_DIV_CORE_S:
            range(0, 2 * L + 2);
            out.push_back((uint32_t)sh.slot_cell.size() + in.a[2 * L + 2]);  // pseudo-slot of the scratch entry
            break;
        case OP_DIV_HEAD_S:
            range(0, L);
            out.push_back((uint32_t)sh.slot_cell.size() + in.a[2 * L + 2]);
            break;
        case OP_DIV_TAIL:
            range(0, 2 * L + 2);
            for (uint32_t s : head_cells(in)) out.push_back(s);
            break;
        case OP_REDUCE_TAIL:
            range(0, L + 1);
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
        case OP_BOOLV:
            for (unsigned i = 0; i < 2 * in.a[1]; i++) out.push_back(sh.tables[in.a[2] + i]);
            break;
        case OP_CHIV:
            for (unsigned i = 0; i < 3 * in.a[1]; i++) out.push_back(sh.tables[in.a[2] + i]);
            break;
        default: break;  // input-only ops
    }
}

// rough single-warp latency of a macro-op in cycles (measured on B200, team mode), for load balancing
inline uint32_t instr_cost(const Instr& in) {
    switch (in.op) {
        case OP_IS_INT_ZERO: return 90000;
        case OP_IS_INT_ZERO_HEAD: return 2500;
        case OP_IS_INT_ZERO_TAIL: return 70000 + 12000 * (in.flags & 3u);
        case OP_DIV_CORE: return 120000;
        case OP_DIV_INV: return 65000 + 7500 * (std::max(1u, in.flags & 3u) - 1);
        case OP_DIV_CORE_S: return 32000;
        case OP_DIV_HEAD_S: return 8000;
        case OP_DIV_TAIL: return 28000;
        case OP_IS_ZERO: return 80000;
        case OP_DECOMPOSE_NATIVE: return 30000;
        case OP_DECOMPOSE_LIMB: return 15000;
        case OP_INT_MUL: return 9000;
        case OP_INT_MUL_HEAD: return 4500;
        case OP_INT_MUL_TAIL: return 7000;
        case OP_REDUCE: return 5000;
        case OP_REDUCE_HEAD: return 3000;
        case OP_REDUCE_TAIL: return 3000;
        case OP_ASSIGN_W: return 3000;
        case OP_LINSUM: return 3000;
        case OP_BOOLV: return 1500 + 400 * in.a[1];
        case OP_CHIV: return 1500 + 700 * in.a[1];
        default: return 1500;
    }
}

struct Schedule;
inline void merge_div_inv(Schedule& sc, unsigned kmax);

struct Schedule {
    std::vector<Instr> program;        // instructions sorted by (level, opcode)
    std::vector<uint32_t> level_start; // level l = program[level_start[l] .. level_start[l+1])
    std::vector<uint32_t> level_mid;   // [level_start[l], level_mid[l]) critical ops, [level_mid[l], level_start[l+1]) deferred (TAIL) ops
    std::vector<uint32_t> pred_off, preds;  // CSR: producers of program[k] (positions in `program`, all < k's level)
    uint32_t n_scratch = 0;                 // scratch entries (64 bytes per instance each) the program uses
    uint32_t max_width = 0;
};

inline Schedule levelise(const Shape& sh, bool split_int_mul = true, bool align_heavy = true) {
    // team-mode program: every OP_INT_MUL becomes HEAD (critical path) + TAIL (off the critical path)
    std::vector<Instr> p;
    p.reserve(sh.program.size() * 5 / 4);
    uint32_t n_scratch = 0;
    std::vector<uint32_t> block_end;  // one past the last slot of the instruction's block
    // is_int_zero TAILs (whole block incl. the Fr inversions; nothing waits for them) are merged: up to three
    // blocks (two for the 4-limb field) become one instruction that inverts all their values with ONE
    // inversion (Montgomery's trick). Operands of block j at a[j(L+1) ..], its first slot in a[11 + j] (j >= 1).
    std::vector<std::pair<Instr, uint32_t>> pend_z;                                  // (original instruction, block end)
    std::vector<std::pair<uint32_t, std::pair<uint32_t, uint32_t>>> extra_blocks;  // (instruction, [out, end)) of blocks j >= 1
    const char* zm = getenv("H2E_ZMERGE");  // tuning: blocks per merged TAIL (1 = one inversion per is_int_zero)
    const unsigned z_merge = zm ? (unsigned)std::max(1, atoi(zm)) : 3u;
    auto flush_z = [&]() {
        if (pend_z.empty()) return;
        unsigned L = limbs_of_field(pend_z[0].first.field);
        Instr m = pend_z[0].first;
        m.op = OP_IS_INT_ZERO_TAIL;
        m.flags = (uint8_t)pend_z.size();
        for (size_t j = 1; j < pend_z.size(); j++) {
            for (unsigned q = 0; q <= L; q++) m.a[j * (L + 1) + q] = pend_z[j].first.a[q];
            m.a[11 + j] = pend_z[j].first.out;
            extra_blocks.push_back({(uint32_t)p.size(), {pend_z[j].first.out, pend_z[j].second}});
        }
        p.push_back(m);
        block_end.push_back(pend_z[0].second);
        pend_z.clear();
    };
    for (size_t i = 0; i < sh.program.size(); i++) {
        uint32_t end = i + 1 < sh.program.size() ? sh.program[i + 1].out : (uint32_t)sh.slot_cell.size();
        if (split_int_mul && sh.program[i].op == OP_INT_MUL) {
            Instr h = sh.program[i], t = sh.program[i];
            h.op = OP_INT_MUL_HEAD;
            t.op = OP_INT_MUL_TAIL;
            t.flags = 3;  // bit 0: the two assign blocks, bit 1: the constraint rows
            p.push_back(h);
            block_end.push_back(h.out);  // HEAD's slots are claimed below
            p.push_back(t);
            block_end.push_back(end);  // nothing depends on TAIL cells
        } else if (split_int_mul && sh.program[i].op == OP_REDUCE) {
            Instr h = sh.program[i], t = sh.program[i];
            h.op = OP_REDUCE_HEAD;
            t.op = OP_REDUCE_TAIL;
            p.push_back(h);
            block_end.push_back(h.out);
            p.push_back(t);
            block_end.push_back(end);
        } else if (split_int_mul && sh.program[i].op == OP_DIV_CORE) {
            // the W inversion only needs the denominator: it becomes its own instruction (result in a
            // scratch entry), so that it runs beside is_int_zero(b) instead of after it
            unsigned L = limbs_of_field(sh.program[i].field);
            Instr inv = {}, core = sh.program[i];
            inv.op = OP_DIV_INV;
            inv.field = core.field;
            inv.out = core.out;
            for (unsigned k = 0; k < L; k++) inv.a[k] = core.a[L + 1 + k];
            inv.a[L] = n_scratch;
            inv.flags = 1;
            core.a[2 * L + 2] = n_scratch;
            n_scratch++;
            p.push_back(inv);
            block_end.push_back(inv.out);
            // ... and the rest into HEAD (the quotient c, which the next macro-ops read) and a deferred TAIL
            Instr tail = core;
            core.op = OP_DIV_HEAD_S;
            tail.op = OP_DIV_TAIL;
            p.push_back(core);
            block_end.push_back(core.out);  // HEAD's cells are claimed below
            p.push_back(tail);
            block_end.push_back(end);
        } else if (split_int_mul && sh.program[i].op == OP_IS_INT_ZERO) {
            // only the condition (last cell of the block) is read by later macro-ops, and it needs no inversion
            Instr h = sh.program[i];
            h.op = OP_IS_INT_ZERO_HEAD;
            h.a[13] = end - 1;
            p.push_back(h);
            block_end.push_back(h.out);
            if (!pend_z.empty() && pend_z[0].first.field != h.field) flush_z();
            pend_z.push_back({sh.program[i], end});
            if (pend_z.size() >= std::min(z_merge, limbs_of_field(h.field) == 3 ? 3u : 2u)) flush_z();
        } else {
            p.push_back(sh.program[i]);
            block_end.push_back(end);
        }
    }
    flush_z();
    size_t n = p.size();
    const uint32_t n_real_slots = (uint32_t)sh.slot_cell.size();
    std::vector<uint32_t> producer(sh.slot_cell.size() + n_scratch, 0);
    for (size_t i = 0; i < n; i++)
        for (uint32_t s = p[i].out; s < block_end[i]; s++) producer[s] = (uint32_t)i;
    for (auto& e : extra_blocks)
        for (uint32_t s = e.second.first; s < e.second.second; s++) producer[s] = e.first;
    for (size_t i = 0; i < n; i++)
        if (p[i].op == OP_INT_MUL_HEAD || p[i].op == OP_REDUCE_HEAD || p[i].op == OP_IS_INT_ZERO_HEAD || p[i].op == OP_DIV_HEAD_S)
            for (uint32_t s : head_cells(p[i])) producer[s] = (uint32_t)i;
    for (size_t i = 0; i < n; i++)
        if (p[i].op == OP_DIV_INV) producer[n_real_slots + p[i].a[limbs_of_field(p[i].field)]] = (uint32_t)i;
    std::vector<uint32_t> level(n, 0);
    std::vector<uint8_t> consumed(n, 0);  // some later instruction reads one of its cells
    std::vector<uint32_t> ins;
    std::vector<uint32_t> pred_off(n + 1, 0), preds;  // distinct producers of every instruction (CSR)
    preds.reserve(n * 4);
    uint32_t n_levels = 0;
    for (size_t i = 0; i < n; i++) {
        instr_inputs(p[i], sh, ins);
        uint32_t lv = 0;
        size_t first = preds.size();
        for (uint32_t s : ins) {
            if (s >= p[i].out && s < n_real_slots && p[i].op != OP_INT_MUL_TAIL && p[i].op != OP_REDUCE_TAIL && p[i].op != OP_DIV_TAIL && p[i].op != OP_IS_INT_ZERO_TAIL) throw std::logic_error("instruction reads a slot it has not seen produced");
            uint32_t pr = producer[s];
            lv = std::max(lv, level[pr] + 1);
            consumed[pr] = 1;
            bool seen = false;
            for (size_t k = first; k < preds.size(); k++) seen |= preds[k] == pr;
            if (!seen) preds.push_back(pr);
        }
        pred_off[i + 1] = (uint32_t)preds.size();
        level[i] = lv;
        n_levels = std::max(n_levels, lv + 1);
    }
    // deferred = nothing reads its output (int_mul TAILs, asserts, cache rows, ...): flagged in Instr::flags bit 7
    for (size_t i = 0; i < n; i++)
        if (!consumed[i]) p[i].flags |= 0x80;
    // Level alignment. A level lasts as long as its slowest instruction, so heavy instructions (the
    // HEADs: a wide product + Barrett) that have slack are moved, within [as-soon-as-possible,
    // as-late-as-possible], into levels that already contain a heavy instruction. The depth of the
    // schedule does not change; the number of levels that pay for a heavy instruction drops.
    if (align_heavy && n_levels > 0) {
        auto heavy = [&](size_t i) { return !(p[i].flags & 0x80) && instr_cost(p[i]) >= 2500; };
        std::vector<uint32_t> hi(n, n_levels - 1);
        for (size_t i = n; i-- > 0;) {
            if (p[i].flags & 0x80) continue;  // deferred instructions constrain nothing
            for (uint32_t k = pred_off[i]; k < pred_off[i + 1]; k++) hi[preds[k]] = std::min(hi[preds[k]], hi[i] - 1);
        }
        std::vector<uint8_t> is_heavy_level(n_levels, 0);
        for (size_t i = 0; i < n; i++)
            if (heavy(i) && hi[i] == level[i]) is_heavy_level[level[i]] = 1;
        for (size_t i = 0; i < n; i++) {
            uint32_t lv = 0;
            for (uint32_t k = pred_off[i]; k < pred_off[i + 1]; k++) lv = std::max(lv, level[preds[k]] + 1);
            if (heavy(i) && !is_heavy_level[lv]) {
                uint32_t l = lv;
                while (l <= hi[i] && !is_heavy_level[l]) l++;
                if (l <= hi[i]) lv = l;
                else is_heavy_level[lv] = 1;
            }
            level[i] = lv;
            n_levels = std::max(n_levels, lv + 1);  // a deferred instruction may land one level past the last barrier
        }
    }
    Schedule sc;
    sc.n_scratch = n_scratch;
    std::vector<uint32_t> count(n_levels + 1, 0);
    for (size_t i = 0; i < n; i++) count[level[i] + 1]++;
    for (uint32_t l = 0; l < n_levels; l++) {
        sc.max_width = std::max(sc.max_width, count[l + 1]);
        count[l + 1] += count[l];
    }
    sc.level_start = count;
    // order: by level; inside a level critical instructions first (heavy first, equal opcodes adjacent),
    // then the deferred ones (nothing depends on their output)
    std::vector<uint32_t> idx(n);
    {
        std::vector<uint32_t> cursor(count.begin(), count.end() - 1);
        for (size_t i = 0; i < n; i++) idx[cursor[level[i]]++] = (uint32_t)i;
    }
    sc.level_mid.resize(n_levels);
    for (uint32_t l = 0; l < n_levels; l++) {
        auto b = idx.begin() + sc.level_start[l], e = idx.begin() + sc.level_start[l + 1];
        auto mid = std::stable_partition(b, e, [&](uint32_t i) { return (p[i].flags & 0x80) == 0; });
        sc.level_mid[l] = (uint32_t)(mid - idx.begin());
        auto by_weight = [&](uint32_t x, uint32_t y) {
            uint32_t wa = instr_cost(p[x]), wb = instr_cost(p[y]);
            return wa != wb ? wa > wb : p[x].op < p[y].op;
        };
        std::stable_sort(b, mid, by_weight);
        std::stable_sort(mid, e, by_weight);
    }
    std::vector<uint32_t> pos(n);
    sc.program.resize(n);
    for (size_t k = 0; k < n; k++) {
        sc.program[k] = p[idx[k]];
        pos[idx[k]] = (uint32_t)k;
    }
    // producers of every instruction, as positions in sc.program
    sc.pred_off.assign(n + 1, 0);
    sc.preds.reserve(preds.size());
    for (size_t k = 0; k < n; k++) {
        uint32_t i = idx[k];
        for (uint32_t q = pred_off[i]; q < pred_off[i + 1]; q++) sc.preds.push_back(pos[preds[q]]);
        sc.pred_off[k + 1] = (uint32_t)sc.preds.size();
    }
    // Tuning (H2E_DIVMERGE = denominators per merged inversion, up to 3). Off by default: measured on the 1000-point MSM x 192
    // instances, 119.2 ms unmerged, 129.6 ms with pairs, 142.9 ms with triples -- the chains of point additions are bound by
    // the latency of each step, and a merged inversion waits for the slowest of its denominators and adds 3 products per member.
    const char* dm = getenv("H2E_DIVMERGE");
    merge_div_inv(sc, dm ? (unsigned)std::max(1, atoi(dm)) : 1u);
    return sc;
}

// (Optional, see H2E_DIVMERGE above.) The W inversions of one dependency level are independent (an MSM adds one point per window in lockstep: 254 int_divs per
// level), and a safegcd inversion costs ~25 modular products: up to three OP_DIV_INV of a level (two for the 4-limb field: the
// operands must fit one instruction) become ONE instruction that inverts the product of its denominators and recovers each
// inverse with 3 products (Montgomery's trick). The members sit in the same level, so every producer of the merged instruction
// is in an earlier level and every consumer in a later one: order and levels of the schedule do not change.
inline void merge_div_inv(Schedule& sc, unsigned kmax) {
    if (kmax <= 1 || sc.program.empty()) return;
    const size_t n = sc.program.size();
    const size_t n_levels = sc.level_start.size() - 1;
    std::vector<Instr> np;
    np.reserve(n);
    std::vector<uint32_t> remap(n), first, count, nstart(n_levels + 1), nmid(n_levels);
    for (size_t l = 0; l < n_levels; l++) {
        nstart[l] = (uint32_t)np.size();
        nmid[l] = 0xffffffffu;
        const uint32_t end = sc.level_start[l + 1];
        uint32_t k = sc.level_start[l];
        while (k < end) {
            if (k == sc.level_mid[l]) nmid[l] = (uint32_t)np.size();
            Instr m = sc.program[k];
            uint32_t members = 1;
            if (m.op == OP_DIV_INV && !(m.flags & 0x80)) {
                const unsigned L = limbs_of_field(m.field), cap = std::min(kmax, 14u / (L + 1));
                while (members < cap && k + members < end && k + members != sc.level_mid[l] && sc.program[k + members].op == OP_DIV_INV &&
                       sc.program[k + members].field == m.field && (sc.program[k + members].flags & 0x83) == 1) {
                    for (unsigned q = 0; q <= L; q++) m.a[members * (L + 1) + q] = sc.program[k + members].a[q];
                    members++;
                }
                m.flags = (uint8_t)((m.flags & ~3u) | members);
            }
            for (uint32_t j = 0; j < members; j++) remap[k + j] = (uint32_t)np.size();
            first.push_back(k);
            count.push_back(members);
            np.push_back(m);
            k += members;
        }
        if (nmid[l] == 0xffffffffu) nmid[l] = (uint32_t)np.size();  // (level_mid == end of the level)
    }
    nstart[n_levels] = (uint32_t)np.size();
    std::vector<uint32_t> poff(np.size() + 1, 0), pr;
    pr.reserve(sc.preds.size());
    for (size_t g = 0; g < np.size(); g++) {
        const size_t begin = pr.size();
        for (uint32_t o = first[g]; o < first[g] + count[g]; o++)
            for (uint32_t q = sc.pred_off[o]; q < sc.pred_off[o + 1]; q++) {
                const uint32_t x = remap[sc.preds[q]];
                bool seen = false;
                for (size_t t = begin; t < pr.size(); t++) seen |= pr[t] == x;
                if (!seen) pr.push_back(x);
            }
        poff[g + 1] = (uint32_t)pr.size();
    }
    sc.program.swap(np);
    sc.level_start.swap(nstart);
    sc.level_mid.swap(nmid);
    sc.pred_off.swap(poff);
    sc.preds.swap(pr);
}

// ---------------------------------------------------------------------------------------------
// Team mode = dataflow execution of the levelised program by many warps per tile.
//
// A tile (32 instances) is evaluated by `twc` critical team warps and `twt` tail team warps spread
// over several CTAs. Every team warp walks its own contiguous instruction stream in order. An
// instruction starts when the instructions that produced its operands have completed: each critical
// warp publishes "number of instructions of my stream completed" in a per-tile progress array (global
// memory, release store), and every instruction carries the (warp, count) pairs it has to wait for.
// There are no barriers: a warp never waits for instructions it does not depend on.
//
// The host assigns instructions to warps by list scheduling over a latency model (instr_cost, plus
// HOP cycles when producer and consumer sit on different warps), in the topological order of the
// levelised program, so every stream is itself in topological order and the execution cannot deadlock
// as long as all CTAs of a tile are co-resident.
struct DepRec {          // 16 bytes, parallel to the instruction streams
    uint32_t n;          // bits 0..15: number of dependencies; bit 16 / 17: publish progress globally / in shared memory
    uint32_t d[3];       // n <= 3: the dependencies; n > 3: d[0], d[1], then d[2] = index into `extra` of the other n - 2
};
static const uint32_t DEP_SEQ_BITS = 20;  // dependency = warp << 20 | (count - 1)

struct TeamStreams {
    uint32_t twc = 0, twt = 0;
    std::vector<Instr> crit;           // streams of the critical team warps, back to back
    std::vector<DepRec> crit_dep;
    std::vector<uint32_t> crit_off;    // [twc + 1] stream bounds
    std::vector<Instr> tail;           // streams of the tail team warps, back to back
    std::vector<DepRec> tail_dep;
    std::vector<uint32_t> tail_off;    // [twt + 1]
    std::vector<uint32_t> extra;       // overflow dependency lists
    double est_cycles = 0;             // modelled makespan of the critical streams
    uint64_t stat_preds = 0, stat_same_warp = 0, stat_same_warp_recent = 0;  // producer locality of critical operands
};

// Team warp numbering: tw = local_warp * G + rank, so tw % G is the CTA (rank) of the stream. A
// hand-over between two warps of the same CTA goes through shared-memory progress counters and a
// CTA-scope fence (hop_local); between CTAs through the global counters and a GPU-scope release
// (hop_global). DepRec::n bit 16 = publish globally, bit 17 = publish in shared memory.
// Layout of the team: `cta_crit` CTAs run `wc` critical streams each and `cta_tail` CTAs run `wt` tail
// streams each. Mixed layout (every CTA hosts both roles): cta_crit = cta_tail = G, tail_rank0 = 0.
// Split layout (a CTA is all-critical or all-tail, so the record stream of the tail warps does not
// compete with the critical path for issue slots and LSU bandwidth of the same SM): cta_crit + cta_tail
// = G, tail_rank0 = cta_crit. Critical stream w runs in CTA w % cta_crit, tail stream w in CTA
// tail_rank0 + w % cta_tail.
// Code locality (split layout only): the macro-ops are fully unrolled (1.3 MB of SASS) and ncu's stall sampling on
// the MSM shapes shows 47 % "no instruction" -- warps of one SM stream different macro-ops through the 32 KB
// instruction cache, which also evicts the one hot loop of the workload, the safegcd inversion (11.5 KB, shared by
// every W and Fr inversion). Tuning option (H2E_INV_CTAS=1): the first `cta_inv` critical CTAs run nothing but
// OP_DIV_INV and the first `tail_inv` tail CTAs nothing but is_int_zero TAILs (one inversion + out-of-line
// products), so their instruction cache holds that loop for the whole pass. 0 = no dedicated CTAs (default).
struct TeamLayout {
    uint32_t cta_crit, wc, cta_tail, wt, tail_rank0;
    uint32_t cta_inv = 0, tail_inv = 0;
};
// dedicated CTAs by modelled work share; only when the inversions are a sizeable part of the program
inline void dedicate_inversion_ctas(const Schedule& sc, TeamLayout& lay) {
    // Measured (MSM n=1000 x 128, after the out-of-line Fr product shrank the is_int_zero TAIL): 122.5 ms with
    // dedicated CTAs, 105-116 ms without -- the load imbalance costs more than the locality gains. Off by default.
    if (lay.tail_rank0 == 0 || !getenv("H2E_INV_CTAS")) return;  // (mixed layout: never)
    double wc = 0, wi = 0, wt = 0, wz = 0;
    for (const Instr& in : sc.program) {
        double c = instr_cost(in);
        if (in.flags & 0x80) {
            wt += c;
            if (in.op == OP_IS_INT_ZERO_TAIL) wz += c;
        } else {
            wc += c;
            if (in.op == OP_DIV_INV) wi += c;
        }
    }
    if (lay.cta_crit >= 4 && wi >= 0.15 * wc)
        lay.cta_inv = (uint32_t)std::min<int>(std::max<int>((int)(lay.cta_crit * wi / wc + 0.5), 1), (int)lay.cta_crit - 1);
    if (lay.cta_tail >= 4 && wz >= 0.15 * wt)
        lay.tail_inv = (uint32_t)std::min<int>(std::max<int>((int)(lay.cta_tail * wz / wt + 0.5), 1), (int)lay.cta_tail - 1);
}

inline TeamStreams build_team_streams(const Schedule& sc, const TeamLayout& lay, double hop_local = 2000.0, double hop_global = 2000.0) {
    TeamStreams ts;
    const uint32_t G = lay.cta_crit;  // stride of the critical stream numbering
    const uint32_t twc = lay.cta_crit * lay.wc, twt = lay.cta_tail * lay.wt;
    ts.twc = twc;
    ts.twt = twt;
    const size_t n = sc.program.size();
    if (twc >= (1u << (32 - DEP_SEQ_BITS))) throw std::logic_error("too many team warps");
    std::vector<uint32_t> warp_of(n, 0xffffffffu), seq_of(n, 0);
    std::vector<double> finish(n, 0.0);
    std::vector<std::vector<uint32_t>> cs(twc), tl(twt);  // positions in sc.program
    std::vector<double> free_at(twc, 0.0);
    auto hop = [&](uint32_t from, uint32_t to) { return from == to ? 0.0 : (from % G == to % G ? hop_local : hop_global); };
    // ---- critical instructions: list scheduling ----
    // candidate warps: the producers' warps, the first free warp of every producer's CTA, the first free warp overall
    // warp classes: 1 = warps of the CTAs dedicated to OP_DIV_INV (lay.cta_inv), 0 = the others
    auto wclass = [&](uint32_t w) { return (w % G) < lay.cta_inv ? 1 : 0; };
    std::set<std::pair<double, uint32_t>> by_free_c[2];          // (free_at, warp) per class
    std::vector<std::set<std::pair<double, uint32_t>>> cta_free(G);  // per CTA
    for (uint32_t w = 0; w < twc; w++) {
        by_free_c[wclass(w)].insert({0.0, w});
        cta_free[w % G].insert({0.0, w});
    }
    std::vector<uint32_t> cand;
    for (size_t k = 0; k < n; k++) {
        const Instr& in = sc.program[k];
        if (in.flags & 0x80) continue;
        const int cls = lay.cta_inv && in.op == OP_DIV_INV ? 1 : 0;
        auto& by_free = by_free_c[cls];
        cand.clear();
        cand.push_back(by_free.begin()->second);
        for (uint32_t q = sc.pred_off[k]; q < sc.pred_off[k + 1]; q++) {
            uint32_t w = warp_of[sc.preds[q]];
            if (wclass(w) != cls) continue;
            cand.push_back(w);
            cand.push_back(cta_free[w % G].begin()->second);
        }
        uint32_t best = cand[0];
        double best_start = 1e300;
        for (uint32_t w : cand) {
            double r = 0.0;
            for (uint32_t q = sc.pred_off[k]; q < sc.pred_off[k + 1]; q++) {
                uint32_t pp = sc.preds[q];
                r = std::max(r, finish[pp] + hop(warp_of[pp], w));
            }
            double st = std::max(free_at[w], r);
            if (st < best_start) {  // candidates are tried free-warp first: ties keep the warp that is free earliest
                best_start = st;
                best = w;
            }
        }
        warp_of[k] = best;
        seq_of[k] = (uint32_t)cs[best].size();
        if (seq_of[k] >= (1u << DEP_SEQ_BITS)) throw std::logic_error("team warp stream too long");
        cs[best].push_back((uint32_t)k);
        finish[k] = best_start + instr_cost(in);
        by_free.erase({free_at[best], best});
        cta_free[best % G].erase({free_at[best], best});
        free_at[best] = finish[k];
        by_free.insert({free_at[best], best});
        cta_free[best % G].insert({free_at[best], best});
        ts.est_cycles = std::max(ts.est_cycles, finish[k]);
    }
    // statistics for tooling: how many operand producers sit on the consumer's own warp, and how recently
    for (size_t k = 0; k < n; k++) {
        if (sc.program[k].flags & 0x80) continue;
        for (uint32_t q = sc.pred_off[k]; q < sc.pred_off[k + 1]; q++) {
            uint32_t pp = sc.preds[q];
            ts.stat_preds++;
            if (warp_of[pp] == warp_of[k]) {
                ts.stat_same_warp++;
                if (seq_of[k] - seq_of[pp] <= 4) ts.stat_same_warp_recent++;
            }
        }
    }
    // ---- deferred instructions: in order of readiness; to a tail warp of the CTA that produced the last
    // operand (its HEAD) unless that CTA's tail warps are clearly busier than the least loaded one ----
    {
        std::vector<std::pair<double, uint32_t>> td;
        std::vector<uint32_t> home(n, 0);
        for (size_t k = 0; k < n; k++) {
            if (!(sc.program[k].flags & 0x80)) continue;
            double r = 0.0;
            for (uint32_t q = sc.pred_off[k]; q < sc.pred_off[k + 1]; q++)
                if (finish[sc.preds[q]] >= r) {
                    r = finish[sc.preds[q]];
                    home[k] = warp_of[sc.preds[q]] % G;
                }
            td.push_back({r, (uint32_t)k});
        }
        std::stable_sort(td.begin(), td.end());
        std::vector<double> load(twt, 0.0);
        for (auto& e : td) {
            uint32_t wl = 0xffffffffu, wg = 0;
            const int tcls = lay.tail_inv && sc.program[e.second].op == OP_IS_INT_ZERO_TAIL ? 1 : 0;
            wg = 0xffffffffu;
            for (uint32_t v = 0; v < twt; v++) {
                if (((v % lay.cta_tail) < lay.tail_inv ? 1 : 0) != tcls) continue;
                if (wg == 0xffffffffu || load[v] < load[wg]) wg = v;
                if (lay.tail_rank0 + v % lay.cta_tail == home[e.second] && (wl == 0xffffffffu || load[v] < load[wl])) wl = v;
            }
            uint32_t w = wg;
            if (wl != 0xffffffffu && std::max(load[wl], e.first + hop_local) <= std::max(load[wg], e.first + hop_global) + 4 * hop_global) w = wl;
            load[w] = std::max(load[w], e.first) + instr_cost(sc.program[e.second]);
            tl[w].push_back(e.second);
        }
    }
    // ---- dependency records ----
    std::vector<uint8_t> publish(n, 0);  // bit 0: some other CTA waits for it, bit 1: another warp of its own CTA does
    auto make_deps = [&](const std::vector<uint32_t>& stream, uint32_t self_warp, uint32_t self_rank, std::vector<DepRec>& out) {
        std::vector<int64_t> seen(twc, -1);  // highest count of warp w this stream has already waited for
        std::vector<std::pair<uint32_t, uint32_t>> need;
        for (uint32_t k : stream) {
            need.clear();
            for (uint32_t q = sc.pred_off[k]; q < sc.pred_off[k + 1]; q++) {
                uint32_t pp = sc.preds[q], w = warp_of[pp];
                if (w == 0xffffffffu) throw std::logic_error("operand produced by a deferred instruction");
                if (w == self_warp) continue;  // program order on the same warp
                if ((int64_t)seq_of[pp] <= seen[w]) continue;
                bool found = false;
                for (auto& e : need)
                    if (e.first == w) {
                        e.second = std::max(e.second, seq_of[pp]);
                        found = true;
                    }
                if (!found) need.push_back({w, seq_of[pp]});
            }
            DepRec r = {};
            r.n = (uint32_t)need.size();
            if (need.size() > 0xffff) throw std::logic_error("too many dependencies");
            std::vector<uint32_t> packed;
            for (auto& e : need) {
                seen[e.first] = e.second;
                publish[cs[e.first][e.second]] |= (e.first % G == self_rank) ? 2 : 1;
                packed.push_back((e.first << DEP_SEQ_BITS) | e.second);
            }
            if (packed.size() <= 3) {
                for (size_t j = 0; j < packed.size(); j++) r.d[j] = packed[j];
            } else {
                r.d[0] = packed[0];
                r.d[1] = packed[1];
                r.d[2] = (uint32_t)ts.extra.size();
                ts.extra.insert(ts.extra.end(), packed.begin() + 2, packed.end());
            }
            out.push_back(r);
        }
    };
    ts.crit_off.push_back(0);
    for (uint32_t w = 0; w < twc; w++) {
        make_deps(cs[w], w, w % G, ts.crit_dep);
        for (uint32_t k : cs[w]) ts.crit.push_back(sc.program[k]);
        ts.crit_off.push_back((uint32_t)ts.crit.size());
    }
    ts.tail_off.push_back(0);
    for (uint32_t w = 0; w < twt; w++) {
        make_deps(tl[w], 0xfffffffeu, lay.tail_rank0 + w % lay.cta_tail, ts.tail_dep);
        for (uint32_t k : tl[w]) ts.tail.push_back(sc.program[k]);
        ts.tail_off.push_back((uint32_t)ts.tail.size());
    }
    // publish bits (known only after every stream's dependencies have been computed)
    {
        size_t o = 0;
        for (uint32_t w = 0; w < twc; w++)
            for (uint32_t k : cs[w]) ts.crit_dep[o++].n |= (uint32_t)publish[k] << 16;
    }
    if (ts.extra.empty()) ts.extra.push_back(0);
    return ts;
}

// Host model of the dataflow execution (tests / tooling): runs the streams round-robin, one
// instruction per warp per sweep, starting an instruction only when its dependency records are
// satisfied by the progress counters. Returns the instructions in the order they started; throws if
// the streams deadlock. Executing that order sequentially must reproduce the program's records.
inline std::vector<Instr> simulate_team_order(const TeamStreams& ts) {
    std::vector<uint32_t> progress(ts.twc, 0);  // published counts
    std::vector<uint32_t> cpos(ts.twc, 0), tpos(ts.twt, 0);
    std::vector<Instr> order;
    order.reserve(ts.crit.size() + ts.tail.size());
    auto ok = [&](const DepRec& r) {
        uint32_t nd = r.n & 0xffff;
        for (uint32_t j = 0; j < nd; j++) {
            uint32_t d = nd <= 3 ? r.d[j] : (j < 2 ? r.d[j] : ts.extra[r.d[2] + j - 2]);
            if (progress[d >> DEP_SEQ_BITS] <= (d & ((1u << DEP_SEQ_BITS) - 1))) return false;
        }
        return true;
    };
    size_t total = ts.crit.size() + ts.tail.size();
    while (order.size() < total) {
        bool moved = false;
        for (uint32_t w = 0; w < ts.twc; w++) {
            uint32_t k = ts.crit_off[w] + cpos[w];
            if (k >= ts.crit_off[w + 1] || !ok(ts.crit_dep[k])) continue;
            order.push_back(ts.crit[k]);
            cpos[w]++;
            if (ts.crit_dep[k].n & (3u << 16)) progress[w] = cpos[w];
            moved = true;
        }
        for (uint32_t w = 0; w < ts.twt; w++) {
            uint32_t k = ts.tail_off[w] + tpos[w];
            if (k >= ts.tail_off[w + 1] || !ok(ts.tail_dep[k])) continue;
            order.push_back(ts.tail[k]);
            tpos[w]++;
            moved = true;
        }
        if (!moved) throw std::logic_error("team streams deadlock");
    }
    return order;
}

}  // namespace h2e
