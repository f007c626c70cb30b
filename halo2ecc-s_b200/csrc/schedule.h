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
        case OP_REDUCE_HEAD: range(0, L); break;
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
        default: break;  // input-only ops
    }
}

// rough single-warp latency of a macro-op in cycles (measured on B200, team mode), for load balancing
inline uint32_t instr_cost(const Instr& in) {
    switch (in.op) {
        case OP_IS_INT_ZERO: return 90000;
        case OP_DIV_CORE: return 120000;
        case OP_IS_ZERO: return 80000;
        case OP_DECOMPOSE_NATIVE: return 30000;
        case OP_DECOMPOSE_LIMB: return 15000;
        case OP_INT_MUL: return 9000;
        case OP_INT_MUL_HEAD: return 4500;
        case OP_INT_MUL_TAIL: return (in.flags & 2) ? 5000 : 3000;
        case OP_REDUCE: return 5000;
        case OP_REDUCE_HEAD: return 3000;
        case OP_REDUCE_TAIL: return 3000;
        case OP_ASSIGN_W: return 3000;
        case OP_LINSUM: return 3000;
        default: return 1500;
    }
}

struct Schedule {
    std::vector<Instr> program;        // instructions sorted by (level, opcode)
    std::vector<uint32_t> level_start; // level l = program[level_start[l] .. level_start[l+1])
    std::vector<uint32_t> level_mid;   // [level_start[l], level_mid[l]) critical ops, [level_mid[l], level_start[l+1]) deferred (TAIL) ops
    uint32_t max_width = 0;
};

inline Schedule levelise(const Shape& sh, bool split_int_mul = true, bool align_heavy = true) {
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
        } else if (split_int_mul && sh.program[i].op == OP_REDUCE) {
            Instr h = sh.program[i], t = sh.program[i];
            h.op = OP_REDUCE_HEAD;
            t.op = OP_REDUCE_TAIL;
            p.push_back(h);
            block_end.push_back(h.out);
            p.push_back(t);
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
        if (p[i].op == OP_INT_MUL_HEAD || p[i].op == OP_REDUCE_HEAD)
            for (uint32_t s : head_cells(p[i])) producer[s] = (uint32_t)i;
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
            if (s >= p[i].out && p[i].op != OP_INT_MUL_TAIL && p[i].op != OP_REDUCE_TAIL) throw std::logic_error("instruction reads a slot it has not seen produced");
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
    auto weight = [](const Instr& in) -> int { return (int)instr_cost(in); };
    sc.level_mid.resize(n_levels);
    for (uint32_t l = 0; l < n_levels; l++) {
        auto b = sc.program.begin() + sc.level_start[l], e = sc.program.begin() + sc.level_start[l + 1];
        // deferred ops (nothing depends on their output) go last; the rest heavy-first
        auto mid = std::stable_partition(b, e, [](const Instr& in) { return (in.flags & 0x80) == 0; });
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

// Per-warp instruction streams of team mode. A tile is evaluated by `twc` critical team warps, which
// walk the levels in lock step (a barrier between levels), and `twt` tail team warps, which execute
// the deferred instructions in level order, each as soon as the level it depends on has completed.
// Every team warp reads its own contiguous stream, so the next instruction is always prefetchable.
struct TeamStreams {
    uint32_t n_levels = 0, twc = 0, twt = 0;
    std::vector<Instr> crit;           // streams of the critical team warps, back to back
    std::vector<uint32_t> crit_off;    // [twc + 1] stream bounds
    std::vector<uint16_t> crit_cnt;    // [twc][n_levels] instructions of warp w in level l
    std::vector<Instr> tail;           // streams of the tail team warps, back to back
    std::vector<uint32_t> tail_off;    // [twt + 1]
    std::vector<uint32_t> tail_ready;  // parallel to `tail`: number of completed levels the instruction needs
};

inline TeamStreams build_team_streams(const Schedule& sc, uint32_t twc, uint32_t twt) {
    TeamStreams ts;
    ts.n_levels = (uint32_t)sc.level_start.size() - 1;
    ts.twc = twc;
    ts.twt = twt;
    std::vector<std::vector<Instr>> cs(twc), tl(twt);
    std::vector<std::vector<uint32_t>> tr(twt);
    ts.crit_cnt.assign((size_t)twc * ts.n_levels, 0);
    size_t t = 0;
    std::vector<uint64_t> load(twc, 0);
    for (uint32_t l = 0; l < ts.n_levels; l++) {
        // longest-processing-time-first: instructions come heaviest first; each goes to the least loaded warp
        std::fill(load.begin(), load.end(), 0);
        for (uint32_t i = sc.level_start[l]; i < sc.level_mid[l]; i++) {
            uint32_t w = 0;
            for (uint32_t k = 1; k < twc; k++)
                if (load[k] < load[w]) w = k;
            load[w] += instr_cost(sc.program[i]);
            cs[w].push_back(sc.program[i]);
            if (++ts.crit_cnt[(size_t)w * ts.n_levels + l] == 0) throw std::logic_error("level too wide for the per-level counter");
        }
        for (uint32_t i = sc.level_mid[l]; i < sc.level_start[l + 1]; i++, t++) {
            tl[t % twt].push_back(sc.program[i]);
            tr[t % twt].push_back(l);
        }
    }
    ts.crit_off.push_back(0);
    for (uint32_t w = 0; w < twc; w++) {
        ts.crit.insert(ts.crit.end(), cs[w].begin(), cs[w].end());
        ts.crit_off.push_back((uint32_t)ts.crit.size());
    }
    ts.tail_off.push_back(0);
    for (uint32_t w = 0; w < twt; w++) {
        ts.tail.insert(ts.tail.end(), tl[w].begin(), tl[w].end());
        ts.tail_ready.insert(ts.tail_ready.end(), tr[w].begin(), tr[w].end());
        ts.tail_off.push_back((uint32_t)ts.tail.size());
    }
    return ts;
}

}  // namespace h2e
