// Record layouts of a traced shape (host side, static).
//
// The witness VM produces one value per advice cell ("slot") per instance. Three layouts of those values
// exist, all tile-interleaved over 32 instances (tile t = instances 32t .. 32t+31, lane = instance % 32):
//
//   WIDE     vals[tile][slot][lane][8 words]                    every cell a 32-byte canonical Fr
//   COMPACT  rec[tile][32 * off_c(slot) + lane * w(slot) + k]   every cell at its static width class w in {1, 4, 8}
//                                                               words (range chunks / bits; limbs; field elements)
//   UNIQUE   rec[tile][32 * off_u(slot) + lane * w(slot) + k]   COMPACT without the cells that are copies: a cell
//                                                               that the reference ties to an older cell with a
//                                                               permutation pair (Records::permutations,
//                                                               context.rs:648-658) always holds that cell's value,
//                                                               so only the root of every copy class is stored
//   PRIMARY  rec[tile][32 * off_p(slot) + lane * w(slot) + k]   UNIQUE without the range-chip chunk cells: the 18-bit
//                                                               chunks of a limb's range rows (assign_nonleading_limb /
//                                                               assign_*_leading_limb / assign_common, context.rs:835-997)
//                                                               are bit fields of the limb's accumulator cell, which is
//                                                               stored; chunk j = (acc >> 18 j) & (2^18 - 1)
//
// Width classes are a property of the macro-op code (which store a call site uses: OutT::c1 / c4 / c8 in
// vm_ops.cuh); `op_widths` restates them per opcode. The restatement is checked against the device code by
// h2e_compact_prepare (a build of the VM whose stores record their width) and by the host emulator's probe
// mode in the CPU tests.
#pragma once
#include <numeric>

#include "tracer.h"

namespace h2e {

enum RecFormat : int { REC_WIDE = 0, REC_COMPACT = 1, REC_UNIQUE = 2, REC_PRIMARY = 3, REC_FORMATS = 4 };

// A derived cell is a bit field of another cell of the same macro-op: value = (cell[slot + rel] >> shift) & (2^18 - 1);
// shift == DER_ZERO: the cell is the constant 0 (zero padding of a leading limb's range row, context.rs:987).
static const uint8_t DER_ZERO = 255;
struct Derived {
    int8_t rel;     // 0 = not derived
    uint8_t shift;
};

struct WidthSink {
    std::vector<uint8_t>& w;
    std::vector<Derived>* der = nullptr;  // optional: derivation of every cell, parallel to `w`
    unsigned dec = 4;                     // chunks a leading limb decomposes (FT::WDEC == FT::DDEC == limbs, vm_ops.cuh)
    void put(unsigned n, uint8_t width) {
        w.insert(w.end(), n, width);
        if (der) der->insert(der->end(), n, Derived{0, 0});
    }
    void chunk(int rel, unsigned shift) {  // a 1-word cell that is a bit field of the cell `rel` slots further on
        w.push_back(1);
        if (der) der->push_back(Derived{(int8_t)rel, (uint8_t)shift});
    }
    void c1(unsigned n = 1) { put(n, 1); }
    void c4(unsigned n = 1) { put(n, 4); }
    void c8(unsigned n = 1) { put(n, 8); }
    void limb3() {                     // emit_limb3: six chunks, then the limb
        for (unsigned j = 0; j < 6; j++) chunk(6 - (int)j, 18 * j);
        c4();
    }
    void lead2() {                     // emit_lead2: `dec` chunks + zero padding, then the limb
        for (unsigned j = 0; j < 4; j++) chunk(4 - (int)j, j < dec ? 18 * j : DER_ZERO);
        c4();
    }
    void common() {                    // emit_common: the value twice
        c1();
        chunk(-1, 0);
    }
    void is_zero_rows() { c8(); c1(); c8(2); c1(); }  // emit_is_zero_rows
    void assign_int(unsigned L) {      // emit_assign_int / emit_assign_int_known
        for (unsigned i = 0; i + 1 < L; i++) limb3();
        lead2();
        c4(L);
        c8();
    }
    void mul_constraints(unsigned L, unsigned M) {  // emit_mul_constraints
        for (unsigned pos = 0; pos < M; pos++) {
            unsigned hi = pos + 1 < L ? pos + 1 : L, lo = pos >= L - 1 ? pos - (L - 1) : 0, n = hi - lo;
            for (unsigned i = 0; i < n; i++) {
                c4(3);
                if (n > 1) c8();
            }
            c8();
        }
        for (unsigned pos = 0; pos < M; pos++) {
            c8();
            if (pos < L) c4();
            if (pos > 0) { c1(); c4(); }
            c8();
            common();
            limb3();
            c1(); c4(); c8();
        }
        c8(4);
    }
};

// width class of every cell `in` writes, in slot order (appended to `w`)
inline void op_widths(const Instr& in, std::vector<uint8_t>& w, std::vector<Derived>* der = nullptr) {
    WidthSink o{w, der};
    const FieldInfo* fi = in.field < F_COUNT ? &field_info((Field)in.field) : nullptr;
    o.dec = fi ? fi->limbs : 4;
    const unsigned L = fi ? fi->limbs : 0, M = fi ? fi->mul_check_limbs : 0, R = fi ? fi->reduce_check_limbs : 0,
                   P = fi ? fi->pure_w_check_limbs : 0;
    switch (in.op) {
        case OP_NOP: break;
        case OP_LOAD_INT:
        case OP_ASSIGN_INT_CONST: o.c4(L); o.c8(); break;
        case OP_ASSIGN_W: o.assign_int(L); break;
        case OP_INT_ADD:
        case OP_INT_SUB: o.c4(3 * L); o.c4(L); o.c8(); break;
        case OP_INT_NEG:
        case OP_MUL_SMALL: o.c4(2 * L); o.c4(L); o.c8(); break;
        case OP_REDUCE:
            o.assign_int(L);
            o.common(); o.c1(); o.c8(2);
            for (unsigned i = 0; i < R; i++) { o.limb3(); o.c1(); o.c4(4); }
            break;
        case OP_INT_MUL:
        case OP_DIV_CORE:
            o.assign_int(L);
            o.assign_int(L);
            o.mul_constraints(L, M);
            break;
        case OP_IS_INT_ZERO:
            o.c4(L); o.c8(); o.is_zero_rows();
            o.c8(2); o.is_zero_rows();
            for (unsigned i = 0; i < P; i++) { o.c4(); o.c8(); o.is_zero_rows(); o.c1(3); }
            o.c1(3);
            break;
        case OP_MASK_INT: o.c8(3 * (L + 1)); break;
        case OP_BISEC_INT: o.c8(5 * (L + 1)); break;
        case OP_SUM_ASSERT_ZERO: o.c4(L); o.c8(2); break;
        case OP_ASSIGN:
        case OP_ASSIGN_CONST:
        case OP_ASSERT_CONST: o.c8(); break;
        case OP_ASSIGN_BIT:
        case OP_ASSERT_EQUAL: o.c8(2); break;
        case OP_LINSUM: o.c8(in.a[0] + 1); break;
        case OP_MUL:
        case OP_BOOL: o.c8(3); break;
        case OP_BISEC: o.c8(5); break;
        case OP_IS_ZERO: o.is_zero_rows(); break;
        case OP_DECOMPOSE_NATIVE:
            for (unsigned i = 0; i < in.a[1]; i++) { o.c1(4); o.c8(); o.c1(2); o.c8(); }
            o.c8();
            break;
        case OP_DECOMPOSE_LIMB:
            for (unsigned i = 0; i < in.a[1]; i++) { o.c1(2); o.c8(); o.c1(); o.c8(); }
            o.c8();
            break;
        case OP_CACHE_INT: o.c8(L + 1); break;
        case OP_SELECT_INT: o.c8(2 * (L + 1)); break;
        case OP_BOOLV: o.c1(3 * in.a[1]); break;  // AssignedCondition cells: bits by construction, stored as one word
        case OP_CHIV: o.c1(6 * in.a[1]); break;
        default: throw std::logic_error("op_widths: opcode " + std::to_string(in.op) + " is not a traced macro-op");
    }
}

// Positions in Instr::a that hold slot numbers (Instr::out is one as well), for every opcode including the
// scheduler's split ops. The DEVICE copy of a program carries references instead (off(slot) << 2 | width class).
inline void instr_slot_fields(const Instr& in, std::vector<uint8_t>& idx) {
    idx.clear();
    const unsigned L = in.field < F_COUNT ? field_info((Field)in.field).limbs : 0;
    auto range = [&](unsigned from, unsigned to) {
        for (unsigned i = from; i < to; i++) idx.push_back((uint8_t)i);
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
        case OP_SUM_ASSERT_ZERO:
        case OP_REDUCE_HEAD: range(0, L); break;
        case OP_DIV_INV:
            for (unsigned j = 0; j < std::max(1u, in.flags & 3u); j++) range(j * (L + 1), j * (L + 1) + L);
            break;
        case OP_REDUCE:
        case OP_REDUCE_TAIL:
        case OP_IS_INT_ZERO:
        case OP_CACHE_INT: range(0, L + 1); break;
        case OP_IS_INT_ZERO_HEAD:
            range(0, L + 1);
            idx.push_back(13);  // the condition cell
            break;
        case OP_IS_INT_ZERO_TAIL:
            for (unsigned j = 0; j < (in.flags & 3u); j++) {
                range(j * (L + 1), (j + 1) * (L + 1));
                if (j) idx.push_back((uint8_t)(11 + j));  // first slot of block j
            }
            break;
        case OP_INT_MUL:
        case OP_INT_MUL_HEAD:
        case OP_INT_MUL_TAIL:
        case OP_DIV_CORE:
        case OP_DIV_CORE_S:
        case OP_DIV_HEAD_S:
        case OP_DIV_TAIL: range(0, 2 * L + 2); break;  // (a[2L+2] of the _S ops is a scratch entry, not a slot)
        case OP_MASK_INT: range(0, L + 2); break;
        case OP_BISEC_INT: range(0, 2 * L + 3); break;
        case OP_LINSUM:
            for (unsigned i = 0; i < in.a[0]; i++) idx.push_back((uint8_t)(2 + 2 * i));
            break;
        case OP_MUL:
        case OP_BOOL:
        case OP_ASSERT_EQUAL: range(0, 2); break;
        case OP_BISEC: range(0, 3); break;
        case OP_IS_ZERO:
        case OP_ASSERT_CONST:
        case OP_DECOMPOSE_NATIVE:
        case OP_DECOMPOSE_LIMB:
        case OP_SELECT_INT: range(0, 1); break;  // (the candidates of OP_SELECT_INT are in Shape::tables)
        default: break;  // input-only ops
    }
}
inline uint32_t slot_ref(uint32_t slot, const uint32_t* off, const uint8_t* width) {
    return (off[slot] << 2) | (width[slot] == 8 ? 2u : (width[slot] == 4 ? 1u : 0u));
}
// slot numbers -> references, in place
inline void translate_program(Instr* prog, size_t n, const uint32_t* off, const uint8_t* width) {
    std::vector<uint8_t> idx;
    for (size_t i = 0; i < n; i++) {
        Instr& in = prog[i];
        instr_slot_fields(in, idx);
        if (in.op != OP_NOP) in.out = slot_ref(in.out, off, width);
        for (uint8_t k : idx) in.a[k] = slot_ref(in.a[k], off, width);
    }
}

struct Layout {
    std::vector<uint8_t> width;         // [n_slots] 1, 4 or 8 words
    std::vector<uint32_t> root;         // [n_slots] the oldest slot holding the same value by a permutation pair (itself if none)
    std::vector<uint32_t> off_compact;  // [n_slots + 1] words per lane before slot s (COMPACT)
    std::vector<uint32_t> off_unique;   // [n_slots + 1] same for UNIQUE; copies take no room (off_unique[s + 1] == off_unique[s])
    std::vector<uint32_t> unique_slots; // slots stored in UNIQUE, ascending
    std::vector<uint32_t> der_src;      // [n_slots] NONE, or the slot this (root) cell is a bit field of
    std::vector<uint8_t> der_shift;     // [n_slots] shift of that bit field (18 bits wide), DER_ZERO = constant 0
    std::vector<uint32_t> off_primary;  // [n_slots + 1] same for PRIMARY; copies and derived cells take no room
    std::vector<uint32_t> primary_slots;  // slots stored in PRIMARY, ascending
    uint64_t n_copies = 0, n_derived = 0;

    const std::vector<uint32_t>& off(int format) const { return format == REC_PRIMARY ? off_primary : (format == REC_UNIQUE ? off_unique : off_compact); }
    const std::vector<uint32_t>& stored_slots(int format) const { return format == REC_PRIMARY ? primary_slots : unique_slots; }
    // words per lane of one tile in `format`
    uint64_t words_per_lane(int format, size_t n_slots) const {
        return format == REC_WIDE ? (uint64_t)n_slots * 8 : off(format).back();
    }
};

inline Layout build_layout(const Shape& sh) {
    Layout lay;
    const size_t n = sh.slot_cell.size();
    lay.width.reserve(n);
    std::vector<Derived> der;
    der.reserve(n);
    for (size_t i = 0; i < sh.program.size(); i++) {
        const Instr& in = sh.program[i];
        if ((size_t)in.out != lay.width.size()) throw std::logic_error("layout: macro-op " + std::to_string(i) + " does not start where the previous one ended");
        op_widths(in, lay.width, &der);
    }
    if (lay.width.size() != n) throw std::logic_error("layout: width table covers " + std::to_string(lay.width.size()) + " of " + std::to_string(n) + " slots");
    // copy classes from the permutation pairs: (older cell, new cell), both advice cells with a slot
    std::vector<std::vector<uint32_t>> cell_slot(3);  // region -> col * height + row -> slot
    for (int r = 0; r < 3; r++) cell_slot[r].assign((size_t)ADV_COLS[r] * std::max<size_t>(sh.height[r], 1), NONE);
    auto key = [&](const Cell& c) { return (size_t)c.col * std::max<size_t>(sh.height[c.region], 1) + c.row; };
    for (size_t s = 0; s < n; s++) cell_slot[sh.slot_cell[s].region][key(sh.slot_cell[s])] = (uint32_t)s;
    lay.root.resize(n);
    std::iota(lay.root.begin(), lay.root.end(), 0u);
    auto find = [&](uint32_t s) {
        while (lay.root[s] != s) {
            lay.root[s] = lay.root[lay.root[s]];
            s = lay.root[s];
        }
        return s;
    };
    for (const auto& p : sh.perms) {
        if (p[0].region > 2 || p[1].region > 2 || p[0].row >= sh.height[p[0].region] || p[1].row >= sh.height[p[1].region]) continue;
        uint32_t a = cell_slot[p[0].region][key(p[0])], b = cell_slot[p[1].region][key(p[1])];
        if (a == NONE || b == NONE) continue;
        a = find(a);
        b = find(b);
        if (a == b) continue;
        if (a < b) lay.root[b] = a;
        else lay.root[a] = b;
    }
    lay.off_compact.assign(n + 1, 0);
    lay.off_unique.assign(n + 1, 0);
    lay.off_primary.assign(n + 1, 0);
    lay.der_src.assign(n, NONE);
    lay.der_shift.assign(n, 0);
    for (size_t s = 0; s < n; s++) lay.root[s] = find((uint32_t)s);
    for (size_t s = 0; s < n; s++) {
        const bool copy = lay.root[s] != s;
        lay.n_copies += copy;
        if (!copy) lay.unique_slots.push_back((uint32_t)s);
        // a root that is a bit field of a cell whose own root is stored (never itself derived: sources are limb
        // accumulators and the first cell of assign_common) is rebuilt by the consumer
        bool derived = false;
        if (!copy && der[s].rel != 0) {
            const uint32_t src = lay.root[(size_t)((int64_t)s + der[s].rel)];
            if (der[s].shift == DER_ZERO || der[src].rel == 0) {
                derived = true;
                lay.der_src[s] = src;
                lay.der_shift[s] = der[s].shift;
                lay.n_derived++;
            }
        }
        if (!copy && !derived) lay.primary_slots.push_back((uint32_t)s);
        lay.off_compact[s + 1] = lay.off_compact[s] + lay.width[s];
        lay.off_unique[s + 1] = lay.off_unique[s] + (copy ? 0 : lay.width[s]);
        lay.off_primary[s + 1] = lay.off_primary[s] + (copy || derived ? 0 : lay.width[s]);
        if ((uint64_t)lay.off_compact[s] + lay.width[s] >= (1ull << 30)) throw std::logic_error("layout: more than 2^30 words per lane");
    }
    return lay;
}

}  // namespace h2e
