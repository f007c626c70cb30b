// Host-side mirror of the reference's chip API for the hot path, in *symbolic* form.
//
// The reference computes witness values on the CPU while it lays rows into `Records`
// (src/context.rs:241-301). Here the same calls (same names, argument meaning and `times`
// bookkeeping) are made ONCE per circuit shape: they lay out rows, fixed cells and permutations
// exactly as the reference does, assign every advice cell a dense slot number, and emit a
// straight-line program of macro-ops (h2e_program.h). The CUDA VM then fills the advice values for
// a whole batch of instances. An AssignedValue therefore carries a slot instead of a value.
//
//   Context / Records writers   src/context.rs:40-46, 590-997
//   BaseChipOps                 src/circuit/base_chip.rs:81-605
//   RangeChipOps                src/circuit/range_chip.rs:262-348
//   SelectChipOps               src/circuit/select_chip.rs:100-162
//   IntegerChipOps              src/circuit/integer_chip.rs:15-686
#pragma once
#include <array>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "fieldinfo.h"

namespace h2e {

enum Chip : uint8_t { BaseChip = 0, RangeChip = 1, SelectChip = 2 };  // assign.rs:5-10
static const int VAR_COLUMNS = 5, MUL_COLUMNS = 2, FIXED_COLUMNS = 9;   // base_chip.rs:14-16
static const int ADV_COLS[3] = {5, 3, 2};
static const int FIX_COLS[3] = {9, 2, 2};
enum RangeAdvCol { ValueAccCol = 0, TaggedRangeCol = 1, CommonRangeCol = 2 };
enum RangeFixCol { AccLinesCol = 0, TagCol = 1 };
enum SelAdvCol { SelValueCol = 0, SelSelectCol = 1 };
enum SelFixCol { EncodeCol = 0, IsLookupCol = 1 };
static const size_t MSM_PREFIX_OFFSET = 1u << 20;  // ecc_chip.rs:20-21
static const size_t MSM_LIMIT = (1u << 8) * MSM_PREFIX_OFFSET;

struct Cell {
    uint8_t region;
    uint8_t col;
    uint32_t row;
};

struct AssignedValue {  // assign.rs:25-29 (val lives on the device, in `slot`)
    Cell cell;
    uint32_t slot;
};
struct AssignedCondition {
    AssignedValue v;
};
struct AssignedInteger {  // assign.rs:31-37
    std::vector<AssignedValue> limbs_le;
    AssignedValue native;
    uint64_t times = 0;
};

// A fixed cell is either a shape-level constant (index into Shape::consts) or, for
// assign_constant of a per-instance value, a copy of an advice slot (FIX_FROM_SLOT | slot).
static const uint32_t FIX_FROM_SLOT = 0x80000000u;
struct FixEntry {
    uint32_t row;
    uint8_t region;
    uint8_t col;
    uint32_t cidx;
};

typedef std::array<uint32_t, 8> Const256;

// Everything about a circuit that is independent of witness values + the value program.
struct Shape {
    size_t height[3] = {0, 0, 0};           // base_height, range_height, select_height
    size_t offset[3] = {0, 0, 0};           // final base/range/select offsets
    std::vector<Cell> slot_cell;            // slot -> advice cell
    std::vector<FixEntry> fixed;            // fixed cells in assignment order
    std::vector<std::array<Cell, 2>> perms; // Records::permutations
    std::vector<Instr> program;
    std::vector<Const256> consts;           // constant pool (device-visible)
    uint32_t n_inputs = 0;                  // per-instance input cells (32 bytes each)
    std::vector<uint32_t> tables;           // slot tables for OP_SELECT_INT (candidate cells)
    std::unordered_map<std::string, uint32_t> const_index;

    uint32_t add_const(const Big& v) {
        Const256 c;
        v.to_words(c.data(), 8);
        std::string key((const char*)c.data(), 32);
        auto it = const_index.find(key);
        if (it != const_index.end()) return it->second;
        uint32_t idx = (uint32_t)consts.size();
        consts.push_back(c);
        const_index[key] = idx;
        return idx;
    }
    // wide (up to 512-bit) constant: two consecutive pool entries (lo, hi); not de-duplicated
    uint32_t add_const_wide(const Big& v) {
        uint32_t w[16];
        v.to_words(w, 16);
        Const256 lo, hi;
        memcpy(lo.data(), w, 32);
        memcpy(hi.data(), w + 8, 32);
        uint32_t idx = (uint32_t)consts.size();
        consts.push_back(lo);
        consts.push_back(hi);
        return idx;
    }
};

// Field element of N on the host (setup-time constants only).
struct NConst {
    Big v;
    static NConst from(uint64_t x) { return NConst{Big(x)}; }
};
inline Big n_neg(const Big& a) { return a.is_zero() ? a : native_modulus() - a % native_modulus(); }

// ValueSchema (assign.rs:123-146). Unassigned(N) carries no value here: it is "a value the
// macro-op computes"; only whether a permutation is recorded matters for the shape.
struct ValueSchema {
    bool assigned;
    AssignedValue av;
    ValueSchema() : assigned(false), av{{0, 0, 0}, NONE} {}
    ValueSchema(const AssignedValue& a) : assigned(true), av(a) {}
    ValueSchema(const AssignedValue* a) : assigned(true), av(*a) {}
};
struct Coeff {  // fixed-cell value for a row: constant, or "copy of the advice cell of this row"
    bool from_adv = false;
    Big v;
    Coeff() {}
    Coeff(const Big& b) : v(b) {}
    Coeff(uint64_t x) : v(x) {}
};
typedef std::pair<ValueSchema, Coeff> Pair;

class Context {  // context.rs:40-46 + Records
   public:
    Shape shape;
    size_t base_offset = 0, range_offset = 0, select_offset = 0;
    int depth = 0;  // >0 while inside a macro-op: nested ops lay out rows but emit no instruction

    const Big R = native_modulus();
    Big ONE = Big(1), ZERO = Big(0), NEG_ONE = native_modulus() - Big(1);

    // ---- slots ----
    uint32_t next_slot() const { return (uint32_t)shape.slot_cell.size(); }
    uint32_t new_slot(uint8_t region, uint8_t col, size_t row) {
        shape.slot_cell.push_back(Cell{region, col, (uint32_t)row});
        return (uint32_t)shape.slot_cell.size() - 1;
    }
    void fix(uint8_t region, uint8_t col, size_t row, const Big& v) {
        shape.fixed.push_back(FixEntry{(uint32_t)row, region, col, shape.add_const(v)});
    }
    void fix_from_slot(uint8_t region, uint8_t col, size_t row, uint32_t slot) {
        shape.fixed.push_back(FixEntry{(uint32_t)row, region, col, FIX_FROM_SLOT | slot});
    }
    void permute(const Cell& a, const Cell& b) { shape.perms.push_back({a, b}); }

    void emit(const Instr& in) { shape.program.push_back(in); }

    struct Macro {  // RAII: emit `in` if we are at top level, then suppress nested emission
        Context& c;
        uint32_t out;
        Macro(Context& ctx, Instr in) : c(ctx), out(ctx.next_slot()) {
            if (c.depth == 0) {
                in.out = out;
                c.emit(in);
            }
            c.depth++;
        }
        ~Macro() { c.depth--; }
        void expect_cells(uint32_t n) const {
            if (c.next_slot() - out != n) throw std::logic_error("macro-op cell count mismatch");
        }
    };
    static Instr mk(Op op, uint8_t field = 0) {
        Instr in;
        memset(&in, 0, sizeof(in));
        in.op = op;
        in.field = field;
        return in;
    }

    // ---- Records::one_line / one_line_with_last (context.rs:634-714) ----
    std::vector<AssignedValue> one_line(const std::vector<Pair>& pairs, const Coeff* constant, const std::vector<Big>& mul,
                                        const Big* next, const Pair* last = nullptr, AssignedValue* last_out = nullptr) {
        if (depth == 0) throw std::logic_error("raw one_line outside a macro-op is not executable on the device");
        size_t offset = base_offset;
        if (offset >= shape.height[0]) shape.height[0] = offset + 1;
        std::vector<AssignedValue> res;
        auto put = [&](int col, const Pair& p) {
            uint32_t s = new_slot(BaseChip, col, offset);
            Cell nc{BaseChip, (uint8_t)col, (uint32_t)offset};
            if (p.first.assigned) permute(p.first.av.cell, nc);
            if (p.second.from_adv)
                fix_from_slot(BaseChip, col, offset, s);
            else
                fix(BaseChip, col, offset, p.second.v);
            return AssignedValue{nc, s};
        };
        for (size_t i = 0; i < pairs.size(); i++) res.push_back(put((int)i, pairs[i]));
        for (size_t i = 0; i < mul.size(); i++) fix(BaseChip, VAR_COLUMNS + i, offset, mul[i]);
        if (next) fix(BaseChip, VAR_COLUMNS + MUL_COLUMNS, offset, *next);
        AssignedValue lastv{{0, 0, 0}, NONE};
        // NOTE: the reference writes the constant before the `last` cell (one_line runs first,
        // context.rs:695-713); order of fixed cells does not matter for the records.
        if (constant) {
            if (constant->from_adv)
                fix_from_slot(BaseChip, VAR_COLUMNS + MUL_COLUMNS + 1, offset, res.at(0).slot);
            else
                fix(BaseChip, VAR_COLUMNS + MUL_COLUMNS + 1, offset, constant->v);
        }
        if (last) lastv = put(VAR_COLUMNS - 1, *last);
        if (last_out) *last_out = lastv;
        base_offset += 1;
        return res;
    }
    AssignedValue one_line_with_last(const std::vector<Pair>& pairs, const Pair& last, const Coeff* constant, const std::vector<Big>& mul,
                                     const Big* next, std::vector<AssignedValue>* cells = nullptr) {
        AssignedValue l;
        auto r = one_line(pairs, constant, mul, next, &last, &l);
        if (cells) *cells = r;
        return l;
    }

    // ---- BaseChipOps (base_chip.rs) ----
    typedef std::pair<const AssignedValue*, Big> Elem;

    // base_chip.rs:110-132
    AssignedValue sum_with_constant_in_one_line(const std::vector<Elem>& elems, const Big* constant) {
        if (elems.size() >= (size_t)VAR_COLUMNS) throw std::logic_error("too many elems");
        Instr in = mk(OP_LINSUM);
        in.a[0] = (uint32_t)elems.size();
        in.a[1] = constant ? shape.add_const(*constant) : NONE;
        for (size_t i = 0; i < elems.size(); i++) {
            in.a[2 + 2 * i] = elems[i].first->slot;
            in.a[3 + 2 * i] = shape.add_const(elems[i].second);
        }
        Macro m(*this, in);
        std::vector<Pair> pairs;
        for (auto& e : elems) pairs.push_back(Pair(ValueSchema(e.first), Coeff(e.second)));
        Coeff k;
        if (constant) k = Coeff(*constant);
        return one_line_with_last(pairs, Pair(ValueSchema(), Coeff(NEG_ONE)), constant ? &k : nullptr, {}, nullptr);
    }
    // base_chip.rs:134-153
    AssignedValue sum_with_constant(const std::vector<Elem>& elems, const Big* constant) {
        size_t columns = VAR_COLUMNS;
        if (elems.size() < columns) return sum_with_constant_in_one_line(elems, constant);
        std::vector<Elem> curr(elems.begin(), elems.begin() + (columns - 1));
        AssignedValue acc = sum_with_constant_in_one_line(curr, constant);
        for (size_t p = columns - 1; p < elems.size(); p += columns - 2) {
            size_t e = std::min(p + columns - 2, elems.size());
            std::vector<Elem> chunk(elems.begin() + p, elems.begin() + e);
            AssignedValue prev = acc;
            chunk.push_back(Elem(&prev, ONE));
            acc = sum_with_constant_in_one_line(chunk, nullptr);
        }
        return acc;
    }
    AssignedValue add(const AssignedValue& a, const AssignedValue& b) { return sum_with_constant({Elem(&a, ONE), Elem(&b, ONE)}, nullptr); }
    AssignedValue add_constant(const AssignedValue& a, const Big& c) { return sum_with_constant({Elem(&a, ONE)}, &c); }
    AssignedValue sub(const AssignedValue& a, const AssignedValue& b) {
        return sum_with_constant({Elem(&a, ONE), Elem(&b, NEG_ONE)}, nullptr);
    }
    // base_chip.rs:176-193
    AssignedValue mul(const AssignedValue& a, const AssignedValue& b) {
        Instr in = mk(OP_MUL);
        in.a[0] = a.slot;
        in.a[1] = b.slot;
        Macro m(*this, in);
        return one_line_with_last({Pair(&a, ZERO), Pair(&b, ZERO)}, Pair(ValueSchema(), NEG_ONE), nullptr, {ONE}, nullptr);
    }
    // base_chip.rs:245-281 (structure only; used inside integer macro-ops)
    struct MulAddTerm {
        const AssignedValue *a, *b, *c;
        Big c_coeff;
    };
    AssignedValue mul_add_with_next_line(const std::vector<MulAddTerm>& ls) {
        if (ls.size() == 1)  // mul_add (base_chip.rs:219-243)
            return one_line_with_last({Pair(ls[0].a, ZERO), Pair(ls[0].b, ZERO), Pair(ls[0].c, ls[0].c_coeff)}, Pair(ValueSchema(), NEG_ONE),
                                      nullptr, {ONE}, nullptr);
        for (size_t i = 0; i < ls.size(); i++)
            one_line_with_last({Pair(ls[i].a, ZERO), Pair(ls[i].b, ZERO), Pair(ls[i].c, ls[i].c_coeff)},
                               Pair(ValueSchema(), i == 0 ? ZERO : ONE), nullptr, {ONE}, &NEG_ONE);
        return one_line_with_last({}, Pair(ValueSchema(), ZERO), nullptr, {}, nullptr);
    }
    // base_chip.rs:298-325
    std::pair<AssignedCondition, AssignedValue> invert(const AssignedValue& a) {
        Instr in = mk(OP_IS_ZERO);
        in.a[0] = a.slot;
        Macro m(*this, in);
        auto cells = one_line({Pair(&a, ZERO), Pair(ValueSchema(), ZERO)}, nullptr, {ONE}, nullptr);
        AssignedValue c1 = cells[1];
        std::vector<AssignedValue> r0;
        Coeff k(NEG_ONE);
        AssignedValue l = one_line_with_last({Pair(&a, ZERO), Pair(ValueSchema(), ZERO)}, Pair(&c1, ONE), &k, {ONE}, nullptr, &r0);
        return {AssignedCondition{l}, r0[1]};
    }
    AssignedCondition is_zero(const AssignedValue& a) { return invert(a).first; }
    // base_chip.rs:344-355. The device reads the value from per-instance input cell `in_cell`.
    AssignedValue assign(uint32_t in_cell) {
        Instr in = mk(OP_ASSIGN);
        in.a[0] = in_cell;
        note_input(in_cell);
        Macro m(*this, in);
        return one_line({Pair(ValueSchema(), ZERO)}, nullptr, {}, nullptr)[0];
    }
    // assign_constant of a shape-level constant
    AssignedValue assign_constant(const Big& v) {
        Instr in = mk(OP_ASSIGN_CONST);
        in.a[0] = 1;
        in.a[1] = shape.add_const(v);
        Macro m(*this, in);
        Coeff k(v);
        return one_line({Pair(ValueSchema(), NEG_ONE)}, &k, {}, nullptr)[0];
    }
    // assign_constant of a per-instance value: the fixed `constant` cell of this row differs per
    // instance and is recorded as a copy of the row's advice cell.
    AssignedValue assign_constant_input(uint32_t in_cell) {
        Instr in = mk(OP_ASSIGN_CONST);
        in.a[0] = 0;
        in.a[1] = in_cell;
        note_input(in_cell);
        Macro m(*this, in);
        Coeff k;
        k.from_adv = true;
        return one_line({Pair(ValueSchema(), NEG_ONE)}, &k, {}, nullptr)[0];
    }
    // base_chip.rs:357-367
    AssignedCondition assign_bit(uint32_t in_cell) {
        Instr in = mk(OP_ASSIGN_BIT);
        in.a[0] = in_cell;
        note_input(in_cell);
        Macro m(*this, in);
        return AssignedCondition{one_line({Pair(ValueSchema(), ONE), Pair(ValueSchema(), ZERO)}, nullptr, {NEG_ONE}, nullptr)[0]};
    }
    // row of assign_bit / assert_bit without an instruction (inside a macro-op)
    AssignedCondition assign_bit_row() {
        return AssignedCondition{one_line({Pair(ValueSchema(), ONE), Pair(ValueSchema(), ZERO)}, nullptr, {NEG_ONE}, nullptr)[0]};
    }
    // base_chip.rs:369-379
    void assert_equal(const AssignedValue& a, const AssignedValue& b) {
        Instr in = mk(OP_ASSERT_EQUAL);
        in.a[0] = a.slot;
        in.a[1] = b.slot;
        Macro m(*this, in);
        one_line({Pair(&a, NEG_ONE), Pair(&b, ONE)}, nullptr, {}, nullptr);
    }
    void assert_constant(const AssignedValue& a, const Big& b, uint32_t status_bit = 0) {
        Instr in = mk(OP_ASSERT_CONST);
        in.a[0] = a.slot;
        in.a[1] = shape.add_const(b);
        in.a[2] = status_bit;
        Macro m(*this, in);
        Coeff k(b);
        one_line({Pair(&a, NEG_ONE)}, &k, {}, nullptr);
    }
    // base_chip.rs:392-467
    AssignedCondition and_(const AssignedCondition& a, const AssignedCondition& b) { return AssignedCondition{mul(a.v, b.v)}; }
    AssignedCondition not_(const AssignedCondition& a) { return AssignedCondition{sum_with_constant({Elem(&a.v, NEG_ONE)}, &ONE)}; }
    AssignedCondition bool_op(int kind, const AssignedCondition& a, const AssignedCondition& b) {
        Instr in = mk(OP_BOOL);
        in.a[0] = a.v.slot;
        in.a[1] = b.v.slot;
        in.a[2] = kind;
        Macro m(*this, in);
        Big two(2), neg_two = n_neg(two);
        switch (kind) {
            case 1:  // or
                return AssignedCondition{
                    one_line_with_last({Pair(&a.v, ONE), Pair(&b.v, ONE)}, Pair(ValueSchema(), NEG_ONE), nullptr, {NEG_ONE}, nullptr)};
            case 2:  // xor
                return AssignedCondition{
                    one_line_with_last({Pair(&a.v, ONE), Pair(&b.v, ONE)}, Pair(ValueSchema(), NEG_ONE), nullptr, {neg_two}, nullptr)};
            case 3: {  // xnor
                Coeff k(ONE);
                return AssignedCondition{
                    one_line_with_last({Pair(&a.v, NEG_ONE), Pair(&b.v, NEG_ONE)}, Pair(ValueSchema(), NEG_ONE), &k, {two}, nullptr)};
            }
            default:  // not_and
                return AssignedCondition{
                    one_line_with_last({Pair(&a.v, ZERO), Pair(&b.v, ONE)}, Pair(ValueSchema(), NEG_ONE), nullptr, {NEG_ONE}, nullptr)};
        }
    }
    // n consecutive, mutually independent or / xor / xnor / not_and rows as ONE macro-op (the rows and their order are those
    // of n bool_op calls); operand slots go to the shape's slot tables
    std::vector<AssignedCondition> bool_vec(int kind, const std::vector<AssignedCondition>& a, const std::vector<AssignedCondition>& b) {
        if (a.size() != b.size() || a.empty()) throw std::logic_error("bool_vec: operand lists differ in length");
        if (depth != 0) {  // inside another macro-op: plain rows
            std::vector<AssignedCondition> r;
            for (size_t i = 0; i < a.size(); i++) r.push_back(bool_op(kind, a[i], b[i]));
            return r;
        }
        Instr in = mk(OP_BOOLV);
        in.a[0] = (uint32_t)kind;
        in.a[1] = (uint32_t)a.size();
        in.a[2] = (uint32_t)shape.tables.size();
        for (size_t i = 0; i < a.size(); i++) {
            shape.tables.push_back(a[i].v.slot);
            shape.tables.push_back(b[i].v.slot);
        }
        Macro m(*this, in);
        std::vector<AssignedCondition> r;
        for (size_t i = 0; i < a.size(); i++) r.push_back(bool_op(kind, a[i], b[i]));
        m.expect_cells((uint32_t)(3 * a.size()));
        return r;
    }
    // keccak's chi on n bits: per element t = not_and(u, v), then xor(s, t), as ONE macro-op
    std::vector<AssignedCondition> chi_vec(const std::vector<AssignedCondition>& u, const std::vector<AssignedCondition>& v,
                                           const std::vector<AssignedCondition>& s) {
        if (u.size() != v.size() || u.size() != s.size() || u.empty()) throw std::logic_error("chi_vec: operand lists differ in length");
        Instr in = mk(OP_CHIV);
        in.a[1] = (uint32_t)u.size();
        in.a[2] = (uint32_t)shape.tables.size();
        const bool top = depth == 0;
        if (top)
            for (size_t i = 0; i < u.size(); i++) {
                shape.tables.push_back(u[i].v.slot);
                shape.tables.push_back(v[i].v.slot);
                shape.tables.push_back(s[i].v.slot);
            }
        Macro m(*this, in);
        std::vector<AssignedCondition> r;
        for (size_t i = 0; i < u.size(); i++) {
            AssignedCondition t = bool_op(4, u[i], v[i]);
            r.push_back(bool_op(2, s[i], t));
        }
        m.expect_cells((uint32_t)(6 * u.size()));
        return r;
    }
    AssignedCondition or_(const AssignedCondition& a, const AssignedCondition& b) { return bool_op(1, a, b); }
    AssignedCondition xor_(const AssignedCondition& a, const AssignedCondition& b) { return bool_op(2, a, b); }
    AssignedCondition xnor(const AssignedCondition& a, const AssignedCondition& b) { return bool_op(3, a, b); }
    AssignedCondition not_and(const AssignedCondition& a, const AssignedCondition& b) { return bool_op(4, a, b); }
    // base_chip.rs:574-598
    AssignedValue bisec(const AssignedCondition& cond, const AssignedValue& a, const AssignedValue& b) {
        Instr in = mk(OP_BISEC);
        in.a[0] = cond.v.slot;
        in.a[1] = a.slot;
        in.a[2] = b.slot;
        Macro m(*this, in);
        return bisec_row(cond, a, b);
    }
    AssignedValue bisec_row(const AssignedCondition& cond, const AssignedValue& a, const AssignedValue& b) {
        return one_line_with_last({Pair(&cond.v, ZERO), Pair(&a, ZERO), Pair(&cond.v, ZERO), Pair(&b, ONE)}, Pair(ValueSchema(), NEG_ONE),
                                  nullptr, {ONE, NEG_ONE}, nullptr);
    }
    AssignedCondition bisec_cond(const AssignedCondition& cond, const AssignedCondition& a, const AssignedCondition& b) {
        return AssignedCondition{bisec(cond, a.v, b.v)};
    }
    // base_chip.rs:487-500. Value asserts become per-instance status bits.
    void assert_true(const AssignedCondition& a) { assert_constant(a.v, ONE); }
    void assert_false(const AssignedCondition& a) { assert_constant(a.v, ZERO); }
    void try_assert_false(const AssignedCondition& a, uint32_t status_bit) { assert_constant(a.v, ZERO, status_bit); }

    void note_input(uint32_t cell, uint32_t n = 1) {
        if (cell + n > shape.n_inputs) shape.n_inputs = cell + n;
    }

    // ---- Records range writers (context.rs:835-997) ----
    AssignedValue assign_one_line_range_value(uint64_t bits) {
        size_t offset = range_offset;
        ensure_range(offset + 1);
        fix(RangeChip, AccLinesCol, offset, Big(1));
        fix(RangeChip, TagCol, offset, Big(bits));
        new_slot(RangeChip, TaggedRangeCol, offset);
        uint32_t s = new_slot(RangeChip, ValueAccCol, offset);
        range_offset += 1;
        return AssignedValue{{RangeChip, ValueAccCol, (uint32_t)offset}, s};
    }
    AssignedValue assign_two_line_range_value(uint64_t bits) {
        const uint64_t C = COMMON_RANGE_BITS;
        if (bits < 2 * C || bits > 4 * C) throw std::logic_error("2-line range bits");
        size_t offset = range_offset;
        ensure_range(offset + 2);
        fix(RangeChip, AccLinesCol, offset, Big(2));
        new_slot(RangeChip, CommonRangeCol, offset);
        new_slot(RangeChip, CommonRangeCol, offset + 1);
        fix(RangeChip, TagCol, offset, Big(bits >= 3 * C ? C : bits % C));
        new_slot(RangeChip, TaggedRangeCol, offset);
        fix(RangeChip, TagCol, offset + 1, Big(bits > 3 * C ? bits - 3 * C : 0));
        new_slot(RangeChip, TaggedRangeCol, offset + 1);
        uint32_t s = new_slot(RangeChip, ValueAccCol, offset);
        range_offset += 2;
        return AssignedValue{{RangeChip, ValueAccCol, (uint32_t)offset}, s};
    }
    AssignedValue assign_three_line_range_value(uint64_t bits) {
        const uint64_t C = COMMON_RANGE_BITS;
        if (bits < 3 * C || bits > 6 * C) throw std::logic_error("3-line range bits");
        size_t offset = range_offset;
        ensure_range(offset + 3);
        fix(RangeChip, AccLinesCol, offset, Big(3));
        new_slot(RangeChip, CommonRangeCol, offset);
        new_slot(RangeChip, CommonRangeCol, offset + 1);
        new_slot(RangeChip, CommonRangeCol, offset + 2);
        fix(RangeChip, TagCol, offset, Big(bits >= 4 * C ? C : bits % C));
        new_slot(RangeChip, TaggedRangeCol, offset);
        fix(RangeChip, TagCol, offset + 1, Big(bits >= 5 * C ? C : (bits > 4 * C ? bits % C : 0)));
        new_slot(RangeChip, TaggedRangeCol, offset + 1);
        fix(RangeChip, TagCol, offset + 2, Big(bits > 5 * C ? bits - 5 * C : 0));
        new_slot(RangeChip, TaggedRangeCol, offset + 2);
        uint32_t s = new_slot(RangeChip, ValueAccCol, offset);
        range_offset += 3;
        return AssignedValue{{RangeChip, ValueAccCol, (uint32_t)offset}, s};
    }
    void ensure_range(size_t offset) {  // context.rs:716-720
        if (offset >= shape.height[1]) shape.height[1] = offset + 1;
    }

    // ---- Records select writers (context.rs:749-801) ----
    void assign_cache_value(const AssignedValue& v, const Big& encode) {
        size_t offset = select_offset;
        if (offset >= shape.height[2]) shape.height[2] = offset + 1;
        new_slot(SelectChip, SelValueCol, offset);
        permute(Cell{SelectChip, SelValueCol, (uint32_t)offset}, v.cell);
        fix(SelectChip, EncodeCol, offset, encode);
        fix(SelectChip, IsLookupCol, offset, ZERO);
        select_offset += 1;
    }
    AssignedValue assign_select_value(const Big& encode, const AssignedValue& selector) {
        size_t offset = select_offset;
        if (offset >= shape.height[2]) shape.height[2] = offset + 1;
        uint32_t s = new_slot(SelectChip, SelValueCol, offset);
        new_slot(SelectChip, SelSelectCol, offset);
        permute(Cell{SelectChip, SelSelectCol, (uint32_t)offset}, selector.cell);
        fix(SelectChip, EncodeCol, offset, encode);
        fix(SelectChip, IsLookupCol, offset, ONE);
        select_offset += 1;
        return AssignedValue{{SelectChip, SelValueCol, (uint32_t)offset}, s};
    }

    void finish() {
        shape.offset[0] = base_offset;
        shape.offset[1] = range_offset;
        shape.offset[2] = select_offset;
    }
};

// IntegerContext<W,N> (context.rs:161-188) with RangeChipOps / IntegerChipOps.
class IntegerContext {
   public:
    Context* ctx;
    Field field;
    const FieldInfo* info;

    IntegerContext(Context* c, Field f) : ctx(c), field(f), info(&field_info(f)) {}

    unsigned L() const { return info->limbs; }

    // ---- RangeChipOps (range_chip.rs:287-347) ----
    AssignedValue assign_common() { return ctx->assign_one_line_range_value(COMMON_RANGE_BITS); }
    AssignedValue assign_range(uint64_t bits) {  // Records::assign_range_value (context.rs:974-997)
        if (bits <= COMMON_RANGE_BITS) return ctx->assign_one_line_range_value(bits);
        if (bits < 2 * COMMON_RANGE_BITS) throw std::logic_error("unreachable range bits");
        if (bits <= 4 * COMMON_RANGE_BITS) return ctx->assign_two_line_range_value(bits);
        if (bits <= 6 * COMMON_RANGE_BITS) return ctx->assign_three_line_range_value(bits);
        throw std::logic_error("unreachable range bits");
    }
    AssignedValue assign_nonleading_limb() { return assign_range(info->limb_bits); }
    AssignedValue assign_w_ceil_leading_limb() { return assign_range(info->w_ceil_bits % info->limb_bits); }
    AssignedValue assign_d_leading_limb() { return assign_range(info->d_bits % info->limb_bits); }

    // ---- helpers ----
    void put_int(Instr& in, int at, const AssignedInteger& a, bool with_native) const {
        for (unsigned i = 0; i < L(); i++) in.a[at + i] = a.limbs_le[i].slot;
        if (with_native) in.a[at + L()] = a.native.slot;
    }
    AssignedValue native_sum(const std::vector<AssignedValue>& limbs) {
        std::vector<Context::Elem> e;
        for (unsigned i = 0; i < limbs.size(); i++) e.push_back(Context::Elem(&limbs[i], info->limb_coeffs[i]));
        return ctx->sum_with_constant(e, nullptr);
    }
    // rows of assign_w / assign_d (integer_chip.rs:236-281)
    AssignedInteger assign_w_rows() {
        AssignedInteger r;
        for (unsigned i = 0; i + 1 < L(); i++) r.limbs_le.push_back(assign_nonleading_limb());
        r.limbs_le.push_back(assign_w_ceil_leading_limb());
        r.native = native_sum(r.limbs_le);
        r.times = 1;
        return r;
    }
    AssignedInteger assign_d_rows() {
        AssignedInteger r;
        for (unsigned i = 0; i + 1 < L(); i++) r.limbs_le.push_back(assign_nonleading_limb());
        r.limbs_le.push_back(assign_d_leading_limb());
        r.native = native_sum(r.limbs_le);
        r.times = 1;
        return r;
    }
    // integer_chip.rs:73-215
    void mul_equation_rows(const AssignedInteger& a, const AssignedInteger& b, const AssignedInteger& d, const AssignedInteger& rem) {
        if (a.times >= info->overflow_limit || b.times >= info->overflow_limit || rem.times != 1)
            throw std::logic_error("times overflow in mul equation (integer_chip.rs:80-82)");
        const Big &ONE = ctx->ONE, &NEG_ONE = ctx->NEG_ONE;
        unsigned l = L();
        std::vector<AssignedValue> limbs;
        for (unsigned pos = 0; pos < info->mul_check_limbs; pos++) {
            unsigned r_bound = std::min(pos + 1, l);
            unsigned l_bound = pos >= l - 1 ? pos - (l - 1) : 0;
            std::vector<Context::MulAddTerm> terms;
            for (unsigned i = l_bound; i < r_bound; i++)
                terms.push_back({&a.limbs_le[i], &b.limbs_le[pos - i], &d.limbs_le[i], n_neg(info->w_modulus_limbs_le[pos - i])});
            limbs.push_back(ctx->mul_add_with_next_line(terms));
        }
        Big B = info->limb_modulus;
        Big borrow = Big(l) * B + Big(2);
        Big c0 = B * borrow, c1 = B * borrow - borrow;
        AssignedValue u = ctx->sum_with_constant({Context::Elem(&limbs[0], ONE), Context::Elem(&rem.limbs_le[0], NEG_ONE)}, &c0);
        AssignedValue v_h = assign_common();
        AssignedValue v_l = assign_nonleading_limb();
        ctx->one_line_with_last({Pair(&v_h, info->limb_coeffs[2]), Pair(&v_l, info->limb_coeffs[1])}, Pair(&u, NEG_ONE), nullptr, {}, nullptr);
        for (unsigned i = 1; i < info->mul_check_limbs; i++) {
            std::vector<Context::Elem> e;
            e.push_back(Context::Elem(&limbs[i], ONE));
            if (i < l) e.push_back(Context::Elem(&rem.limbs_le[i], NEG_ONE));
            e.push_back(Context::Elem(&v_h, info->limb_coeffs[1]));
            e.push_back(Context::Elem(&v_l, info->limb_coeffs[0]));
            AssignedValue ui = ctx->sum_with_constant(e, &c1);
            v_h = assign_common();
            v_l = assign_nonleading_limb();
            ctx->one_line_with_last({Pair(&v_h, info->limb_coeffs[2]), Pair(&v_l, info->limb_coeffs[1])}, Pair(&ui, NEG_ONE), nullptr, {},
                                    nullptr);
        }
        // native (integer_chip.rs:195-215)
        ctx->one_line({Pair(&a.native, ctx->ZERO), Pair(&b.native, ctx->ZERO), Pair(&d.native, info->w_native), Pair(&rem.native, ONE)},
                      nullptr, {NEG_ONE}, nullptr);
    }

    // ---- IntegerChipOps ----
    // harness prelude for tests/benches: an integer whose (possibly overflowed) limbs come from
    // per-instance input cells in_cell, in_cell+2, ... (each logical input is 64 bytes = 2 cells)
    // packed: the L limbs sit back to back, 16 bytes each, from input cell in_cell on (one 64-byte logical input)
    AssignedInteger load_int(uint64_t times, uint32_t in_cell, bool packed = false) {
        Instr in = Context::mk(OP_LOAD_INT, field);
        in.a[0] = in_cell;
        in.a[1] = packed ? 1 : 0;
        ctx->note_input(in_cell, packed ? 2 : 2 * L());
        Context::Macro m(*ctx, in);
        AssignedInteger r;
        for (unsigned i = 0; i <= L(); i++) {
            AssignedValue v = ctx->one_line({Pair(ValueSchema(), ctx->ZERO)}, nullptr, {}, nullptr)[0];
            if (i < L())
                r.limbs_le.push_back(v);
            else
                r.native = v;
        }
        r.times = times;
        return r;
    }
    // integer_chip.rs:236-258
    AssignedInteger assign_w(uint32_t in_cell) {
        Instr in = Context::mk(OP_ASSIGN_W, field);
        in.a[0] = in_cell;
        ctx->note_input(in_cell, 2);
        Context::Macro m(*ctx, in);
        return assign_w_rows();
    }
    // assign_w of a shape-level constant value (e.g. the curve generator in msm_unsafe)
    AssignedInteger assign_w_static(const Big& w) {
        Instr in = Context::mk(OP_ASSIGN_W, field);
        in.a[0] = ctx->shape.add_const_wide(w);
        in.a[1] = 1;
        Context::Macro m(*ctx, in);
        return assign_w_rows();
    }
    // integer_chip.rs:580-598 (shape-level constant)
    AssignedInteger assign_int_constant(const Big& w) {
        Instr in = Context::mk(OP_ASSIGN_INT_CONST, field);
        in.a[0] = 1;
        in.a[1] = ctx->shape.add_const_wide(w);
        Context::Macro m(*ctx, in);
        AssignedInteger r;
        for (unsigned i = 0; i < L(); i++) r.limbs_le.push_back(ctx->assign_constant((w >> (i * info->limb_bits)).low_bits(info->limb_bits)));
        r.native = ctx->assign_constant(w % info->n_modulus);
        r.times = 1;
        return r;
    }
    // per-instance constant (e.g. the G2 points the pairing tests pass as constants)
    AssignedInteger assign_int_constant_input(uint32_t in_cell) {
        Instr in = Context::mk(OP_ASSIGN_INT_CONST, field);
        in.a[0] = 0;
        in.a[1] = in_cell;
        ctx->note_input(in_cell, 2);
        Context::Macro m(*ctx, in);
        AssignedInteger r;
        for (unsigned i = 0; i < L(); i++) r.limbs_le.push_back(ctx->assign_constant_input(in_cell));
        r.native = ctx->assign_constant_input(in_cell);
        r.times = 1;
        return r;
    }
    // integer_chip.rs:283-373
    AssignedInteger reduce(const AssignedInteger& a) {
        if (a.times == 1) return a;
        if (a.times >= info->overflow_limit) throw std::logic_error("times overflow in reduce (integer_chip.rs:293)");
        Instr in = Context::mk(OP_REDUCE, field);
        put_int(in, 0, a, true);
        Context::Macro m(*ctx, in);
        const Big &ONE = ctx->ONE, &NEG_ONE = ctx->NEG_ONE;
        AssignedInteger rem = assign_w_rows();
        AssignedValue d = assign_common();
        ctx->one_line_with_last({Pair(&d, info->w_native), Pair(&rem.native, ONE)}, Pair(&a.native, NEG_ONE), nullptr, {}, nullptr);
        bool have_last = false;
        AssignedValue last_v;
        Big B = info->limb_modulus;
        for (unsigned i = 0; i < info->reduce_check_limbs; i++) {
            AssignedValue v = assign_nonleading_limb();
            Coeff k(B * Big(info->overflow_limit) - Big(i == 0 ? 0 : info->overflow_limit));
            ctx->one_line_with_last({Pair(&d, info->w_modulus_limbs_le[i]), Pair(&rem.limbs_le[i], ONE), Pair(&a.limbs_le[i], NEG_ONE),
                                     have_last ? Pair(&last_v, ONE) : Pair(ValueSchema(), ctx->ZERO)},
                                    Pair(&v, n_neg(B)), &k, {}, nullptr);
            last_v = v;
            have_last = true;
        }
        return rem;
    }
    // integer_chip.rs:375-382
    AssignedInteger conditionally_reduce(const AssignedInteger& a) {
        uint64_t threshold = 1ull << (OVERFLOW_BITS - 2);
        return a.times > threshold ? reduce(a) : a;
    }
    // integer_chip.rs:384-406
    AssignedInteger int_add(const AssignedInteger& a, const AssignedInteger& b) {
        AssignedInteger r;
        {
            Instr in = Context::mk(OP_INT_ADD, field);
            put_int(in, 0, a, false);
            put_int(in, L(), b, false);
            in.a[2 * L()] = a.native.slot;
            in.a[2 * L() + 1] = b.native.slot;
            Context::Macro m(*ctx, in);
            for (unsigned i = 0; i < L(); i++) r.limbs_le.push_back(ctx->add(a.limbs_le[i], b.limbs_le[i]));
            r.native = native_sum(r.limbs_le);
            r.times = a.times + b.times;
        }
        return conditionally_reduce(r);
    }
    // integer_chip.rs:408-437
    AssignedInteger int_sub(const AssignedInteger& a, const AssignedInteger& b) {
        if (b.times < 1 || b.times >= info->overflow_limit) throw std::logic_error("int_sub: b.times out of range");
        AssignedInteger r;
        {
            Instr in = Context::mk(OP_INT_SUB, field);
            put_int(in, 0, a, false);
            put_int(in, L(), b, false);
            in.a[2 * L()] = (uint32_t)b.times;
            in.a[2 * L() + 1] = a.native.slot;
            in.a[2 * L() + 2] = b.native.slot;
            Context::Macro m(*ctx, in);
            const auto& upper = info->w_modulus_of_ceil_times[b.times];
            for (unsigned i = 0; i < L(); i++)
                r.limbs_le.push_back(
                    ctx->sum_with_constant({Context::Elem(&a.limbs_le[i], ctx->ONE), Context::Elem(&b.limbs_le[i], ctx->NEG_ONE)}, &upper[i]));
            r.native = native_sum(r.limbs_le);
            r.times = a.times + b.times + 1;
        }
        return conditionally_reduce(r);
    }
    // integer_chip.rs:439-464
    AssignedInteger int_neg(const AssignedInteger& a) {
        if (a.times < 1 || a.times >= info->overflow_limit) throw std::logic_error("int_neg: a.times out of range");
        AssignedInteger r;
        {
            Instr in = Context::mk(OP_INT_NEG, field);
            put_int(in, 0, a, false);
            in.a[L()] = (uint32_t)a.times;
            in.a[L() + 1] = a.native.slot;
            Context::Macro m(*ctx, in);
            const auto& upper = info->w_modulus_of_ceil_times[a.times];
            for (unsigned i = 0; i < L(); i++)
                r.limbs_le.push_back(ctx->sum_with_constant({Context::Elem(&a.limbs_le[i], ctx->NEG_ONE)}, &upper[i]));
            r.native = native_sum(r.limbs_le);
            r.times = a.times + 1;
        }
        return conditionally_reduce(r);
    }
    // integer_chip.rs:466-483
    AssignedInteger int_mul(const AssignedInteger& a, const AssignedInteger& b) {
        Instr in = Context::mk(OP_INT_MUL, field);
        put_int(in, 0, a, true);
        put_int(in, L() + 1, b, true);
        Context::Macro m(*ctx, in);
        AssignedInteger rem = assign_w_rows();
        AssignedInteger d = assign_d_rows();
        mul_equation_rows(a, b, d, rem);
        return rem;
    }
    AssignedInteger int_square(const AssignedInteger& a) { return int_mul(a, a); }
    // integer_chip.rs:485-491
    AssignedInteger int_unsafe_invert(const AssignedInteger& x) {
        AssignedInteger one = assign_int_constant(Big(1));
        auto r = int_div(one, x);
        ctx->assert_false(r.first);
        return r.second;
    }
    // integer_chip.rs:493-538
    std::pair<AssignedCondition, AssignedInteger> int_div(const AssignedInteger& a_in, const AssignedInteger& b_in) {
        AssignedInteger b = reduce(b_in);
        AssignedCondition is_b_zero = is_int_zero(b);
        AssignedCondition a_coeff = ctx->not_(is_b_zero);
        AssignedInteger a;
        {
            AssignedInteger ar = reduce(a_in);
            Instr in = Context::mk(OP_MASK_INT, field);
            put_int(in, 0, ar, true);
            in.a[L() + 1] = a_coeff.v.slot;
            Context::Macro m(*ctx, in);
            for (unsigned i = 0; i < L(); i++) a.limbs_le.push_back(ctx->mul(ar.limbs_le[i], a_coeff.v));
            a.native = ctx->mul(ar.native, a_coeff.v);
            a.times = ar.times;
        }
        Instr in = Context::mk(OP_DIV_CORE, field);
        put_int(in, 0, a, true);
        put_int(in, L() + 1, b, true);
        Context::Macro m(*ctx, in);
        AssignedInteger c = assign_w_rows();
        AssignedInteger d = assign_d_rows();
        mul_equation_rows(b, c, d, a);
        return {is_b_zero, c};
    }
    // integer_chip.rs:540-578 (a.times == 1 after reduce)
    AssignedCondition is_int_zero(const AssignedInteger& a_in) {
        AssignedInteger a = reduce(a_in);
        Instr in = Context::mk(OP_IS_INT_ZERO, field);
        put_int(in, 0, a, true);
        Context::Macro m(*ctx, in);
        // is_pure_zero
        std::vector<Context::Elem> e;
        for (auto& v : a.limbs_le) e.push_back(Context::Elem(&v, ctx->ONE));
        AssignedValue sum = ctx->sum_with_constant(e, nullptr);
        AssignedCondition is_zero = ctx->is_zero(sum);
        // is_pure_w_modulus
        if (a.times != 1) throw std::logic_error("is_pure_w_modulus: times != 1");
        AssignedValue native_diff = ctx->add_constant(a.native, n_neg(info->w_native));
        AssignedCondition is_eq = ctx->is_zero(native_diff);
        for (unsigned i = 0; i < info->pure_w_check_limbs; i++) {
            AssignedValue limb_diff = ctx->add_constant(a.limbs_le[i], n_neg(info->w_modulus_limbs_le[i]));
            AssignedCondition is_limb_eq = ctx->is_zero(limb_diff);
            is_eq = ctx->and_(is_eq, is_limb_eq);
        }
        return ctx->or_(is_zero, is_eq);
    }
    AssignedCondition is_int_equal(const AssignedInteger& a, const AssignedInteger& b) {  // integer_chip.rs:47-54
        AssignedInteger diff = int_sub(a, b);
        return is_int_zero(diff);
    }
    // integer_chip.rs:600-612
    void assert_int_equal(const AssignedInteger& a, const AssignedInteger& b) {
        AssignedInteger diff = int_sub(a, b);
        diff = reduce(diff);
        Instr in = Context::mk(OP_SUM_ASSERT_ZERO, field);
        put_int(in, 0, diff, false);
        Context::Macro m(*ctx, in);
        std::vector<Context::Elem> e;
        for (auto& v : diff.limbs_le) e.push_back(Context::Elem(&v, ctx->ONE));
        AssignedValue sum = ctx->sum_with_constant(e, nullptr);
        ctx->assert_constant(sum, ctx->ZERO);
    }
    // integer_chip.rs:618-658
    AssignedInteger int_mul_small_constant(const AssignedInteger& a_in, uint64_t b) {
        uint64_t threshold = 1ull << (OVERFLOW_BITS - 2);
        if (b >= threshold) throw std::logic_error("int_mul_small_constant: b too large (integer_chip.rs:624)");
        AssignedInteger a = (a_in.times * b >= info->overflow_limit) ? reduce(a_in) : a_in;
        AssignedInteger r;
        {
            Instr in = Context::mk(OP_MUL_SMALL, field);
            put_int(in, 0, a, false);
            in.a[L()] = (uint32_t)b;
            in.a[L() + 1] = a.native.slot;
            Context::Macro m(*ctx, in);
            for (unsigned i = 0; i < L(); i++) r.limbs_le.push_back(ctx->sum_with_constant({Context::Elem(&a.limbs_le[i], Big(b))}, nullptr));
            r.native = native_sum(r.limbs_le);
            r.times = a.times * b;
        }
        return conditionally_reduce(r);
    }
    // integer_chip.rs:660-681
    AssignedInteger bisec_int(const AssignedCondition& cond, const AssignedInteger& a, const AssignedInteger& b) {
        Instr in = Context::mk(OP_BISEC_INT, field);
        in.a[0] = cond.v.slot;
        put_int(in, 1, a, true);
        put_int(in, 2 + L(), b, true);
        Context::Macro m(*ctx, in);
        AssignedInteger r;
        for (unsigned i = 0; i < L(); i++) r.limbs_le.push_back(ctx->bisec_row(cond, a.limbs_le[i], b.limbs_le[i]));
        r.native = ctx->bisec_row(cond, a.native, b.native);
        r.times = std::max(a.times, b.times);
        return r;
    }
};

}  // namespace h2e
