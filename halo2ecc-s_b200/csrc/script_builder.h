// Builds a Shape from an op-script: a flat list of (opcode, nargs, args...) chip calls. This is the
// C-ABI's generic circuit-construction entry (h2e_shape_from_script): a binding in any language
// can replay the exact call sequence it would make against the reference's traits.
// Logical input i (64 bytes) occupies per-instance input cells 2i and 2i+1.
#pragma once
#include "tracer.h"

namespace h2e {

enum ScriptOp : uint32_t {
    S_LOAD_INT = 0,
    S_ASSIGN_W = 1,
    S_ASSIGN_INT_CONSTANT = 2,
    S_INT_ADD = 3,
    S_INT_SUB = 4,
    S_INT_NEG = 5,
    S_INT_MUL = 6,
    S_INT_SQUARE = 7,
    S_INT_DIV = 8,
    S_REDUCE = 9,
    S_MUL_SMALL_CONST = 10,
    S_BISEC_INT = 11,
    S_IS_INT_ZERO = 12,
    S_IS_INT_EQUAL = 13,
    S_ASSERT_INT_EQUAL = 14,
    S_INT_UNSAFE_INVERT = 15,
    S_ASSIGN = 20,
    S_ASSIGN_CONSTANT = 21,
    S_ASSIGN_BIT = 22,
    S_AND = 23,
    S_OR = 24,
    S_NOT = 25,
    S_XOR = 26,
    S_XNOR = 27,
    S_NOT_AND = 28,
    S_BISEC = 29,
    S_ADD = 30,
    S_SUB = 31,
    S_MUL = 32,
    S_ASSERT_TRUE = 34,
    S_ASSERT_FALSE = 35,
    S_IS_ZERO = 36,
    S_ASSERT_EQUAL = 37,
};

inline void run_script(Context& ctx, Field field, const uint32_t* s, size_t n, const std::vector<Big>& statics) {
    IntegerContext ic(&ctx, field);
    std::vector<AssignedInteger> ints;
    std::vector<AssignedValue> vals;
    auto C = [&](uint32_t i) { return AssignedCondition{vals.at(i)}; };
    size_t p = 0;
    while (p < n) {
        if (p + 2 > n) throw std::runtime_error("truncated script");
        uint32_t op = s[p], na = s[p + 1];
        const uint32_t* a = s + p + 2;
        p += 2 + na;
        if (p > n) throw std::runtime_error("truncated script");
        switch (op) {
            case S_LOAD_INT: ints.push_back(ic.load_int(a[0], 2 * a[1])); break;
            case S_ASSIGN_W: ints.push_back(ic.assign_w(2 * a[0])); break;
            case S_ASSIGN_INT_CONSTANT:
                ints.push_back(a[0] == 0 ? ic.assign_int_constant_input(2 * a[1]) : ic.assign_int_constant(statics.at(a[1])));
                break;
            case S_INT_ADD: ints.push_back(ic.int_add(ints.at(a[0]), ints.at(a[1]))); break;
            case S_INT_SUB: ints.push_back(ic.int_sub(ints.at(a[0]), ints.at(a[1]))); break;
            case S_INT_NEG: ints.push_back(ic.int_neg(ints.at(a[0]))); break;
            case S_INT_MUL: ints.push_back(ic.int_mul(ints.at(a[0]), ints.at(a[1]))); break;
            case S_INT_SQUARE: ints.push_back(ic.int_square(ints.at(a[0]))); break;
            case S_INT_DIV: {
                auto r = ic.int_div(ints.at(a[0]), ints.at(a[1]));
                vals.push_back(r.first.v);
                ints.push_back(r.second);
                break;
            }
            case S_REDUCE: ints.push_back(ic.reduce(ints.at(a[0]))); break;
            case S_MUL_SMALL_CONST: ints.push_back(ic.int_mul_small_constant(ints.at(a[0]), a[1])); break;
            case S_BISEC_INT: ints.push_back(ic.bisec_int(C(a[0]), ints.at(a[1]), ints.at(a[2]))); break;
            case S_IS_INT_ZERO: vals.push_back(ic.is_int_zero(ints.at(a[0])).v); break;
            case S_IS_INT_EQUAL: vals.push_back(ic.is_int_equal(ints.at(a[0]), ints.at(a[1])).v); break;
            case S_ASSERT_INT_EQUAL: ic.assert_int_equal(ints.at(a[0]), ints.at(a[1])); break;
            case S_INT_UNSAFE_INVERT: ints.push_back(ic.int_unsafe_invert(ints.at(a[0]))); break;
            case S_ASSIGN: vals.push_back(ctx.assign(2 * a[0])); break;
            case S_ASSIGN_CONSTANT:
                vals.push_back(a[0] == 0 ? ctx.assign_constant_input(2 * a[1]) : ctx.assign_constant(statics.at(a[1]) % native_modulus()));
                break;
            case S_ASSIGN_BIT: vals.push_back(ctx.assign_bit(2 * a[0]).v); break;
            case S_AND: vals.push_back(ctx.and_(C(a[0]), C(a[1])).v); break;
            case S_OR: vals.push_back(ctx.or_(C(a[0]), C(a[1])).v); break;
            case S_NOT: vals.push_back(ctx.not_(C(a[0])).v); break;
            case S_XOR: vals.push_back(ctx.xor_(C(a[0]), C(a[1])).v); break;
            case S_XNOR: vals.push_back(ctx.xnor(C(a[0]), C(a[1])).v); break;
            case S_NOT_AND: vals.push_back(ctx.not_and(C(a[0]), C(a[1])).v); break;
            case S_BISEC: vals.push_back(ctx.bisec(C(a[0]), vals.at(a[1]), vals.at(a[2]))); break;
            case S_ADD: vals.push_back(ctx.add(vals.at(a[0]), vals.at(a[1]))); break;
            case S_SUB: vals.push_back(ctx.sub(vals.at(a[0]), vals.at(a[1]))); break;
            case S_MUL: vals.push_back(ctx.mul(vals.at(a[0]), vals.at(a[1]))); break;
            case S_ASSERT_TRUE: ctx.assert_true(C(a[0])); break;
            case S_ASSERT_FALSE: ctx.assert_false(C(a[0])); break;
            case S_IS_ZERO: vals.push_back(ctx.is_zero(vals.at(a[0])).v); break;
            case S_ASSERT_EQUAL: ctx.assert_equal(vals.at(a[0]), vals.at(a[1])); break;
            default: throw std::runtime_error("unknown script op");
        }
    }
    ctx.finish();
}

}  // namespace h2e
