// ORACLE (test infrastructure, NOT product code).
// A tiny op-script runner so tests can drive arbitrary sequences of the reference's chip ops
// (the same call sequences the reference tests make, e.g. src/tests/integer_chip.rs:11-99) through
// the oracle, and the same script through the product's builder API, and compare records.
#pragma once
#include "pairing.h"

namespace orc {

enum ScriptOp : uint32_t {
    S_LOAD_INT = 0,             // times, in_idx(L limb values)          -> int   (harness: cells via `assign` rows)
    S_ASSIGN_W = 1,             // in_idx                                -> int   integer_chip.rs:236
    S_ASSIGN_INT_CONSTANT = 2,  // src(0 input,1 static), idx            -> int   integer_chip.rs:580
    S_INT_ADD = 3,              // a b -> int
    S_INT_SUB = 4,              // a b -> int
    S_INT_NEG = 5,              // a -> int
    S_INT_MUL = 6,              // a b -> int
    S_INT_SQUARE = 7,           // a -> int
    S_INT_DIV = 8,              // a b -> val(cond), int
    S_REDUCE = 9,               // a -> int
    S_MUL_SMALL_CONST = 10,     // a k -> int
    S_BISEC_INT = 11,           // cond a b -> int
    S_IS_INT_ZERO = 12,         // a -> val
    S_IS_INT_EQUAL = 13,        // a b -> val
    S_ASSERT_INT_EQUAL = 14,    // a b
    S_INT_UNSAFE_INVERT = 15,   // a -> int
    S_ASSIGN = 20,              // in_idx -> val
    S_ASSIGN_CONSTANT = 21,     // src, idx -> val
    S_ASSIGN_BIT = 22,          // in_idx -> val
    S_AND = 23,
    S_OR = 24,
    S_NOT = 25,
    S_XOR = 26,
    S_XNOR = 27,
    S_NOT_AND = 28,
    S_BISEC = 29,  // cond a b -> val
    S_ADD = 30,
    S_SUB = 31,
    S_MUL = 32,
    S_ASSERT_TRUE = 34,
    S_ASSERT_FALSE = 35,
    S_IS_ZERO = 36,
    S_ASSERT_EQUAL = 37,
    // EccChipBaseOps / EccChipScalarOps on the curve of the script's field (see the product's script_builder.h)
    S_ASSIGN_POINT = 40,
    S_TO_POINT_WITH_CURVATURE = 41,
    S_ECC_ADD = 42,
    S_ECC_DOUBLE = 43,
    S_ECC_NEG = 44,
    S_ECC_REDUCE = 45,
    S_ECC_ASSERT_EQUAL = 46,
    S_ECC_ENCODE = 47,
    S_MSM = 48,
    S_ASSIGN_G2_CONSTANT = 50,
    S_CHECK_PAIRING = 51,
};

struct ScriptRunner {
    IntegerContext& ic;
    const std::vector<BN>& inputs;
    const std::vector<BN>& statics;
    std::vector<AssignedInteger> ints;
    std::vector<AssignedValue> vals;
    std::vector<AssignedPoint> points;
    std::vector<AssignedPointWithCurvature> pwcs;
    std::shared_ptr<Context> ctx;  // needed by the ECC ops
    int field = -1;
    std::vector<AssignedG2Affine> g2s;
    std::unique_ptr<EccContext> ecc;
    std::unique_ptr<PairingContext> pairing;
    PairingContext& PC() {
        if (!pairing) pairing.reset(new PairingContext(E(), field == 0));
        return *pairing;
    }

    ScriptRunner(IntegerContext& i, const std::vector<BN>& in, const std::vector<BN>& st) : ic(i), inputs(in), statics(st) {}

    EccContext& E() {
        if (!ecc) {
            ORC_ASSERT(ctx && (field == 0 || field == 1));
            if (field == 0) ecc.reset(new EccContext(EccContext::native(ctx, bn256_g1(), BN256_FQ(), true)));
            else ecc.reset(new EccContext(EccContext::general(ctx, bls12_381_g1(), BLS12_381_FQ())));
        }
        return *ecc;
    }
    HostPoint host_point(uint32_t in_idx, bool with_z) const {
        HostPoint p;
        p.x = inputs.at(in_idx);
        p.y = inputs.at(in_idx + 1);
        p.identity = with_z && !inputs.at(in_idx + 2).is_zero();
        return p;
    }

    const BN& src(uint32_t kind, uint32_t idx) const { return kind == 0 ? inputs.at(idx) : statics.at(idx); }

    // Harness prelude (not a reference op): materialise an integer whose limbs are arbitrary
    // (possibly overflowed, `times` > 1) values as if earlier ops had produced it. Cells are
    // created with BaseChipOps::assign (base_chip.rs:351-355) so that permutations are well-formed.
    AssignedInteger load_int(uint64_t times, uint32_t in_idx) {
        std::vector<AssignedValue> limbs;
        N native = n_from(0);
        for (uint64_t i = 0; i < ic.info->limbs; i++) {
            N lv = bn_to_n(inputs.at(in_idx + i));
            limbs.push_back(ic.base().assign(lv));
            native = n_add(native, n_mul(lv, ic.info->limb_coeffs[i]));
        }
        AssignedValue nat = ic.base().assign(native);
        return AssignedInteger(limbs, nat, times);
    }

    void run(const uint32_t* s, size_t n) {
        size_t p = 0;
        while (p < n) {
            uint32_t op = s[p], na = s[p + 1];
            const uint32_t* a = s + p + 2;
            p += 2 + na;
            BaseOps b = ic.base();
            switch (op) {
                case S_LOAD_INT: ints.push_back(load_int(a[0], a[1])); break;
                case S_ASSIGN_W: ints.push_back(ic.assign_w(inputs.at(a[0]))); break;
                case S_ASSIGN_INT_CONSTANT: ints.push_back(ic.assign_int_constant(src(a[0], a[1]))); break;
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
                case S_BISEC_INT: ints.push_back(ic.bisec_int(AssignedCondition(vals.at(a[0])), ints.at(a[1]), ints.at(a[2]))); break;
                case S_IS_INT_ZERO: vals.push_back(ic.is_int_zero(ints.at(a[0])).v); break;
                case S_IS_INT_EQUAL: vals.push_back(ic.is_int_equal(ints.at(a[0]), ints.at(a[1])).v); break;
                case S_ASSERT_INT_EQUAL: ic.assert_int_equal(ints.at(a[0]), ints.at(a[1])); break;
                case S_INT_UNSAFE_INVERT: ints.push_back(ic.int_unsafe_invert(ints.at(a[0]))); break;
                case S_ASSIGN: vals.push_back(b.assign(bn_to_n(inputs.at(a[0])))); break;
                case S_ASSIGN_CONSTANT: vals.push_back(b.assign_constant(bn_to_n(src(a[0], a[1])))); break;
                case S_ASSIGN_BIT: vals.push_back(b.assign_bit(bn_to_n(inputs.at(a[0]))).v); break;
                case S_AND: vals.push_back(b.and_(AssignedCondition(vals.at(a[0])), AssignedCondition(vals.at(a[1]))).v); break;
                case S_OR: vals.push_back(b.or_(AssignedCondition(vals.at(a[0])), AssignedCondition(vals.at(a[1]))).v); break;
                case S_NOT: vals.push_back(b.not_(AssignedCondition(vals.at(a[0]))).v); break;
                case S_XOR: vals.push_back(b.xor_(AssignedCondition(vals.at(a[0])), AssignedCondition(vals.at(a[1]))).v); break;
                case S_XNOR: vals.push_back(b.xnor(AssignedCondition(vals.at(a[0])), AssignedCondition(vals.at(a[1]))).v); break;
                case S_NOT_AND: vals.push_back(b.not_and(AssignedCondition(vals.at(a[0])), AssignedCondition(vals.at(a[1]))).v); break;
                case S_BISEC: vals.push_back(b.bisec(AssignedCondition(vals.at(a[0])), vals.at(a[1]), vals.at(a[2]))); break;
                case S_ADD: vals.push_back(b.add(vals.at(a[0]), vals.at(a[1]))); break;
                case S_SUB: vals.push_back(b.sub(vals.at(a[0]), vals.at(a[1]))); break;
                case S_MUL: vals.push_back(b.mul(vals.at(a[0]), vals.at(a[1]))); break;
                case S_ASSERT_TRUE: b.assert_true(AssignedCondition(vals.at(a[0]))); break;
                case S_ASSERT_FALSE: b.assert_false(AssignedCondition(vals.at(a[0]))); break;
                case S_IS_ZERO: vals.push_back(b.is_zero(vals.at(a[0])).v); break;
                case S_ASSERT_EQUAL: b.assert_equal(vals.at(a[0]), vals.at(a[1])); break;
                case S_ASSIGN_POINT: points.push_back(E().assign_point(host_point(a[0], true))); break;
                case S_TO_POINT_WITH_CURVATURE: pwcs.push_back(E().to_point_with_curvature(points.at(a[0]))); break;
                case S_ECC_ADD: points.push_back(E().ecc_add(pwcs.at(a[0]), points.at(a[1]))); break;
                case S_ECC_DOUBLE: points.push_back(E().ecc_double(pwcs.at(a[0]))); break;
                case S_ECC_NEG: points.push_back(E().ecc_neg(points.at(a[0]))); break;
                case S_ECC_REDUCE: points.push_back(E().ecc_reduce(points.at(a[0]))); break;
                case S_ECC_ASSERT_EQUAL: E().ecc_assert_equal(points.at(a[0]), points.at(a[1])); break;
                case S_ECC_ENCODE:
                    for (const AssignedValue& v : E().ecc_encode(points.at(a[0]))) vals.push_back(v);
                    break;
                case S_MSM: {
                    ORC_ASSERT(field == 0);
                    uint32_t m = a[0];
                    std::vector<AssignedPoint> ps;
                    std::vector<AssignedScalar> ss;
                    for (uint32_t i = 0; i < m; i++) ps.push_back(points.at(a[1 + i]));
                    for (uint32_t i = 0; i < m; i++) {
                        AssignedScalar sc;
                        sc.v = vals.at(a[1 + m + i]);
                        ss.push_back(sc);
                    }
                    points.push_back(E().msm_unsafe(ps, ss, host_point(a[1 + 2 * m], false), host_point(a[2 + 2 * m], false)));
                    break;
                }
                case S_ASSIGN_G2_CONSTANT: {
                    AssignedFq2 x = PC().fq2_assign_constant(HFq2{inputs.at(a[0]), inputs.at(a[0] + 1)});
                    AssignedFq2 y = PC().fq2_assign_constant(HFq2{inputs.at(a[0] + 2), inputs.at(a[0] + 3)});
                    AssignedValue z = E().bc().assign_constant(n_from(0));
                    g2s.push_back(AssignedG2Affine{x, y, AssignedCondition(z)});
                    break;
                }
                case S_CHECK_PAIRING: {
                    uint32_t m = a[0];
                    std::vector<std::pair<const AssignedPoint*, const AssignedG2Affine*>> terms;
                    for (uint32_t i = 0; i < m; i++) terms.push_back({&points.at(a[1 + 2 * i]), &g2s.at(a[2 + 2 * i])});
                    PC().check_pairing(terms);
                    break;
                }
                default: ORC_ASSERT(!"unknown script op");
            }
        }
    }
};

}  // namespace orc
