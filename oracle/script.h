// ORACLE (test infrastructure, NOT product code).
// A tiny op-script runner so tests can drive arbitrary sequences of the reference's chip ops
// (the same call sequences the reference tests make, e.g. src/tests/integer_chip.rs:11-99) through
// the oracle, and the same script through the product's builder API, and compare records.
#pragma once
#include "keccak.h"
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
    S_LOAD_INT_PACKED = 16,     // times, in_idx: limb i = bits [128 i, 128 i + 128) of ONE input value
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
    S_PAIRING = 52,
    S_MULTI_MILLER_LOOP = 53,
    S_FINAL_EXPONENTIATION = 54,
    // Fq2 / Fq6 / Fq12ChipOps (fq12.rs:10-459), see the product's script_builder.h for the argument lists
    S_FQ2_FROM_INTS = 60, S_FQ2_ADD = 61, S_FQ2_SUB = 62, S_FQ2_MUL = 63, S_FQ2_NEG = 64, S_FQ2_DOUBLE = 65, S_FQ2_MUL_BY_NONRESIDUE = 66,
    S_FQ2_UNSAFE_INVERT = 67, S_FQ2_REDUCE = 68, S_FQ2_FROBENIUS_MAP = 69, S_FQ2_ASSERT_EQUAL = 70, S_FQ2_PARTS = 71,
    S_FQ6_FROM_FQ2S = 75, S_FQ6_ADD = 76, S_FQ6_SUB = 77, S_FQ6_MUL = 78, S_FQ6_NEG = 79, S_FQ6_UNSAFE_INVERT = 80, S_FQ6_MUL_BY_1 = 81,
    S_FQ6_MUL_BY_01 = 82, S_FQ6_FROBENIUS_MAP = 83, S_FQ6_ASSERT_EQUAL = 84,
    S_FQ12_FROM_FQ6S = 90, S_FQ12_MUL = 91, S_FQ12_MUL_BY_014 = 92, S_FQ12_MUL_BY_034 = 93, S_FQ12_CYCLOTOMIC_SQUARE = 94,
    S_FQ12_UNSAFE_INVERT = 95, S_FQ12_FROBENIUS_MAP = 96, S_FQ12_ASSERT_EQ = 97, S_FQ12_ASSERT_ONE = 98, S_FQ12_PARTS = 99,
    S_ECC_REDUCE_WITH_CURVATURE = 100, S_ECC_MUL = 101, S_ASSIGN_SCALAR_W = 102, S_MSM_GENERAL = 103,
    // KeccakChipOps (keccak_chip.rs:53-307), see the product's script_builder.h for the argument lists
    S_KECCAK_HASH = 110, S_KECCAK_INIT = 111, S_KECCAK_ABSORB = 112, S_KECCAK_PERMUTE = 113, S_KECCAK_STEP = 114,
    S_KECCAK_DECOMPOSE_U256 = 115, S_KECCAK_COMPOSE = 116, S_KECCAK_LANE = 117,
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
    std::vector<AssignedFq2> fq2s;
    std::vector<AssignedFq6> fq6s;
    std::vector<AssignedFq12> fq12s;
    std::vector<AssignedInteger> sints;
    std::vector<std::unique_ptr<KeccakOps::AssignedState>> kstates;
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
    AssignedInteger load_int(uint64_t times, uint32_t in_idx, bool packed = false) {
        std::vector<AssignedValue> limbs;
        N native = n_from(0);
        for (uint64_t i = 0; i < ic.info->limbs; i++) {
            N lv = bn_to_n(packed ? (inputs.at(in_idx) >> (128 * i)) & ((BN(1) << 128) - BN(1)) : inputs.at(in_idx + i));
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
                case S_LOAD_INT_PACKED: ints.push_back(load_int(a[0], a[1], true)); break;
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
                case S_CHECK_PAIRING:
                case S_PAIRING:
                case S_MULTI_MILLER_LOOP: {
                    uint32_t m = a[0];
                    std::vector<std::pair<const AssignedPoint*, const AssignedG2Affine*>> terms;
                    for (uint32_t i = 0; i < m; i++) terms.push_back({&points.at(a[1 + 2 * i]), &g2s.at(a[2 + 2 * i])});
                    if (op == S_CHECK_PAIRING) {
                        PC().check_pairing(terms);
                    } else if (op == S_PAIRING) {
                        fq12s.push_back(PC().pairing(terms));
                    } else {
                        const bool bn = field == 0;
                        std::vector<AssignedG2Prepared> prepared;
                        for (auto& t : terms) prepared.push_back(bn ? PC().bn_prepare_g2(*t.second) : PC().bls_prepare_g2(*t.second));
                        std::vector<std::pair<const AssignedPoint*, const AssignedG2Prepared*>> pt;
                        for (size_t i = 0; i < terms.size(); i++) pt.push_back({terms[i].first, &prepared[i]});
                        fq12s.push_back(bn ? PC().bn_multi_miller_loop(pt) : PC().bls_multi_miller_loop(pt));
                    }
                    break;
                }
                case S_FINAL_EXPONENTIATION:
                    fq12s.push_back(field == 0 ? PC().bn_final_exponentiation(fq12s.at(a[0])) : PC().bls_final_exponentiation(fq12s.at(a[0])));
                    break;
                case S_FQ2_FROM_INTS: fq2s.push_back(AssignedFq2{ints.at(a[0]), ints.at(a[1])}); break;
                case S_FQ2_ADD: fq2s.push_back(PC().fq2_add(fq2s.at(a[0]), fq2s.at(a[1]))); break;
                case S_FQ2_SUB: fq2s.push_back(PC().fq2_sub(fq2s.at(a[0]), fq2s.at(a[1]))); break;
                case S_FQ2_MUL: fq2s.push_back(PC().fq2_mul(fq2s.at(a[0]), fq2s.at(a[1]))); break;
                case S_FQ2_NEG: fq2s.push_back(PC().fq2_neg(fq2s.at(a[0]))); break;
                case S_FQ2_DOUBLE: fq2s.push_back(PC().fq2_double(fq2s.at(a[0]))); break;
                case S_FQ2_MUL_BY_NONRESIDUE: fq2s.push_back(PC().fq2_mul_by_nonresidue(fq2s.at(a[0]))); break;
                case S_FQ2_UNSAFE_INVERT: fq2s.push_back(PC().fq2_unsafe_invert(fq2s.at(a[0]))); break;
                case S_FQ2_REDUCE: fq2s.push_back(PC().fq2_reduce(fq2s.at(a[0]))); break;
                case S_FQ2_FROBENIUS_MAP: fq2s.push_back(PC().fq2_frobenius_map(fq2s.at(a[0]), a[1])); break;
                case S_FQ2_ASSERT_EQUAL: PC().fq2_assert_equal(fq2s.at(a[0]), fq2s.at(a[1])); break;
                case S_FQ2_PARTS:
                    ints.push_back(fq2s.at(a[0]).first);
                    ints.push_back(fq2s.at(a[0]).second);
                    break;
                case S_FQ6_FROM_FQ2S: fq6s.push_back(AssignedFq6{fq2s.at(a[0]), fq2s.at(a[1]), fq2s.at(a[2])}); break;
                case S_FQ6_ADD: fq6s.push_back(PC().fq6_add(fq6s.at(a[0]), fq6s.at(a[1]))); break;
                case S_FQ6_SUB: fq6s.push_back(PC().fq6_sub(fq6s.at(a[0]), fq6s.at(a[1]))); break;
                case S_FQ6_MUL: fq6s.push_back(PC().fq6_mul(fq6s.at(a[0]), fq6s.at(a[1]))); break;
                case S_FQ6_NEG: fq6s.push_back(PC().fq6_neg(fq6s.at(a[0]))); break;
                case S_FQ6_UNSAFE_INVERT: fq6s.push_back(PC().fq6_unsafe_invert(fq6s.at(a[0]))); break;
                case S_FQ6_MUL_BY_1: fq6s.push_back(PC().fq6_mul_by_1(fq6s.at(a[0]), fq2s.at(a[1]))); break;
                case S_FQ6_MUL_BY_01: fq6s.push_back(PC().fq6_mul_by_01(fq6s.at(a[0]), fq2s.at(a[1]), fq2s.at(a[2]))); break;
                case S_FQ6_FROBENIUS_MAP: fq6s.push_back(PC().fq6_frobenius_map(fq6s.at(a[0]), a[1])); break;
                case S_FQ6_ASSERT_EQUAL: PC().fq6_assert_equal(fq6s.at(a[0]), fq6s.at(a[1])); break;
                case S_FQ12_FROM_FQ6S: fq12s.push_back(AssignedFq12{fq6s.at(a[0]), fq6s.at(a[1])}); break;
                case S_FQ12_MUL: fq12s.push_back(PC().fq12_mul(fq12s.at(a[0]), fq12s.at(a[1]))); break;
                case S_FQ12_MUL_BY_014: fq12s.push_back(PC().fq12_mul_by_014(fq12s.at(a[0]), fq2s.at(a[1]), fq2s.at(a[2]), fq2s.at(a[3]))); break;
                case S_FQ12_MUL_BY_034: fq12s.push_back(PC().fq12_mul_by_034(fq12s.at(a[0]), fq2s.at(a[1]), fq2s.at(a[2]), fq2s.at(a[3]))); break;
                case S_FQ12_CYCLOTOMIC_SQUARE: fq12s.push_back(PC().fq12_cyclotomic_square(fq12s.at(a[0]))); break;
                case S_FQ12_UNSAFE_INVERT: fq12s.push_back(PC().fq12_unsafe_invert(fq12s.at(a[0]))); break;
                case S_FQ12_FROBENIUS_MAP: fq12s.push_back(PC().fq12_frobenius_map(fq12s.at(a[0]), a[1])); break;
                case S_FQ12_ASSERT_EQ: PC().fq12_assert_eq(fq12s.at(a[0]), fq12s.at(a[1])); break;
                case S_FQ12_ASSERT_ONE: PC().fq12_assert_one(fq12s.at(a[0])); break;
                case S_FQ12_PARTS:
                    fq6s.push_back(fq12s.at(a[0]).c0);
                    fq6s.push_back(fq12s.at(a[0]).c1);
                    break;
                case S_ECC_REDUCE_WITH_CURVATURE: pwcs.push_back(E().ecc_reduce_with_curvature(points.at(a[0]))); break;
                case S_ECC_MUL: {
                    ORC_ASSERT(field == 0);
                    AssignedScalar sc;
                    sc.v = vals.at(a[1]);
                    points.push_back(E().msm_unsafe({points.at(a[0])}, {sc}, host_point(a[2], false), host_point(a[3], false)));
                    break;
                }
                case S_ASSIGN_SCALAR_W:
                    ORC_ASSERT(field == 1);
                    sints.push_back(E().scalar->assign_w(inputs.at(a[0])));
                    break;
                case S_MSM_GENERAL: {
                    ORC_ASSERT(field == 1);
                    uint32_t m = a[0];
                    std::vector<AssignedPoint> ps;
                    std::vector<AssignedScalar> ss;
                    for (uint32_t i = 0; i < m; i++) ps.push_back(points.at(a[1 + i]));
                    for (uint32_t i = 0; i < m; i++) {
                        AssignedScalar sc;
                        sc.i = sints.at(a[1 + m + i]);
                        ss.push_back(sc);
                    }
                    points.push_back(E().msm_unsafe(ps, ss, host_point(a[1 + 2 * m], false), host_point(a[2 + 2 * m], false)));
                    break;
                }
                case S_KECCAK_HASH: {
                    std::vector<AssignedValue> in;
                    for (uint32_t i = 0; i < a[0]; i++) in.push_back(vals.at(a[1 + i]));
                    vals.push_back(KeccakOps(b.c).hash(in));
                    break;
                }
                case S_KECCAK_INIT: kstates.emplace_back(new KeccakOps::AssignedState(KeccakOps(b.c).init())); break;
                case S_KECCAK_ABSORB: {
                    std::vector<AssignedCondition> bits;
                    for (size_t i = 0; i < KeccakOps::ABSORB_BITS_RATE; i++) bits.push_back(AssignedCondition(vals.at(a[1 + i])));
                    KeccakOps(b.c).absorb(*kstates.at(a[0]), bits);
                    break;
                }
                case S_KECCAK_PERMUTE: KeccakOps(b.c).permute(*kstates.at(a[0])); break;
                case S_KECCAK_STEP: {
                    KeccakOps k(b.c);
                    KeccakOps::AssignedState& st = *kstates.at(a[0]);
                    if (a[1] == 0) k.theta(st);
                    else if (a[1] == 1) k.rho_and_pi(st);
                    else if (a[1] == 2) k.xi(st);
                    else k.iota(st, a[2]);
                    break;
                }
                case S_KECCAK_DECOMPOSE_U256:
                    for (const AssignedCondition& bit : KeccakOps(b.c).decompose_scalar_as_u256_be(vals.at(a[0]))) vals.push_back(bit.v);
                    break;
                case S_KECCAK_COMPOSE: {
                    std::vector<AssignedCondition> bits;
                    for (uint32_t i = 0; i < a[0]; i++) bits.push_back(AssignedCondition(vals.at(a[1 + i])));
                    vals.push_back(KeccakOps(b.c).compose_to_scalar_be(bits));
                    break;
                }
                case S_KECCAK_LANE:
                    for (const AssignedCondition& bit : (*kstates.at(a[0]))[a[1]][a[2]]) vals.push_back(bit.v);
                    break;
                default: ORC_ASSERT(!"unknown script op");
            }
        }
    }
};

}  // namespace orc
