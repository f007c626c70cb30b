// Builds a Shape from an op-script: a flat list of (opcode, nargs, args...) chip calls. This is the
// C-ABI's generic circuit-construction entry (h2e_shape_from_script): a binding in any language
// can replay the exact call sequence it would make against the reference's traits.
// Logical input i (64 bytes) occupies per-instance input cells 2i and 2i+1.
#pragma once
#include "circuits.h"
#include "keccak_tracer.h"

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
    S_LOAD_INT_PACKED = 16,  // times, in_idx: like S_LOAD_INT with the L limbs packed into ONE logical input, limb i = bits [128 i, 128 i + 128)
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
    // ---- EccChipBaseOps / EccChipScalarOps on the curve of the script's field (bn256 G1 for
    //      H2E_FIELD_BN256_FQ, bls12_381 G1 for H2E_FIELD_BLS12_381_FQ); results go to the point / curvature lists ----
    S_ASSIGN_POINT = 40,             // in_idx (x, y, z = 3 logical inputs)        -> point   ecc_chip.rs:458-487
    S_TO_POINT_WITH_CURVATURE = 41,  // point                                      -> pwc     ecc_chip.rs:779-794
    S_ECC_ADD = 42,                  // pwc, point                                 -> point   ecc_chip.rs:606-669
    S_ECC_DOUBLE = 43,               // pwc                                        -> point   ecc_chip.rs:671-690
    S_ECC_NEG = 44,                  // point                                      -> point
    S_ECC_REDUCE = 45,               // point                                      -> point
    S_ECC_ASSERT_EQUAL = 46,         // point, point
    S_ECC_ENCODE = 47,               // point                                      -> 3 vals  ecc_chip.rs:710-732
    S_MSM = 48,                      // n, n points, n scalar vals, r1 in_idx, r2 in_idx -> point   (native scalars; ecc_chip.rs:373-408
                                     //   with the blinding points r1, r2 as inputs: 2 logical inputs each)
    // ---- PairingChipOps (src/circuit/pairing_chip.rs:13-176) ----
    S_ASSIGN_G2_CONSTANT = 50,       // in_idx (x.c0, x.c1, y.c0, y.c1 = 4 logical inputs) -> g2   (G2 as per-instance constants, as the
                                     //   reference's pairing tests assign it: native_scalar_pairing_chip.rs:74-93)
    S_CHECK_PAIRING = 51,            // n, then n x (point, g2)                              pairing_chip.rs:170-176
    S_PAIRING = 52,                  // n, then n x (point, g2)                    -> fq12   pairing_chip.rs:157-168 (no final assert)
    S_MULTI_MILLER_LOOP = 53,        // n, then n x (point, g2)                    -> fq12   prepare_g2 + multi_miller_loop
    S_FINAL_EXPONENTIATION = 54,     // fq12                                       -> fq12
    // ---- Fq2 / Fq6 / Fq12ChipOps (src/circuit/fq12.rs:10-459); elements live in their own result lists ----
    S_FQ2_FROM_INTS = 60,            // int c0, int c1                             -> fq2    (AssignedFq2 = (AssignedInteger, AssignedInteger))
    S_FQ2_ADD = 61, S_FQ2_SUB = 62, S_FQ2_MUL = 63,     // fq2, fq2                -> fq2
    S_FQ2_NEG = 64, S_FQ2_DOUBLE = 65, S_FQ2_MUL_BY_NONRESIDUE = 66, S_FQ2_UNSAFE_INVERT = 67, S_FQ2_REDUCE = 68,   // fq2 -> fq2
    S_FQ2_FROBENIUS_MAP = 69,        // fq2, power                                 -> fq2
    S_FQ2_ASSERT_EQUAL = 70,         // fq2, fq2
    S_FQ2_PARTS = 71,                // fq2                                        -> 2 ints (c0, c1)
    S_FQ6_FROM_FQ2S = 75,            // fq2 c0, c1, c2                             -> fq6
    S_FQ6_ADD = 76, S_FQ6_SUB = 77, S_FQ6_MUL = 78,     // fq6, fq6                -> fq6
    S_FQ6_NEG = 79, S_FQ6_UNSAFE_INVERT = 80,           // fq6                     -> fq6
    S_FQ6_MUL_BY_1 = 81,             // fq6, fq2 b1                                -> fq6
    S_FQ6_MUL_BY_01 = 82,            // fq6, fq2 b0, fq2 b1                        -> fq6
    S_FQ6_FROBENIUS_MAP = 83,        // fq6, power                                 -> fq6
    S_FQ6_ASSERT_EQUAL = 84,         // fq6, fq6
    S_FQ12_FROM_FQ6S = 90,           // fq6 c0, c1                                 -> fq12
    S_FQ12_MUL = 91,                 // fq12, fq12                                 -> fq12
    S_FQ12_MUL_BY_014 = 92, S_FQ12_MUL_BY_034 = 93,     // fq12, fq2, fq2, fq2     -> fq12
    S_FQ12_CYCLOTOMIC_SQUARE = 94, S_FQ12_UNSAFE_INVERT = 95,   // fq12            -> fq12
    S_FQ12_FROBENIUS_MAP = 96,       // fq12, power                                -> fq12
    S_FQ12_ASSERT_EQ = 97,           // fq12, fq12
    S_FQ12_ASSERT_ONE = 98,          // fq12
    S_FQ12_PARTS = 99,               // fq12                                       -> 2 fq6 -> (pushes c0, c1 to the fq6 list)
    // ---- more of EccChipBaseOps / EccChipScalarOps ----
    S_ECC_REDUCE_WITH_CURVATURE = 100,  // point                                   -> pwc    ecc_chip.rs:692-708
    S_ECC_MUL = 101,                 // point, scalar val, r1 in_idx, r2 in_idx    -> point  ecc_chip.rs:416-420 (one-term msm; native scalar)
    S_ASSIGN_SCALAR_W = 102,         // in_idx                                     -> scalar int (general-scalar context: assign_w in the scalar field)
    S_MSM_GENERAL = 103,             // n, n points, n scalar ints, r1 in_idx, r2 in_idx -> point  general_scalar_ecc_chip.rs:96-147 (bls12_381)
    // ---- KeccakChipOps (src/circuit/keccak_chip.rs:53-307); states live in their own list and are updated in place ----
    S_KECCAK_HASH = 110,             // n, n vals                                  -> val    keccak_chip.rs:231-300
    S_KECCAK_INIT = 111,             //                                            -> state
    S_KECCAK_ABSORB = 112,           // state, 1088 vals (bits)                              keccak_chip.rs:142-166
    S_KECCAK_PERMUTE = 113,          // state
    S_KECCAK_STEP = 114,             // state, which (0 theta, 1 rho_and_pi, 2 xi, 3 iota), round
    S_KECCAK_DECOMPOSE_U256 = 115,   // val                                        -> 256 vals (bits, most significant first)
    S_KECCAK_COMPOSE = 116,          // n, n vals (bits, most significant first)   -> val    compose_to_scalar_be
    S_KECCAK_LANE = 117,             // state, x, y                                -> 64 vals (the lane's bits)
};

// Argument count of every fixed-arity script op (-1: variadic, checked where it is decoded; -2: unknown opcode).
inline int script_arity(uint32_t op) {
    switch (op) {
        case S_LOAD_INT: case S_LOAD_INT_PACKED: case S_ASSIGN_INT_CONSTANT: case S_INT_ADD: case S_INT_SUB: case S_INT_MUL: case S_INT_DIV:
        case S_MUL_SMALL_CONST: case S_IS_INT_EQUAL: case S_ASSERT_INT_EQUAL: case S_ASSIGN_CONSTANT: case S_AND: case S_OR:
        case S_XOR: case S_XNOR: case S_NOT_AND: case S_ADD: case S_SUB: case S_MUL: case S_ASSERT_EQUAL: case S_ECC_ADD:
        case S_ECC_ASSERT_EQUAL:
            return 2;
        case S_BISEC_INT: case S_BISEC:
            return 3;
        case S_ASSIGN_W: case S_INT_NEG: case S_INT_SQUARE: case S_REDUCE: case S_IS_INT_ZERO: case S_INT_UNSAFE_INVERT:
        case S_ASSIGN: case S_ASSIGN_BIT: case S_NOT: case S_ASSERT_TRUE: case S_ASSERT_FALSE: case S_IS_ZERO:
        case S_ASSIGN_POINT: case S_TO_POINT_WITH_CURVATURE: case S_ECC_DOUBLE: case S_ECC_NEG: case S_ECC_REDUCE:
        case S_ECC_ENCODE: case S_ASSIGN_G2_CONSTANT:
            return 1;
        case S_FQ2_FROM_INTS: case S_FQ2_ADD: case S_FQ2_SUB: case S_FQ2_MUL: case S_FQ2_FROBENIUS_MAP: case S_FQ2_ASSERT_EQUAL:
        case S_FQ6_ADD: case S_FQ6_SUB: case S_FQ6_MUL: case S_FQ6_MUL_BY_1: case S_FQ6_FROBENIUS_MAP: case S_FQ6_ASSERT_EQUAL:
        case S_FQ12_FROM_FQ6S: case S_FQ12_MUL: case S_FQ12_FROBENIUS_MAP: case S_FQ12_ASSERT_EQ:
            return 2;
        case S_FQ2_NEG: case S_FQ2_DOUBLE: case S_FQ2_MUL_BY_NONRESIDUE: case S_FQ2_UNSAFE_INVERT: case S_FQ2_REDUCE: case S_FQ2_PARTS:
        case S_FQ6_NEG: case S_FQ6_UNSAFE_INVERT: case S_FQ12_CYCLOTOMIC_SQUARE: case S_FQ12_UNSAFE_INVERT: case S_FQ12_ASSERT_ONE:
        case S_FQ12_PARTS: case S_FINAL_EXPONENTIATION: case S_ECC_REDUCE_WITH_CURVATURE: case S_ASSIGN_SCALAR_W:
            return 1;
        case S_FQ6_FROM_FQ2S: case S_FQ6_MUL_BY_01: case S_KECCAK_STEP: case S_KECCAK_LANE:
            return 3;
        case S_KECCAK_INIT:
            return 0;
        case S_KECCAK_PERMUTE: case S_KECCAK_DECOMPOSE_U256:
            return 1;
        case S_KECCAK_ABSORB:
            return 1 + (int)KeccakOps::RATE_BITS;
        case S_FQ12_MUL_BY_014: case S_FQ12_MUL_BY_034: case S_ECC_MUL:
            return 4;
        case S_MSM: case S_CHECK_PAIRING: case S_PAIRING: case S_MULTI_MILLER_LOOP: case S_MSM_GENERAL: case S_KECCAK_HASH: case S_KECCAK_COMPOSE:
            return -1;
        default: return -2;
    }
}

inline void run_script(Context& ctx, Field field, const uint32_t* s, size_t n, const std::vector<Big>& statics) {
    IntegerContext ic(&ctx, field);
    std::vector<AssignedInteger> ints;
    std::vector<AssignedValue> vals;
    std::vector<AssignedPoint> points;
    std::vector<AssignedPointWithCurvature> pwcs;
    std::vector<AssignedG2Affine> g2s;
    std::vector<AssignedFq2> fq2s;
    std::vector<AssignedFq6> fq6s;
    std::vector<AssignedFq12> fq12s;
    std::vector<AssignedInteger> sints;  // integers of the scalar field (general-scalar context)
    std::vector<std::unique_ptr<KeccakOps::State>> kstates;
    KeccakOps keccak(&ctx);
    std::unique_ptr<EccContext> ecc;
    std::unique_ptr<PairingOps> pairing;
    auto E = [&]() -> EccContext& {
        if (!ecc) {
            if (field == F_BN256_FQ) ecc.reset(new EccContext(&ctx, curve_bn256_g1(), true, true));
            else if (field == F_BLS12_381_FQ) ecc.reset(new EccContext(&ctx, curve_bls12_381_g1(), false, true));
            else throw std::runtime_error("this field is not the base field of a supported curve");
        }
        return *ecc;
    };
    auto PC = [&]() -> PairingOps& {
        if (!pairing) pairing.reset(new PairingOps(E(), field == F_BN256_FQ));
        return *pairing;
    };
    auto C = [&](uint32_t i) { return AssignedCondition{vals.at(i)}; };
    size_t p = 0;
    while (p < n) {
        if (p + 2 > n) throw std::runtime_error("truncated script");
        uint32_t op = s[p], na = s[p + 1];
        const uint32_t* a = s + p + 2;
        p += 2 + na;
        if (p > n || p < 2 + (size_t)na) throw std::runtime_error("truncated script");
        {
            const int ar = script_arity(op);
            if (ar == -2) throw std::runtime_error("unknown script op " + std::to_string(op));
            if (ar >= 0 && na != (uint32_t)ar)
                throw std::runtime_error("script op " + std::to_string(op) + " takes " + std::to_string(ar) + " arguments, record has " + std::to_string(na));
            if (ar == -1 && na < 1) throw std::runtime_error("variadic script op " + std::to_string(op) + " needs its count argument");
        }
        switch (op) {
            case S_LOAD_INT: ints.push_back(ic.load_int(a[0], 2 * a[1])); break;
            case S_LOAD_INT_PACKED: ints.push_back(ic.load_int(a[0], 2 * a[1], true)); break;
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
            case S_ASSIGN_POINT: points.push_back(E().assign_point(PointInput{2 * a[0], 2 * (a[0] + 1)}, 2 * (a[0] + 2))); break;
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
                if (field != F_BN256_FQ) throw std::runtime_error("script MSM takes native scalars: bn256 only");
                uint32_t m = a[0];
                if (m == 0 || m > (1u << 24) || na != 2 * m + 3) throw std::runtime_error("bad MSM record");
                std::vector<AssignedPoint> ps;
                std::vector<AssignedScalar> ss;
                for (uint32_t i = 0; i < m; i++) ps.push_back(points.at(a[1 + i]));
                for (uint32_t i = 0; i < m; i++) {
                    AssignedScalar sc;
                    sc.v = vals.at(a[1 + m + i]);
                    ss.push_back(sc);
                }
                uint32_t r1 = a[1 + 2 * m], r2 = a[2 + 2 * m];
                points.push_back(E().msm_unsafe(ps, ss, PointInput{2 * r1, 2 * (r1 + 1)}, PointInput{2 * r2, 2 * (r2 + 1)}));
                break;
            }
            case S_ASSIGN_G2_CONSTANT: g2s.push_back(g2_constant_input(ctx, PC(), a[0])); break;
            case S_CHECK_PAIRING:
            case S_PAIRING:
            case S_MULTI_MILLER_LOOP: {
                uint32_t m = a[0];
                if (m == 0 || m > (1u << 24) || na != 2 * m + 1) throw std::runtime_error("bad pairing record");
                std::vector<std::pair<const AssignedPoint*, const AssignedG2Affine*>> terms;
                for (uint32_t i = 0; i < m; i++) terms.push_back({&points.at(a[1 + 2 * i]), &g2s.at(a[2 + 2 * i])});
                if (op == S_CHECK_PAIRING) {
                    PC().check_pairing(terms);
                } else if (op == S_PAIRING) {
                    fq12s.push_back(PC().pairing(terms));
                } else {
                    std::vector<AssignedG2Prepared> prepared;
                    for (auto& t : terms) prepared.push_back(PC().prepare_g2(*t.second));
                    PairingOps::Terms pt;
                    for (size_t i = 0; i < terms.size(); i++) pt.push_back({terms[i].first, &prepared[i]});
                    fq12s.push_back(PC().multi_miller_loop(pt));
                }
                break;
            }
            case S_FINAL_EXPONENTIATION: fq12s.push_back(PC().final_exponentiation(fq12s.at(a[0]))); break;
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
                ints.push_back(fq2s.at(a[0]).c0);
                ints.push_back(fq2s.at(a[0]).c1);
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
                if (field != F_BN256_FQ) throw std::runtime_error("script ecc_mul takes a native scalar: bn256 only");
                AssignedScalar sc;
                sc.v = vals.at(a[1]);
                points.push_back(E().msm_unsafe({points.at(a[0])}, {sc}, PointInput{2 * a[2], 2 * (a[2] + 1)}, PointInput{2 * a[3], 2 * (a[3] + 1)}));
                break;
            }
            case S_ASSIGN_SCALAR_W:
                if (field != F_BLS12_381_FQ) throw std::runtime_error("scalar-field integers exist in the general-scalar context only (bls12_381)");
                sints.push_back(E().scalar.assign_w(2 * a[0]));
                break;
            case S_MSM_GENERAL: {
                if (field != F_BLS12_381_FQ) throw std::runtime_error("general-scalar MSM: bls12_381 only");
                uint32_t m = a[0];
                if (m == 0 || m > (1u << 24) || na != 2 * m + 3) throw std::runtime_error("bad MSM record");
                std::vector<AssignedPoint> ps;
                std::vector<AssignedScalar> ss;
                for (uint32_t i = 0; i < m; i++) ps.push_back(points.at(a[1 + i]));
                for (uint32_t i = 0; i < m; i++) {
                    AssignedScalar sc;
                    sc.i = sints.at(a[1 + m + i]);
                    ss.push_back(sc);
                }
                uint32_t r1 = a[1 + 2 * m], r2 = a[2 + 2 * m];
                points.push_back(E().msm_unsafe(ps, ss, PointInput{2 * r1, 2 * (r1 + 1)}, PointInput{2 * r2, 2 * (r2 + 1)}));
                break;
            }
            case S_KECCAK_HASH: {
                const uint32_t m = a[0];
                if (m == 0 || m > (1u << 16) || na != m + 1) throw std::runtime_error("bad keccak hash record");
                std::vector<AssignedValue> in;
                for (uint32_t i = 0; i < m; i++) in.push_back(vals.at(a[1 + i]));
                vals.push_back(keccak.hash(in));
                break;
            }
            case S_KECCAK_INIT: kstates.emplace_back(new KeccakOps::State(keccak.init())); break;
            case S_KECCAK_ABSORB: {
                std::vector<AssignedCondition> bits;
                for (size_t i = 0; i < KeccakOps::RATE_BITS; i++) bits.push_back(C(a[1 + i]));
                keccak.absorb(*kstates.at(a[0]), bits.data(), bits.size());
                break;
            }
            case S_KECCAK_PERMUTE: keccak.permute(*kstates.at(a[0])); break;
            case S_KECCAK_STEP: {
                KeccakOps::State& st = *kstates.at(a[0]);
                if (a[2] >= (uint32_t)KeccakOps::N_R) throw std::runtime_error("keccak round out of range");
                switch (a[1]) {
                    case 0: keccak.theta(st); break;
                    case 1: keccak.rho_and_pi(st); break;
                    case 2: keccak.xi(st); break;
                    case 3: keccak.iota(st, (int)a[2]); break;
                    default: throw std::runtime_error("unknown keccak step");
                }
                break;
            }
            case S_KECCAK_DECOMPOSE_U256:
                for (const AssignedCondition& b : keccak.decompose_scalar_as_u256_be(vals.at(a[0]))) vals.push_back(b.v);
                break;
            case S_KECCAK_COMPOSE: {
                const uint32_t m = a[0];
                if (m > (1u << 20) || na != m + 1) throw std::runtime_error("bad keccak compose record");
                std::vector<AssignedCondition> bits;
                for (uint32_t i = 0; i < m; i++) bits.push_back(C(a[1 + i]));
                vals.push_back(keccak.compose_to_scalar_be(bits));
                break;
            }
            case S_KECCAK_LANE: {
                if (a[1] >= 5 || a[2] >= 5) throw std::runtime_error("keccak lane out of range");
                for (const AssignedCondition& b : (*kstates.at(a[0]))[a[1]][a[2]]) vals.push_back(b.v);
                break;
            }
            default: throw std::runtime_error("unknown script op");
        }
    }
    ctx.finish();
}

}  // namespace h2e
