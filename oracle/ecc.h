// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the ECC / MSM chips:
//   EccChipBaseOps / EccChipScalarOps / ParallelClone / Offset   src/circuit/ecc_chip.rs:23-1009
//   NativeScalarEccContext impl                                  src/circuit/native_scalar_ecc_chip.rs:27-194
//   GeneralScalarEccContext impl                                 src/circuit/general_scalar_ecc_chip.rs:25-169
// Curve points enter as affine coordinates (canonical integers) + identity flag; the random
// blinding points r1, r2 of msm_unsafe (ecc_chip.rs:378-379, `Scalar::rand()`) are explicit inputs.
#pragma once
#include "chips.h"

namespace orc {

struct UnsafeError {
    int code;  // 1 AddSameOrNegPoint, 2 AddIdentity, 3 AssignIdentity (ecc_chip.rs:23-28)
};

struct HostPoint {  // C::CurveExt in affine form
    BN x, y;
    bool identity = false;
};

struct AssignedPoint {
    AssignedInteger x, y;
    AssignedCondition z;
};
struct AssignedNonZeroPoint {
    AssignedInteger x, y;
};
struct AssignedCurvature {
    AssignedInteger v;
    AssignedCondition z;
};
struct AssignedPointWithCurvature {
    AssignedInteger x, y;
    AssignedCondition z;
    AssignedCurvature curvature;
    AssignedPoint to_point() const { return AssignedPoint{x, y, z}; }
};

struct Offset {  // ecc_chip.rs:36-62
    size_t range_offset_diff = 0, base_offset_diff = 0, select_offset_diff = 0;
    Offset operator-(const Offset& r) const {
        return Offset{range_offset_diff - r.range_offset_diff, base_offset_diff - r.base_offset_diff,
                      select_offset_diff - r.select_offset_diff};
    }
    Offset scale(size_t n) const { return Offset{range_offset_diff * n, base_offset_diff * n, select_offset_diff * n}; }
    bool operator==(const Offset& r) const {
        return range_offset_diff == r.range_offset_diff && base_offset_diff == r.base_offset_diff &&
               select_offset_diff == r.select_offset_diff;
    }
};

// AssignedScalar: AssignedValue for the native-scalar context, AssignedInteger for the general one.
struct AssignedScalar {
    AssignedValue v;
    AssignedInteger i;
};

struct CurveParams {
    BN b;            // curve constant (y^2 = x^3 + b)
    HostPoint gen;   // generator
    BN scalar_mod;   // scalar field modulus
    uint64_t scalar_num_bits;
};

inline CurveParams bn256_g1() {
    CurveParams c;
    c.b = BN(3);
    c.gen.x = BN(1);
    c.gen.y = BN(2);
    c.scalar_mod = BN256_FR();
    c.scalar_num_bits = 254;
    return c;
}
inline CurveParams bls12_381_g1() {
    CurveParams c;
    c.b = BN(4);
    c.gen.x = BN::from_hex("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb");
    c.gen.y = BN::from_hex("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1");
    c.scalar_mod = BLS12_381_FR();
    c.scalar_num_bits = 255;
    return c;
}

// One struct plays both NativeScalarEccContext<C> (context.rs:190-213) and
// GeneralScalarEccContext<C,N> (context.rs:215-239).
struct EccContext {
    IntegerContext base;                      // base_integer_ctx / .0
    std::shared_ptr<IntegerContext> scalar;   // scalar_integer_ctx (general only)
    std::shared_ptr<Context> native_ctx;
    size_t msm_prefix;                        // usize::MAX disables the select chip
    bool native_scalar;
    CurveParams curve;

    static EccContext native(std::shared_ptr<Context> c, const CurveParams& cp, const BN& base_mod, bool with_select) {
        return EccContext{IntegerContext(c, base_mod), nullptr, c, with_select ? 0 : (size_t)-1, true, cp};
    }
    static EccContext general(std::shared_ptr<Context> c, const CurveParams& cp, const BN& base_mod) {
        EccContext e{IntegerContext(c, base_mod), std::make_shared<IntegerContext>(c, cp.scalar_mod), c, 0, false, cp};
        return e;
    }

    BaseOps bc() { return BaseOps(native_ctx.get()); }
    bool has_select_chip() const { return native_scalar ? msm_prefix < (size_t)-1 : true; }

    // ---- ParallelClone (native_scalar_ecc_chip.rs:50-90, general_scalar_ecc_chip.rs:43-91) ----
    void apply_offset_diff(const Offset& d) {
        native_ctx->base_offset += d.base_offset_diff;
        native_ctx->range_offset += d.range_offset_diff;
        native_ctx->select_offset += d.select_offset_diff;
    }
    EccContext clone_with_offset(const Offset& d) const {
        auto ctx = std::make_shared<Context>(native_ctx->clone_without_permutation());
        ctx->base_offset += d.base_offset_diff;
        ctx->range_offset += d.range_offset_diff;
        ctx->select_offset += d.select_offset_diff;
        EccContext e{IntegerContext(ctx, base.info), scalar ? std::make_shared<IntegerContext>(ctx, scalar->info) : nullptr, ctx, msm_prefix,
                     native_scalar, curve};
        return e;
    }
    EccContext clone_without_offset() const { return clone_with_offset(Offset{}); }
    Offset offset() const { return Offset{native_ctx->range_offset, native_ctx->base_offset, native_ctx->select_offset}; }
    void merge(EccContext& other) {
        Records& record = native_ctx->records;
        Records& record_other = other.native_ctx->records;
        record.permutations.insert(record.permutations.end(), record_other.permutations.begin(), record_other.permutations.end());
        record_other.permutations.clear();
        record.base_height = std::max(record.base_height, record_other.base_height);
        record.range_height = std::max(record.select_height, record_other.range_height);  // sic (native_scalar_ecc_chip.rs:87)
        record.select_height = std::max(record.select_height, record_other.select_height);
    }

    // ---- EccChipBaseOps (ecc_chip.rs:438-1009) ----
    // ecc_chip.rs:441-456
    AssignedPoint assign_constant_point(const HostPoint& c) {
        BN x = c.identity ? BN(0) : c.x, y = c.identity ? BN(0) : c.y;
        AssignedInteger ax = base.assign_int_constant(x);
        AssignedInteger ay = base.assign_int_constant(y);
        AssignedValue z = bc().assign_constant(n_from(c.identity ? 1 : 0));
        return AssignedPoint{ax, ay, AssignedCondition(z)};
    }
    // ecc_chip.rs:458-487
    AssignedPoint assign_point(const HostPoint& c) {
        BN xv = c.identity ? BN(0) : c.x, yv = c.identity ? BN(0) : c.y;
        AssignedInteger x = base.assign_w(xv);
        AssignedInteger y = base.assign_w(yv);
        AssignedCondition z = bc().assign_bit(n_from(c.identity ? 1 : 0));
        AssignedInteger b = base.assign_int_constant(curve.b);
        AssignedInteger y2 = base.int_square(y);
        AssignedInteger x2 = base.int_square(x);
        AssignedInteger x3 = base.int_mul(x2, x);
        AssignedInteger right = base.int_add(x3, b);
        AssignedCondition eq = base.is_int_equal(y2, right);
        AssignedCondition eq_or_identity = bc().or_(eq, z);
        bc().assert_true(eq_or_identity);
        return AssignedPoint{x, y, z};
    }
    // ecc_chip.rs:489-512
    AssignedNonZeroPoint assign_non_zero_point(const HostPoint& c) {
        ORC_ASSERT(!c.identity);
        AssignedInteger x = base.assign_w(c.x);
        AssignedInteger y = base.assign_w(c.y);
        AssignedInteger b = base.assign_int_constant(curve.b);
        AssignedInteger y2 = base.int_square(y);
        AssignedInteger x2 = base.int_square(x);
        AssignedInteger x3 = base.int_mul(x2, x);
        AssignedInteger right = base.int_add(x3, b);
        base.assert_int_equal(y2, right);
        return AssignedNonZeroPoint{x, y};
    }
    // ecc_chip.rs:514-529
    AssignedPointWithCurvature assign_identity() {
        AssignedInteger zero = base.assign_int_constant(BN(0));
        AssignedValue one = bc().assign_constant(n_from(1));
        return AssignedPointWithCurvature{zero, zero, AssignedCondition(one), AssignedCurvature{zero, AssignedCondition(one)}};
    }
    // ecc_chip.rs:531-578
    AssignedPoint bisec_point(const AssignedCondition& cond, const AssignedPoint& a, const AssignedPoint& b) {
        AssignedInteger x = base.bisec_int(cond, a.x, b.x);
        AssignedInteger y = base.bisec_int(cond, a.y, b.y);
        AssignedCondition z = bc().bisec_cond(cond, a.z, b.z);
        return AssignedPoint{x, y, z};
    }
    AssignedCurvature bisec_curvature(const AssignedCondition& cond, const AssignedCurvature& a, const AssignedCurvature& b) {
        AssignedInteger v = base.bisec_int(cond, a.v, b.v);
        AssignedCondition z = bc().bisec_cond(cond, a.z, b.z);
        return AssignedCurvature{v, z};
    }
    AssignedPointWithCurvature bisec_point_with_curvature(const AssignedCondition& cond, const AssignedPointWithCurvature& a,
                                                          const AssignedPointWithCurvature& b) {
        AssignedInteger x = base.bisec_int(cond, a.x, b.x);
        AssignedInteger y = base.bisec_int(cond, a.y, b.y);
        AssignedCondition z = bc().bisec_cond(cond, a.z, b.z);
        AssignedCurvature c = bisec_curvature(cond, a.curvature, b.curvature);
        return AssignedPointWithCurvature{x, y, z, c};
    }
    // ecc_chip.rs:580-604
    AssignedPoint lambda_to_point(const AssignedCurvature& lambda, const AssignedPoint& a, const AssignedPoint& b) {
        const AssignedInteger& l = lambda.v;
        AssignedInteger l_square = base.int_square(l);
        AssignedInteger t = base.int_sub(l_square, a.x);
        AssignedInteger cx = base.int_sub(t, b.x);
        t = base.int_sub(a.x, cx);
        t = base.int_mul(t, l);
        AssignedInteger cy = base.int_sub(t, a.y);
        return AssignedPoint{cx, cy, lambda.z};
    }
    // ecc_chip.rs:606-628
    AssignedPoint ecc_add(const AssignedPointWithCurvature& a, const AssignedPoint& b) {
        AssignedInteger diff_x = base.int_sub(a.x, b.x);
        AssignedInteger diff_y = base.int_sub(a.y, b.y);
        auto dv = base.int_div(diff_y, diff_x);
        AssignedCondition x_eq = dv.first;
        AssignedInteger tangent = dv.second;
        AssignedCondition y_eq = base.is_int_zero(diff_y);
        AssignedCondition eq = bc().and_(x_eq, y_eq);
        AssignedCurvature tg{tangent, x_eq};
        AssignedCurvature lambda = bisec_curvature(eq, a.curvature, tg);
        AssignedPoint a_p = a.to_point();
        AssignedPoint p = lambda_to_point(lambda, a_p, b);
        p = bisec_point(a.z, b, p);
        p = bisec_point(b.z, a_p, p);
        return p;
    }
    // ecc_chip.rs:630-642
    AssignedPoint ecc_double(const AssignedPointWithCurvature& a) {
        ORC_ASSERT(!(curve.scalar_mod - BN(1)).bit(0));
        AssignedPoint a_p = a.to_point();
        AssignedPoint p = lambda_to_point(a.curvature, a_p, a_p);
        p.z = bc().bisec_cond(a.z, a.z, p.z);
        return p;
    }
    // ecc_chip.rs:644-658
    void ecc_assert_equal(const AssignedPoint& a, const AssignedPoint& b) {
        AssignedCondition eq_x = base.is_int_equal(a.x, b.x);
        AssignedCondition eq_y = base.is_int_equal(a.y, b.y);
        AssignedCondition eq_z = bc().xnor(a.z, b.z);
        AssignedCondition eq_xy = bc().and_(eq_x, eq_y);
        AssignedCondition eq_xyz = bc().and_(eq_xy, eq_z);
        AssignedCondition is_both_identity = bc().and_(a.z, b.z);
        AssignedCondition eq = bc().or_(eq_xyz, is_both_identity);
        bc().assert_true(eq);
    }
    // ecc_chip.rs:660-666
    AssignedPoint ecc_neg(const AssignedPoint& a) { return AssignedPoint{a.x, base.int_neg(a.y), a.z}; }
    // ecc_chip.rs:668-675
    AssignedPoint ecc_reduce(const AssignedPoint& a) {
        AssignedInteger x = base.reduce(a.x);
        AssignedInteger y = base.reduce(a.y);
        AssignedCondition z = a.z;
        AssignedPointWithCurvature identity = assign_identity();
        return bisec_point(z, identity.to_point(), AssignedPoint{x, y, z});
    }
    // ecc_chip.rs:677-693
    AssignedPointWithCurvature ecc_reduce_with_curvature(const AssignedPoint& a_in) {
        AssignedPoint a = ecc_reduce(a_in);
        AssignedInteger x_square = base.int_square(a.x);
        AssignedInteger numerator = base.int_mul_small_constant(x_square, 3);
        AssignedInteger denominator = base.int_mul_small_constant(a.y, 2);
        auto zv = base.int_div(numerator, denominator);
        AssignedInteger v = base.reduce(zv.second);
        return AssignedPointWithCurvature{a.x, a.y, a.z, AssignedCurvature{v, zv.first}};
    }
    // ecc_chip.rs:695-708
    AssignedPointWithCurvature to_point_with_curvature(const AssignedPoint& a) {
        AssignedInteger x_square = base.int_square(a.x);
        AssignedInteger numerator = base.int_mul_small_constant(x_square, 3);
        AssignedInteger denominator = base.int_mul_small_constant(a.y, 2);
        auto zv = base.int_div(numerator, denominator);
        return AssignedPointWithCurvature{a.x, a.y, a.z, AssignedCurvature{zv.second, zv.first}};
    }
    // ecc_chip.rs:710-732
    std::vector<AssignedValue> ecc_encode(const AssignedPoint& p_in) {
        AssignedPoint p = ecc_reduce(p_in);
        N shift = bn_to_n(bn_pow2(base.info->limb_bits));
        N one = n_from(1);
        AssignedValue s0 = bc().sum_with_constant({BaseOps::Elem(&p.x.limbs_le[0], one), BaseOps::Elem(&p.x.limbs_le[1], shift)}, nullptr);
        AssignedValue s1 = bc().sum_with_constant({BaseOps::Elem(&p.x.limbs_le[2], one), BaseOps::Elem(&p.y.limbs_le[0], shift)}, nullptr);
        AssignedValue s2 = bc().sum_with_constant({BaseOps::Elem(&p.y.limbs_le[1], one), BaseOps::Elem(&p.y.limbs_le[2], shift)}, nullptr);
        return {s0, s1, s2};
    }
    // ecc_chip.rs:734-777
    void assign_cache_integer(const AssignedInteger& p, size_t sc, size_t g, size_t& offset) {
        ORC_ASSERT(p.times == 1);
        for (size_t j = 0; j < base.info->limbs; j++) {
            base.assign_cache_value(p.limbs_le[j], offset, g, sc);
            offset += 1;
        }
        base.assign_cache_value(p.native, offset, g, sc);
        offset += 1;
    }
    AssignedInteger assign_selected_integer(const AssignedInteger& p, const AssignedValue& sc, size_t g, size_t& offset) {
        std::vector<AssignedValue> limbs_le;
        for (size_t j = 0; j < base.info->limbs; j++) {
            limbs_le.push_back(base.assign_selected_value(p.limbs_le[j], offset, g, sc));
            offset += 1;
        }
        AssignedValue native = base.assign_selected_value(p.native, offset, g, sc);
        offset += 1;
        return AssignedInteger(limbs_le, native, 1);
    }
    // ecc_chip.rs:779-812
    void assign_cache_point(const AssignedPointWithCurvature& p, size_t g, size_t sc) {
        size_t i = 0;
        assign_cache_integer(p.x, sc, g, i);
        assign_cache_integer(p.y, sc, g, i);
        base.assign_cache_value(p.z.v, i, g, sc);
        i += 1;
        assign_cache_integer(p.curvature.v, sc, g, i);
        base.assign_cache_value(p.curvature.z.v, i, g, sc);
    }
    AssignedPointWithCurvature assign_selected_point(const AssignedPointWithCurvature& p, const AssignedValue& sc, size_t g) {
        size_t i = 0;
        AssignedInteger x = assign_selected_integer(p.x, sc, g, i);
        AssignedInteger y = assign_selected_integer(p.y, sc, g, i);
        AssignedValue z = base.assign_selected_value(p.z.v, i, g, sc);
        i += 1;
        AssignedInteger c_v = assign_selected_integer(p.curvature.v, sc, g, i);
        AssignedValue c_z = base.assign_selected_value(p.curvature.z.v, i, g, sc);
        return AssignedPointWithCurvature{x, y, AssignedCondition(z), AssignedCurvature{c_v, AssignedCondition(c_z)}};
    }
    // ecc_chip.rs:814-838
    AssignedNonZeroPoint lambda_to_point_non_zero(const AssignedInteger& lambda, const AssignedNonZeroPoint& a,
                                                  const AssignedNonZeroPoint& b) {
        const AssignedInteger& l = lambda;
        AssignedInteger l_square = base.int_square(l);
        AssignedInteger t = base.int_sub(l_square, a.x);
        AssignedInteger cx = base.int_sub(t, b.x);
        t = base.int_sub(a.x, cx);
        t = base.int_mul(t, l);
        AssignedInteger cy = base.int_sub(t, a.y);
        return AssignedNonZeroPoint{cx, cy};
    }
    // ecc_chip.rs:840-858
    AssignedNonZeroPoint ecc_add_unsafe(const AssignedNonZeroPoint& a, const AssignedNonZeroPoint& b) {
        AssignedInteger diff_x = base.int_sub(a.x, b.x);
        AssignedInteger diff_y = base.int_sub(a.y, b.y);
        auto dv = base.int_div(diff_y, diff_x);
        bool succeed = bc().try_assert_false(dv.first);
        AssignedNonZeroPoint res = lambda_to_point_non_zero(dv.second, a, b);
        if (!succeed) throw UnsafeError{1};
        return res;
    }
    // ecc_chip.rs:860-882
    AssignedNonZeroPoint ecc_double_unsafe(const AssignedNonZeroPoint& a) {
        AssignedInteger x_square = base.int_square(a.x);
        AssignedInteger numerator = base.int_mul_small_constant(x_square, 3);
        AssignedInteger denominator = base.int_mul_small_constant(a.y, 2);
        auto zv = base.int_div(numerator, denominator);
        bool succeed = bc().try_assert_false(zv.first);
        AssignedNonZeroPoint res = lambda_to_point_non_zero(zv.second, a, a);
        if (!succeed) throw UnsafeError{2};
        return res;
    }
    // ecc_chip.rs:884-911
    AssignedNonZeroPoint ecc_neg_non_zero(const AssignedNonZeroPoint& a) { return AssignedNonZeroPoint{a.x, base.int_neg(a.y)}; }
    AssignedNonZeroPoint ecc_reduce_non_zero(const AssignedNonZeroPoint& a) {
        AssignedInteger x = base.reduce(a.x);
        AssignedInteger y = base.reduce(a.y);
        return AssignedNonZeroPoint{x, y};
    }
    AssignedNonZeroPoint ecc_bisec_non_zero_point(const AssignedCondition& cond, const AssignedNonZeroPoint& a,
                                                  const AssignedNonZeroPoint& b) {
        AssignedInteger x = base.bisec_int(cond, a.x, b.x);
        AssignedInteger y = base.bisec_int(cond, a.y, b.y);
        return AssignedNonZeroPoint{x, y};
    }
    // ecc_chip.rs:913-933
    AssignedNonZeroPoint bisec_candidate_non_zero(const std::vector<AssignedNonZeroPoint>& candidates,
                                                  const std::vector<AssignedCondition>& group_bits) {
        std::vector<AssignedNonZeroPoint> curr = candidates;
        for (auto& bit : group_bits) {
            std::vector<AssignedNonZeroPoint> next;
            for (size_t k = 0; k + 1 < curr.size() + 1 && k < curr.size(); k += 2) {
                ORC_ASSERT(k + 1 < curr.size());
                next.push_back(ecc_bisec_non_zero_point(bit, curr[k + 1], curr[k]));
            }
            curr = next;
        }
        ORC_ASSERT(curr.size() == 1);
        return curr[0];
    }
    // ecc_chip.rs:935-953
    std::pair<AssignedValue, AssignedNonZeroPoint> pick_candidate_non_zero(const std::vector<AssignedNonZeroPoint>& candidates,
                                                                           const std::vector<AssignedCondition>& group_bits) {
        std::vector<BaseOps::Elem> index_vec;
        for (size_t i = 0; i < group_bits.size(); i++) index_vec.push_back(BaseOps::Elem(&group_bits[i].v, n_from(1ull << i)));
        AssignedValue index = bc().sum_with_constant(index_vec, nullptr);
        size_t index_i = (size_t)(index.val.w[0] & 0xff);  // byte 0 of the repr (ecc_chip.rs:949)
        return {index, candidates.at(index_i)};
    }
    // ecc_chip.rs:955-973
    AssignedNonZeroPoint assign_selected_point_non_zero(const AssignedNonZeroPoint& p, const AssignedValue& sc, size_t g) {
        size_t i = 0;
        AssignedInteger x = assign_selected_integer(p.x, sc, g, i);
        AssignedInteger y = assign_selected_integer(p.y, sc, g, i);
        return AssignedNonZeroPoint{x, y};
    }
    void assign_cache_point_non_zero(const AssignedNonZeroPoint& p, size_t g, size_t sc) {
        size_t i = 0;
        assign_cache_integer(p.x, sc, g, i);
        assign_cache_integer(p.y, sc, g, i);
    }
    // ecc_chip.rs:975-1008
    void ecc_assert_equal_non_zero(const AssignedNonZeroPoint& a, const AssignedNonZeroPoint& b) {
        base.assert_int_equal(a.x, b.x);
        base.assert_int_equal(a.y, b.y);
    }
    AssignedPoint ecc_non_zero_point_downgrade(const AssignedNonZeroPoint& a) {
        AssignedValue zero = bc().assign_constant(n_from(0));
        return AssignedPoint{a.x, a.y, AssignedCondition(zero)};
    }
    AssignedNonZeroPoint ecc_bisec_to_non_zero_point(const AssignedPoint& a, const AssignedNonZeroPoint& b) {
        AssignedInteger x = base.bisec_int(a.z, b.x, a.x);
        AssignedInteger y = base.bisec_int(a.z, b.y, a.y);
        return AssignedNonZeroPoint{x, y};
    }

    // ---- EccChipScalarOps ----
    size_t get_and_increase_msm_prefix() {
        size_t ret = msm_prefix;
        ORC_ASSERT(ret < MSM_LIMIT);
        msm_prefix += MSM_PREFIX_OFFSET;
        return ret;
    }
    // native_scalar_ecc_chip.rs:97-171 / general_scalar_ecc_chip.rs:96-147, WINDOW_SIZE = 1
    std::vector<AssignedCondition> decompose_scalar(const AssignedScalar& s) {
        std::vector<AssignedCondition> bits;
        if (native_scalar) {
            N one = n_from(1), two = n_from(2), four = n_from(4);
            BN s_bn = s.v.val;
            AssignedValue v = s.v;
            for (uint64_t i = 0; i < curve.scalar_num_bits / 2; i++) {
                AssignedCondition b0 = bc().assign_bit(n_from(s_bn.bit(i * 2) ? 1 : 0));
                AssignedCondition b1 = bc().assign_bit(n_from(s_bn.bit(i * 2 + 1) ? 1 : 0));
                N v_next = bn_to_n(s_bn >> (i * 2 + 2));
                auto cells = bc().one_line_with_last({Pair(ValueSchema(v_next), four), Pair(&b1.v, two), Pair(&b0.v, one)},
                                                     Pair(&v, n_neg(one)), nullptr, {}, nullptr);
                v = cells.first[0];
                bits.push_back(b0);
                bits.push_back(b1);
            }
            if (curve.scalar_num_bits % 2 == 1) {
                bc().assert_bit(v);
                bits.push_back(AssignedCondition(v));
            } else {
                bc().assert_constant(v, n_from(0));
            }
            // WINDOW_SIZE == 1: no padding
        } else {
            N zero = n_from(0), one = n_from(1), two = n_from(2);
            N two_inv;
            ORC_ASSERT(n_inv(two, two_inv));
            AssignedInteger sr = scalar->reduce(s.i);
            for (auto& l : sr.limbs_le) {
                BN v = l.val;
                AssignedValue rest = l;
                for (uint64_t j = 0; j < scalar->info->limb_bits; j++) {
                    AssignedCondition b = bc().assign_bit(n_from(v.bit(j) ? 1 : 0));
                    N nv = n_mul(n_sub(rest.val, b.v.val), two_inv);
                    rest = bc().one_line_with_last({Pair(&rest, n_neg(one)), Pair(&b.v, one)}, Pair(ValueSchema(nv), two), nullptr, {}, nullptr)
                               .second;
                    bits.push_back(b);
                }
                bc().assert_constant(rest, zero);
            }
        }
        std::vector<AssignedCondition> res(bits.rbegin(), bits.rend());  // chunks(1) then reverse
        return res;
    }
    AssignedScalar ecc_bisec_scalar(const AssignedCondition& cond, const AssignedScalar& a, const AssignedScalar& b) {
        AssignedScalar r;
        if (native_scalar)
            r.v = bc().bisec(cond, a.v, b.v);
        else
            r.i = scalar->bisec_int(cond, a.i, b.i);
        return r;
    }
    AssignedScalar ecc_assign_constant_zero_scalar() {
        AssignedScalar r;
        if (native_scalar)
            r.v = bc().assign_constant(n_from(0));
        else
            r.i = scalar->assign_int_constant(BN(0));
        return r;
    }

    // ecc_chip.rs:91-221 and 223-371 share everything but candidate selection.
    AssignedPoint msm_batch_on_group_non_zero(const std::vector<AssignedNonZeroPoint>& points_in, const std::vector<AssignedScalar>& scalars,
                                              const HostPoint& rand_acc_point_h, const HostPoint& rand_line_point_h, bool with_select) {
        if (with_select) ORC_ASSERT(points_in.size() <= MSM_PREFIX_OFFSET);
        std::vector<AssignedNonZeroPoint> points;
        for (auto& p : points_in) points.push_back(ecc_reduce_non_zero(p));
        AssignedNonZeroPoint rand_acc_point = assign_non_zero_point(rand_acc_point_h);
        AssignedNonZeroPoint rand_line_point = assign_non_zero_point(rand_line_point_h);
        AssignedNonZeroPoint rand_acc_point_neg = ecc_reduce_non_zero(ecc_neg_non_zero(rand_acc_point));
        AssignedNonZeroPoint rand_line_point_neg = ecc_reduce_non_zero(ecc_neg_non_zero(rand_line_point));

        size_t best_group_size = with_select ? 5 : 2;
        size_t n_group = (points.size() + best_group_size - 1) / best_group_size;
        size_t group_size = (points.size() + n_group - 1) / n_group;

        std::vector<std::vector<AssignedNonZeroPoint>> candidates;
        size_t group_prefix = with_select ? get_and_increase_msm_prefix() : 0;
        size_t n_groups = (points.size() + group_size - 1) / group_size;
        for (size_t group_index = 0; group_index < n_groups; group_index++) {
            size_t c0 = group_index * group_size, c1 = std::min(points.size(), c0 + group_size);
            const AssignedNonZeroPoint& init = (group_index % 2 == 0) ? rand_line_point : rand_line_point_neg;
            candidates.push_back({init});
            if (with_select) assign_cache_point_non_zero(init, group_prefix + group_index, 0);
            std::vector<AssignedNonZeroPoint>& cl = candidates.back();
            for (uint32_t i = 1; i < (1u << (c1 - c0)); i++) {
                uint32_t pos = __builtin_ctz(i);  // i.reverse_bits().leading_zeros()
                uint32_t other = i - (1u << pos);
                AssignedNonZeroPoint p = ecc_add_unsafe(cl[other], points[c0 + pos]);
                p = ecc_reduce_non_zero(p);
                if (with_select) assign_cache_point_non_zero(p, group_prefix + group_index, i);
                cl.push_back(p);
            }
        }

        std::vector<std::vector<AssignedCondition>> bits;
        for (auto& s : scalars) bits.push_back(decompose_scalar(s));
        size_t windows = bits[0].size();

        auto window_body = [&](EccContext& ops, size_t wi) {
            AssignedNonZeroPoint acc = rand_acc_point_neg;
            for (size_t group_index = 0; group_index < n_groups; group_index++) {
                size_t c0 = group_index * group_size, c1 = std::min(points.size(), c0 + group_size);
                std::vector<AssignedCondition> group_bits;
                for (size_t k = c0; k < c1; k++) group_bits.push_back(bits[k][wi]);
                AssignedNonZeroPoint ci;
                if (with_select) {
                    auto pk = ops.pick_candidate_non_zero(candidates[group_index], group_bits);
                    ci = ops.assign_selected_point_non_zero(pk.second, pk.first, group_index + group_prefix);
                } else {
                    ci = ops.bisec_candidate_non_zero(candidates[group_index], group_bits);
                }
                acc = ops.ecc_add_unsafe(ci, acc);
            }
            return acc;
        };

        EccContext predict_ops = clone_without_offset();
        Offset offset_before = predict_ops.offset();
        std::vector<AssignedNonZeroPoint> line_acc_arr;
        line_acc_arr.push_back(window_body(predict_ops, 0));
        Offset offset_after = predict_ops.offset();
        Offset offset_diff = offset_after - offset_before;
        merge(predict_ops);

        // windows 1.. run on clones placed at offset_diff * i (rayon par_iter in the reference;
        // sequential here -- rows are disjoint and permutations are merged in window order).
        std::vector<EccContext> cloned_ops;
        for (size_t i = 1; i < windows; i++) cloned_ops.push_back(clone_with_offset(offset_diff.scale(i)));
        for (size_t i = 1; i < windows; i++) {
            EccContext& op = cloned_ops[i - 1];
            Offset ob = op.offset();
            line_acc_arr.push_back(window_body(op, i));
            ORC_ASSERT(offset_diff == (op.offset() - ob));
        }
        for (auto& op : cloned_ops) merge(op);
        apply_offset_diff(offset_diff.scale(windows));

        AssignedNonZeroPoint acc = rand_acc_point;
        for (size_t wi = 0; wi < windows; wi++) {
            acc = ecc_double_unsafe(acc);
            acc = ecc_add_unsafe(line_acc_arr[wi], acc);
            if (n_groups % 2 == 1) acc = ecc_add_unsafe(acc, rand_line_point_neg);
        }
        AssignedPoint accp = ecc_non_zero_point_downgrade(acc);
        AssignedPointWithCurvature accc = to_point_with_curvature(accp);
        AssignedPoint carry = ecc_non_zero_point_downgrade(rand_acc_point_neg);
        return ecc_add(accc, carry);
    }

    // ecc_chip.rs:373-408 with r1, r2 given
    AssignedPoint msm_unsafe(const std::vector<AssignedPoint>& points, const std::vector<AssignedScalar>& scalars, const HostPoint& r1,
                             const HostPoint& r2) {
        std::vector<AssignedNonZeroPoint> non_zero_points;
        std::vector<AssignedScalar> normalized_scalars;
        AssignedNonZeroPoint non_zero_p = assign_non_zero_point(curve.gen);
        AssignedScalar s_zero = ecc_assign_constant_zero_scalar();
        for (size_t i = 0; i < points.size(); i++) {
            AssignedScalar s = ecc_bisec_scalar(points[i].z, s_zero, scalars[i]);
            AssignedNonZeroPoint p = ecc_bisec_to_non_zero_point(points[i], non_zero_p);
            non_zero_points.push_back(p);
            normalized_scalars.push_back(s);
        }
        return msm_batch_on_group_non_zero(non_zero_points, normalized_scalars, r1, r2, has_select_chip());
    }
};

}  // namespace orc
