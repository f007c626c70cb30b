// Symbolic mirror of the reference's ECC / MSM chips (see tracer.h for the model).
//   EccChipBaseOps / EccChipScalarOps / ParallelClone / Offset   src/circuit/ecc_chip.rs:23-1009
//   NativeScalarEccContext                                       src/circuit/native_scalar_ecc_chip.rs:27-194
//   GeneralScalarEccContext                                      src/circuit/general_scalar_ecc_chip.rs:25-169
// Points enter as per-instance input cells. msm_unsafe's random blinding points (ecc_chip.rs:378-379)
// are per-instance inputs too; an "unsafe" add hitting equal x-coordinates sets the instance's
// status bit (UnsafeError) instead of aborting the trace.
#pragma once
#include "tracer.h"

namespace h2e {

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
struct AssignedScalar {  // AssignedValue (native scalar) or AssignedInteger (general scalar)
    AssignedValue v;
    AssignedInteger i;
};
struct PointInput {  // per-instance input cells of an affine point (64-byte logical inputs)
    uint32_t x_cell, y_cell;
};

struct Offset {  // ecc_chip.rs:36-62
    size_t range = 0, base = 0, select = 0;
    Offset operator-(const Offset& r) const { return Offset{range - r.range, base - r.base, select - r.select}; }
    Offset scale(size_t n) const { return Offset{range * n, base * n, select * n}; }
    bool operator==(const Offset& r) const { return range == r.range && base == r.base && select == r.select; }
};

struct CurveInfo {
    Field base_field;
    Field scalar_field;  // general scalar only
    Big b;
    Big gen_x, gen_y;
    unsigned scalar_num_bits;
};
inline CurveInfo curve_bn256_g1() { return CurveInfo{F_BN256_FQ, F_BLS12_381_FR, Big(3), Big(1), Big(2), 254}; }
inline CurveInfo curve_bls12_381_g1() {
    return CurveInfo{F_BLS12_381_FQ, F_BLS12_381_FR, Big(4),
                     Big::from_hex("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"),
                     Big::from_hex("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1"), 255};
}

// NativeScalarEccContext<C> (context.rs:190-213) and GeneralScalarEccContext<C,N> (215-239).
class EccContext {
   public:
    Context* ctx;
    IntegerContext base;
    IntegerContext scalar;  // only meaningful when !native_scalar
    size_t msm_prefix;
    bool native_scalar;
    CurveInfo curve;

    EccContext(Context* c, const CurveInfo& ci, bool native, bool with_select = true)
        : ctx(c), base(c, ci.base_field), scalar(c, ci.scalar_field), msm_prefix(with_select ? 0 : (size_t)-1), native_scalar(native), curve(ci) {}

    bool has_select_chip() const { return native_scalar ? msm_prefix < (size_t)-1 : true; }
    unsigned L() const { return base.L(); }

    // ---- EccChipBaseOps ----
    // ecc_chip.rs:458-487. `identity_cell`: input cell holding z (0/1).
    AssignedPoint assign_point(const PointInput& p, uint32_t identity_cell) {
        AssignedInteger x = base.assign_w(p.x_cell);
        AssignedInteger y = base.assign_w(p.y_cell);
        AssignedCondition z = ctx->assign_bit(identity_cell);
        AssignedInteger b = base.assign_int_constant(curve.b);
        AssignedInteger y2 = base.int_square(y);
        AssignedInteger x2 = base.int_square(x);
        AssignedInteger x3 = base.int_mul(x2, x);
        AssignedInteger right = base.int_add(x3, b);
        AssignedCondition eq = base.is_int_equal(y2, right);
        AssignedCondition eq_or_identity = ctx->or_(eq, z);
        ctx->assert_true(eq_or_identity);
        return AssignedPoint{x, y, z};
    }
    // ecc_chip.rs:489-512
    AssignedNonZeroPoint assign_non_zero_point(const PointInput& p) {
        AssignedInteger x = base.assign_w(p.x_cell);
        AssignedInteger y = base.assign_w(p.y_cell);
        return finish_non_zero_point(x, y);
    }
    // same for a shape-level constant point (the generator in msm_unsafe, ecc_chip.rs:383)
    AssignedNonZeroPoint assign_non_zero_point_static(const Big& xv, const Big& yv) {
        AssignedInteger x = base.assign_w_static(xv);
        AssignedInteger y = base.assign_w_static(yv);
        return finish_non_zero_point(x, y);
    }
    AssignedNonZeroPoint finish_non_zero_point(const AssignedInteger& x, const AssignedInteger& y) {
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
        AssignedInteger zero = base.assign_int_constant(Big(0));
        AssignedValue one = ctx->assign_constant(Big(1));
        return AssignedPointWithCurvature{zero, zero, AssignedCondition{one}, AssignedCurvature{zero, AssignedCondition{one}}};
    }
    // ecc_chip.rs:531-560
    AssignedPoint bisec_point(const AssignedCondition& cond, const AssignedPoint& a, const AssignedPoint& b) {
        AssignedInteger x = base.bisec_int(cond, a.x, b.x);
        AssignedInteger y = base.bisec_int(cond, a.y, b.y);
        AssignedCondition z = ctx->bisec_cond(cond, a.z, b.z);
        return AssignedPoint{x, y, z};
    }
    AssignedCurvature bisec_curvature(const AssignedCondition& cond, const AssignedCurvature& a, const AssignedCurvature& b) {
        AssignedInteger v = base.bisec_int(cond, a.v, b.v);
        AssignedCondition z = ctx->bisec_cond(cond, a.z, b.z);
        return AssignedCurvature{v, z};
    }
    // ecc_chip.rs:580-604
    AssignedPoint lambda_to_point(const AssignedCurvature& lambda, const AssignedPoint& a, const AssignedPoint& b) {
        AssignedInteger l_square = base.int_square(lambda.v);
        AssignedInteger t = base.int_sub(l_square, a.x);
        AssignedInteger cx = base.int_sub(t, b.x);
        t = base.int_sub(a.x, cx);
        t = base.int_mul(t, lambda.v);
        AssignedInteger cy = base.int_sub(t, a.y);
        return AssignedPoint{cx, cy, lambda.z};
    }
    // ecc_chip.rs:606-628
    AssignedPoint ecc_add(const AssignedPointWithCurvature& a, const AssignedPoint& b) {
        AssignedInteger diff_x = base.int_sub(a.x, b.x);
        AssignedInteger diff_y = base.int_sub(a.y, b.y);
        auto dv = base.int_div(diff_y, diff_x);
        AssignedCondition y_eq = base.is_int_zero(diff_y);
        AssignedCondition eq = ctx->and_(dv.first, y_eq);
        AssignedCurvature tg{dv.second, dv.first};
        AssignedCurvature lambda = bisec_curvature(eq, a.curvature, tg);
        AssignedPoint a_p = a.to_point();
        AssignedPoint p = lambda_to_point(lambda, a_p, b);
        p = bisec_point(a.z, b, p);
        p = bisec_point(b.z, a_p, p);
        return p;
    }
    // ecc_chip.rs:630-642
    AssignedPoint ecc_double(const AssignedPointWithCurvature& a) {
        AssignedPoint a_p = a.to_point();
        AssignedPoint p = lambda_to_point(a.curvature, a_p, a_p);
        p.z = ctx->bisec_cond(a.z, a.z, p.z);
        return p;
    }
    // ecc_chip.rs:644-658
    void ecc_assert_equal(const AssignedPoint& a, const AssignedPoint& b) {
        AssignedCondition eq_x = base.is_int_equal(a.x, b.x);
        AssignedCondition eq_y = base.is_int_equal(a.y, b.y);
        AssignedCondition eq_z = ctx->xnor(a.z, b.z);
        AssignedCondition eq_xy = ctx->and_(eq_x, eq_y);
        AssignedCondition eq_xyz = ctx->and_(eq_xy, eq_z);
        AssignedCondition is_both_identity = ctx->and_(a.z, b.z);
        AssignedCondition eq = ctx->or_(eq_xyz, is_both_identity);
        ctx->assert_true(eq);
    }
    // ecc_chip.rs:660-708
    AssignedPoint ecc_neg(const AssignedPoint& a) { return AssignedPoint{a.x, base.int_neg(a.y), a.z}; }
    AssignedPoint ecc_reduce(const AssignedPoint& a) {
        AssignedInteger x = base.reduce(a.x);
        AssignedInteger y = base.reduce(a.y);
        AssignedPointWithCurvature identity = assign_identity();
        return bisec_point(a.z, identity.to_point(), AssignedPoint{x, y, a.z});
    }
    AssignedPointWithCurvature to_point_with_curvature(const AssignedPoint& a) {
        AssignedInteger x_square = base.int_square(a.x);
        AssignedInteger numerator = base.int_mul_small_constant(x_square, 3);
        AssignedInteger denominator = base.int_mul_small_constant(a.y, 2);
        auto zv = base.int_div(numerator, denominator);
        return AssignedPointWithCurvature{a.x, a.y, a.z, AssignedCurvature{zv.second, zv.first}};
    }
    AssignedPointWithCurvature ecc_reduce_with_curvature(const AssignedPoint& a_in) {
        AssignedPoint a = ecc_reduce(a_in);
        AssignedInteger x_square = base.int_square(a.x);
        AssignedInteger numerator = base.int_mul_small_constant(x_square, 3);
        AssignedInteger denominator = base.int_mul_small_constant(a.y, 2);
        auto zv = base.int_div(numerator, denominator);
        AssignedInteger v = base.reduce(zv.second);
        return AssignedPointWithCurvature{a.x, a.y, a.z, AssignedCurvature{v, zv.first}};
    }
    // ecc_chip.rs:710-732
    std::vector<AssignedValue> ecc_encode(const AssignedPoint& p_in) {
        AssignedPoint p = ecc_reduce(p_in);
        Big shift = Big::pow2(base.info->limb_bits) % native_modulus();
        AssignedValue s0 = ctx->sum_with_constant({Context::Elem(&p.x.limbs_le[0], ctx->ONE), Context::Elem(&p.x.limbs_le[1], shift)}, nullptr);
        AssignedValue s1 = ctx->sum_with_constant({Context::Elem(&p.x.limbs_le[2], ctx->ONE), Context::Elem(&p.y.limbs_le[0], shift)}, nullptr);
        AssignedValue s2 = ctx->sum_with_constant({Context::Elem(&p.y.limbs_le[1], ctx->ONE), Context::Elem(&p.y.limbs_le[2], shift)}, nullptr);
        return {s0, s1, s2};
    }
    // select_chip.rs:118-122
    static Big encode_offset(size_t g, size_t offset, size_t limb_offset) {
        return (Big(offset) << 128) + (Big(g) << 64) + Big(limb_offset);
    }
    // ecc_chip.rs:734-751 (one macro-op per integer: L+1 cache rows)
    void assign_cache_integer(const AssignedInteger& p, size_t sc, size_t g, size_t& offset) {
        if (p.times != 1) throw std::logic_error("assign_cache_integer: times != 1");
        Instr in = Context::mk(OP_CACHE_INT, base.field);
        base.put_int(in, 0, p, true);
        Context::Macro m(*ctx, in);
        for (unsigned j = 0; j < L(); j++) {
            ctx->assign_cache_value(p.limbs_le[j], encode_offset(g, sc, offset));
            offset += 1;
        }
        ctx->assign_cache_value(p.native, encode_offset(g, sc, offset));
        offset += 1;
    }
    // ecc_chip.rs:753-777 with the value-dependent choice of pick_candidate_non_zero (935-953)
    // folded in: the device reads the index cell and copies that candidate's cells.
    AssignedInteger assign_selected_integer(const std::vector<const AssignedInteger*>& candidates, const AssignedValue& sc, size_t g,
                                            size_t& offset) {
        Instr in = Context::mk(OP_SELECT_INT, base.field);
        in.a[0] = sc.slot;
        in.a[1] = (uint32_t)ctx->shape.tables.size();
        in.a[2] = (uint32_t)candidates.size();
        for (auto* c : candidates) {
            for (unsigned j = 0; j < L(); j++) ctx->shape.tables.push_back(c->limbs_le[j].slot);
            ctx->shape.tables.push_back(c->native.slot);
        }
        Context::Macro m(*ctx, in);
        AssignedInteger r;
        for (unsigned j = 0; j < L(); j++) {
            r.limbs_le.push_back(ctx->assign_select_value(encode_offset(g, 0, offset), sc));
            offset += 1;
        }
        r.native = ctx->assign_select_value(encode_offset(g, 0, offset), sc);
        offset += 1;
        r.times = 1;
        return r;
    }
    // ecc_chip.rs:814-882
    AssignedNonZeroPoint lambda_to_point_non_zero(const AssignedInteger& l, const AssignedNonZeroPoint& a, const AssignedNonZeroPoint& b) {
        AssignedInteger l_square = base.int_square(l);
        AssignedInteger t = base.int_sub(l_square, a.x);
        AssignedInteger cx = base.int_sub(t, b.x);
        t = base.int_sub(a.x, cx);
        t = base.int_mul(t, l);
        AssignedInteger cy = base.int_sub(t, a.y);
        return AssignedNonZeroPoint{cx, cy};
    }
    AssignedNonZeroPoint ecc_add_unsafe(const AssignedNonZeroPoint& a, const AssignedNonZeroPoint& b) {
        AssignedInteger diff_x = base.int_sub(a.x, b.x);
        AssignedInteger diff_y = base.int_sub(a.y, b.y);
        auto dv = base.int_div(diff_y, diff_x);
        ctx->try_assert_false(dv.first, ST_ADD_SAME_OR_NEG);
        return lambda_to_point_non_zero(dv.second, a, b);
    }
    AssignedNonZeroPoint ecc_double_unsafe(const AssignedNonZeroPoint& a) {
        AssignedInteger x_square = base.int_square(a.x);
        AssignedInteger numerator = base.int_mul_small_constant(x_square, 3);
        AssignedInteger denominator = base.int_mul_small_constant(a.y, 2);
        auto zv = base.int_div(numerator, denominator);
        ctx->try_assert_false(zv.first, ST_ADD_IDENTITY);
        return lambda_to_point_non_zero(zv.second, a, a);
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
            for (size_t k = 0; k + 1 < curr.size(); k += 2) next.push_back(ecc_bisec_non_zero_point(bit, curr[k + 1], curr[k]));
            curr = next;
        }
        if (curr.size() != 1) throw std::logic_error("bisec_candidate_non_zero");
        return curr[0];
    }
    // ecc_chip.rs:935-967: index = sum(bits * 2^i); the selected point's cells are copies of
    // candidates[index] (no permutation recorded for the value, context.rs:769-801).
    AssignedNonZeroPoint pick_and_select_candidate_non_zero(const std::vector<AssignedNonZeroPoint>& candidates,
                                                            const std::vector<AssignedCondition>& group_bits, size_t g) {
        std::vector<Context::Elem> index_vec;
        for (size_t i = 0; i < group_bits.size(); i++) index_vec.push_back(Context::Elem(&group_bits[i].v, Big(1ull << i)));
        AssignedValue index = ctx->sum_with_constant(index_vec, nullptr);
        std::vector<const AssignedInteger*> xs, ys;
        for (auto& c : candidates) {
            xs.push_back(&c.x);
            ys.push_back(&c.y);
        }
        size_t i = 0;
        AssignedInteger x = assign_selected_integer(xs, index, g, i);
        AssignedInteger y = assign_selected_integer(ys, index, g, i);
        return AssignedNonZeroPoint{x, y};
    }
    void assign_cache_point_non_zero(const AssignedNonZeroPoint& p, size_t g, size_t sc) {  // ecc_chip.rs:969-973
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
        AssignedValue zero = ctx->assign_constant(Big(0));
        return AssignedPoint{a.x, a.y, AssignedCondition{zero}};
    }
    AssignedNonZeroPoint ecc_bisec_to_non_zero_point(const AssignedPoint& a, const AssignedNonZeroPoint& b) {
        AssignedInteger x = base.bisec_int(a.z, b.x, a.x);
        AssignedInteger y = base.bisec_int(a.z, b.y, a.y);
        return AssignedNonZeroPoint{x, y};
    }

    // ---- EccChipScalarOps ----
    size_t get_and_increase_msm_prefix() {
        size_t ret = msm_prefix;
        if (ret >= MSM_LIMIT) throw std::logic_error("msm prefix overflow");
        msm_prefix += MSM_PREFIX_OFFSET;
        return ret;
    }
    // decompose_scalar::<1> (native_scalar_ecc_chip.rs:97-171, general_scalar_ecc_chip.rs:96-147):
    // returns bits most-significant window first.
    std::vector<AssignedCondition> decompose_scalar(const AssignedScalar& s) {
        std::vector<AssignedCondition> bits;
        if (native_scalar) {
            if (curve.scalar_num_bits % 2) throw std::logic_error("odd NUM_BITS not supported");
            Instr in = Context::mk(OP_DECOMPOSE_NATIVE);
            in.a[0] = s.v.slot;
            in.a[1] = curve.scalar_num_bits / 2;
            Context::Macro m(*ctx, in);
            AssignedValue v = s.v;
            for (unsigned i = 0; i < curve.scalar_num_bits / 2; i++) {
                AssignedCondition b0 = ctx->assign_bit_row();
                AssignedCondition b1 = ctx->assign_bit_row();
                std::vector<AssignedValue> cells;
                ctx->one_line_with_last({Pair(ValueSchema(), Big(4)), Pair(&b1.v, Big(2)), Pair(&b0.v, ctx->ONE)}, Pair(&v, ctx->NEG_ONE), nullptr,
                                        {}, nullptr, &cells);
                v = cells[0];
                bits.push_back(b0);
                bits.push_back(b1);
            }
            ctx->assert_constant(v, ctx->ZERO);
        } else {
            AssignedInteger sr = scalar.reduce(s.i);
            Big two_inv = (native_modulus() + Big(1)) >> 1;
            (void)two_inv;
            for (auto& l : sr.limbs_le) {
                Instr in = Context::mk(OP_DECOMPOSE_LIMB);
                in.a[0] = l.slot;
                in.a[1] = scalar.info->limb_bits;
                Context::Macro m(*ctx, in);
                AssignedValue rest = l;
                for (unsigned j = 0; j < scalar.info->limb_bits; j++) {
                    AssignedCondition b = ctx->assign_bit_row();
                    rest = ctx->one_line_with_last({Pair(&rest, ctx->NEG_ONE), Pair(&b.v, ctx->ONE)}, Pair(ValueSchema(), Big(2)), nullptr, {},
                                                   nullptr);
                    bits.push_back(b);
                }
                ctx->assert_constant(rest, ctx->ZERO);
            }
        }
        return std::vector<AssignedCondition>(bits.rbegin(), bits.rend());
    }
    AssignedScalar ecc_bisec_scalar(const AssignedCondition& cond, const AssignedScalar& a, const AssignedScalar& b) {
        AssignedScalar r;
        if (native_scalar)
            r.v = ctx->bisec(cond, a.v, b.v);
        else
            r.i = scalar.bisec_int(cond, a.i, b.i);
        return r;
    }
    AssignedScalar ecc_assign_constant_zero_scalar() {
        AssignedScalar r;
        if (native_scalar)
            r.v = ctx->assign_constant(Big(0));
        else
            r.i = scalar.assign_int_constant(Big(0));
        return r;
    }

    // ---- ParallelClone on heights/offsets (native_scalar_ecc_chip.rs:50-90) ----
    struct Fork {  // the part of a cloned Context that differs from its parent
        size_t height[3];
        size_t base_offset, range_offset, select_offset;
    };
    Fork snapshot() const {
        return Fork{{ctx->shape.height[0], ctx->shape.height[1], ctx->shape.height[2]}, ctx->base_offset, ctx->range_offset, ctx->select_offset};
    }
    void restore(const Fork& f) {
        for (int i = 0; i < 3; i++) ctx->shape.height[i] = f.height[i];
        ctx->base_offset = f.base_offset;
        ctx->range_offset = f.range_offset;
        ctx->select_offset = f.select_offset;
    }
    static void merge(Fork& self, const Fork& other) {
        self.height[0] = std::max(self.height[0], other.height[0]);
        self.height[1] = std::max(self.height[2], other.height[1]);  // sic: native_scalar_ecc_chip.rs:87
        self.height[2] = std::max(self.height[2], other.height[2]);
    }

    // ecc_chip.rs:91-221 and 223-371
    AssignedPoint msm_batch_on_group_non_zero(const std::vector<AssignedNonZeroPoint>& points_in, const std::vector<AssignedScalar>& scalars,
                                              const PointInput& r1, const PointInput& r2, bool with_select) {
        if (with_select && points_in.size() > MSM_PREFIX_OFFSET) throw std::logic_error("too many points");
        std::vector<AssignedNonZeroPoint> points;
        for (auto& p : points_in) points.push_back(ecc_reduce_non_zero(p));
        AssignedNonZeroPoint rand_acc_point = assign_non_zero_point(r1);
        AssignedNonZeroPoint rand_line_point = assign_non_zero_point(r2);
        AssignedNonZeroPoint rand_acc_point_neg = ecc_reduce_non_zero(ecc_neg_non_zero(rand_acc_point));
        AssignedNonZeroPoint rand_line_point_neg = ecc_reduce_non_zero(ecc_neg_non_zero(rand_line_point));

        size_t best_group_size = with_select ? 5 : 2;
        size_t n_group = (points.size() + best_group_size - 1) / best_group_size;
        size_t group_size = (points.size() + n_group - 1) / n_group;
        size_t n_groups = (points.size() + group_size - 1) / group_size;
        size_t group_prefix = with_select ? get_and_increase_msm_prefix() : 0;

        std::vector<std::vector<AssignedNonZeroPoint>> candidates;
        for (size_t gi = 0; gi < n_groups; gi++) {
            size_t c0 = gi * group_size, c1 = std::min(points.size(), c0 + group_size);
            const AssignedNonZeroPoint& init = (gi % 2 == 0) ? rand_line_point : rand_line_point_neg;
            candidates.push_back({init});
            if (with_select) assign_cache_point_non_zero(init, group_prefix + gi, 0);
            for (uint32_t i = 1; i < (1u << (c1 - c0)); i++) {
                uint32_t pos = __builtin_ctz(i);
                uint32_t other = i - (1u << pos);
                AssignedNonZeroPoint p = ecc_add_unsafe(candidates.back()[other], points[c0 + pos]);
                p = ecc_reduce_non_zero(p);
                if (with_select) assign_cache_point_non_zero(p, group_prefix + gi, i);
                candidates.back().push_back(p);
            }
        }
        std::vector<std::vector<AssignedCondition>> bits;
        for (auto& s : scalars) bits.push_back(decompose_scalar(s));
        size_t windows = bits[0].size();

        auto window_body = [&](size_t wi) {
            AssignedNonZeroPoint acc = rand_acc_point_neg;
            for (size_t gi = 0; gi < n_groups; gi++) {
                size_t c0 = gi * group_size, c1 = std::min(points.size(), c0 + group_size);
                std::vector<AssignedCondition> group_bits;
                for (size_t k = c0; k < c1; k++) group_bits.push_back(bits[k][wi]);
                AssignedNonZeroPoint ci = with_select ? pick_and_select_candidate_non_zero(candidates[gi], group_bits, gi + group_prefix)
                                                      : bisec_candidate_non_zero(candidates[gi], group_bits);
                acc = ecc_add_unsafe(ci, acc);
            }
            return acc;
        };
        // window 0 on a "predict" clone; windows 1.. on clones at offset_diff * i. All clones write
        // disjoint rows of the same store, so tracing them in order on one Context is equivalent;
        // only the height bookkeeping of clone/merge is replayed literally.
        Fork self = snapshot();
        std::vector<AssignedNonZeroPoint> line_acc_arr;
        line_acc_arr.push_back(window_body(0));
        Fork predict = snapshot();
        Offset offset_diff{predict.range_offset - self.range_offset, predict.base_offset - self.base_offset,
                           predict.select_offset - self.select_offset};
        merge(self, predict);
        for (size_t i = 1; i < windows; i++) {
            Fork clone = self;
            clone.base_offset += offset_diff.base * i;
            clone.range_offset += offset_diff.range * i;
            clone.select_offset += offset_diff.select * i;
            restore(clone);
            line_acc_arr.push_back(window_body(i));
            Fork after = snapshot();
            Offset d{after.range_offset - clone.range_offset, after.base_offset - clone.base_offset, after.select_offset - clone.select_offset};
            if (!(d == offset_diff)) throw std::logic_error("window offset diff mismatch (ecc_chip.rs:337-339)");
            merge(self, after);
        }
        self.base_offset += offset_diff.base * windows;
        self.range_offset += offset_diff.range * windows;
        self.select_offset += offset_diff.select * windows;
        restore(self);

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

    // ecc_chip.rs:373-408 with r1, r2 as inputs
    AssignedPoint msm_unsafe(const std::vector<AssignedPoint>& points, const std::vector<AssignedScalar>& scalars, const PointInput& r1,
                             const PointInput& r2) {
        std::vector<AssignedNonZeroPoint> non_zero_points;
        std::vector<AssignedScalar> normalized_scalars;
        AssignedNonZeroPoint non_zero_p = assign_non_zero_point_static(curve.gen_x, curve.gen_y);
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

}  // namespace h2e
