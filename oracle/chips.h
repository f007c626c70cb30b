// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the ops halves of the primitive chips and of the non-native integer chip:
//   BaseChipOps    src/circuit/base_chip.rs:81-605
//   RangeChipOps   src/circuit/range_chip.rs:262-348
//   SelectChipOps  src/circuit/select_chip.rs:100-162
//   IntegerChipOps src/circuit/integer_chip.rs:15-686
#pragma once
#include "core.h"

namespace orc {

// ------------------------------- BaseChipOps for Context -------------------------------------
struct BaseOps {
    Context* c;
    explicit BaseOps(Context* ctx) : c(ctx) {}

    int var_columns() const { return VAR_COLUMNS; }

    // base_chip.rs:516-541
    std::vector<AssignedValue> one_line(const std::vector<Pair>& pairs, const N* constant, const std::vector<N>& mul,
                                        const N* next) {
        std::vector<AssignedValue> res;
        for (size_t i = 0; i < pairs.size(); i++) res.push_back(AssignedValue(BaseChip, i, c->base_offset, pairs[i].first.val));
        c->records.one_line(c->base_offset, pairs, constant, mul, next);
        c->base_offset += 1;
        return res;
    }
    std::vector<AssignedValue> one_line_add(const std::vector<Pair>& pairs, const N* constant) {
        return one_line(pairs, constant, {}, nullptr);
    }
    // base_chip.rs:543-572
    std::pair<std::vector<AssignedValue>, AssignedValue> one_line_with_last(const std::vector<Pair>& pairs, const Pair& last,
                                                                            const N* constant, const std::vector<N>& mul,
                                                                            const N* next) {
        std::vector<AssignedValue> res0;
        for (size_t i = 0; i < pairs.size(); i++) res0.push_back(AssignedValue(BaseChip, i, c->base_offset, pairs[i].first.val));
        AssignedValue res1(BaseChip, VAR_COLUMNS - 1, c->base_offset, last.first.val);
        c->records.one_line_with_last(c->base_offset, pairs, last, constant, mul, next);
        c->base_offset += 1;
        return {res0, res1};
    }

    typedef std::pair<const AssignedValue*, N> Elem;

    // base_chip.rs:110-132
    AssignedValue sum_with_constant_in_one_line(const std::vector<Elem>& elems, const N* constant) {
        ORC_ASSERT(elems.size() < (size_t)var_columns());
        N sum = n_mul(elems[0].first->val, elems[0].second);
        for (size_t i = 1; i < elems.size(); i++) sum = n_add(sum, n_mul(elems[i].first->val, elems[i].second));
        if (constant) sum = n_add(*constant, sum);
        std::vector<Pair> pairs;
        for (auto& e : elems) pairs.push_back(Pair(ValueSchema(e.first), e.second));
        return one_line_with_last(pairs, Pair(ValueSchema(sum), n_neg(n_from(1))), constant, {}, nullptr).second;
    }
    // base_chip.rs:134-153
    AssignedValue sum_with_constant(const std::vector<Elem>& elems, const N* constant) {
        size_t columns = var_columns();
        if (elems.size() < columns) return sum_with_constant_in_one_line(elems, constant);
        std::vector<Elem> curr(elems.begin(), elems.begin() + (columns - 1));
        AssignedValue acc = sum_with_constant_in_one_line(curr, constant);
        for (size_t p = columns - 1; p < elems.size(); p += columns - 2) {
            size_t e = std::min(p + columns - 2, elems.size());
            std::vector<Elem> chunk(elems.begin() + p, elems.begin() + e);
            AssignedValue prev = acc;
            chunk.push_back(Elem(&prev, n_from(1)));
            acc = sum_with_constant_in_one_line(chunk, nullptr);
        }
        return acc;
    }
    // base_chip.rs:155-174
    AssignedValue add(const AssignedValue& a, const AssignedValue& b) {
        return sum_with_constant({Elem(&a, n_from(1)), Elem(&b, n_from(1))}, nullptr);
    }
    AssignedValue add_constant(const AssignedValue& a, const N& k) { return sum_with_constant({Elem(&a, n_from(1))}, &k); }
    AssignedValue sub(const AssignedValue& a, const AssignedValue& b) {
        return sum_with_constant({Elem(&a, n_from(1)), Elem(&b, n_neg(n_from(1)))}, nullptr);
    }
    // base_chip.rs:176-193
    AssignedValue mul(const AssignedValue& a, const AssignedValue& b) {
        N one = n_from(1), zero = n_from(0);
        N cv = n_mul(a.val, b.val);
        return one_line_with_last({Pair(&a, zero), Pair(&b, zero)}, Pair(ValueSchema(cv), n_neg(one)), nullptr, {one}, nullptr).second;
    }
    // base_chip.rs:195-217
    AssignedValue mul_add_constant(const AssignedValue& a, const AssignedValue& b, const N& k) {
        N one = n_from(1), zero = n_from(0);
        N d = n_add(n_mul(a.val, b.val), k);
        return one_line_with_last({Pair(&a, zero), Pair(&b, zero)}, Pair(ValueSchema(d), n_neg(one)), &k, {one}, nullptr).second;
    }
    // base_chip.rs:219-243
    AssignedValue mul_add(const AssignedValue& a, const AssignedValue& b, const N& ab_coeff, const AssignedValue& cc,
                          const N& c_coeff) {
        N one = n_from(1), zero = n_from(0);
        N d = n_add(n_mul(n_mul(a.val, b.val), ab_coeff), n_mul(cc.val, c_coeff));
        return one_line_with_last({Pair(&a, zero), Pair(&b, zero), Pair(&cc, c_coeff)}, Pair(ValueSchema(d), n_neg(one)), nullptr,
                                  {ab_coeff}, nullptr)
            .second;
    }
    struct MulAddTerm {
        const AssignedValue *a, *b, *c;
        N c_coeff;
    };
    // base_chip.rs:245-281
    AssignedValue mul_add_with_next_line(const std::vector<MulAddTerm>& ls) {
        ORC_ASSERT(ls.size() > 0);
        if (ls.size() == 1) return mul_add(*ls[0].a, *ls[0].b, n_from(1), *ls[0].c, ls[0].c_coeff);
        N one = n_from(1), zero = n_from(0), neg_one = n_neg(one);
        N t = zero;
        for (size_t i = 0; i < ls.size(); i++) {
            one_line_with_last({Pair(ls[i].a, zero), Pair(ls[i].b, zero), Pair(ls[i].c, ls[i].c_coeff)},
                               i == 0 ? Pair(ValueSchema(t), zero) : Pair(ValueSchema(t), one), nullptr, {one}, &neg_one);
            t = n_add(n_add(n_mul(ls[i].a->val, ls[i].b->val), n_mul(ls[i].c->val, ls[i].c_coeff)), t);
        }
        return one_line_with_last({}, Pair(ValueSchema(t), zero), nullptr, {}, nullptr).second;
    }
    // base_chip.rs:283-296
    AssignedValue invert_unsafe(const AssignedValue& a) {
        N b;
        ORC_ASSERT(n_inv(a.val, b));
        N one = n_from(1), zero = n_from(0), neg_one = n_neg(one);
        return one_line({Pair(&a, zero), Pair(ValueSchema(b), zero)}, &neg_one, {one}, nullptr)[1];
    }
    // base_chip.rs:298-321
    std::pair<AssignedCondition, AssignedValue> invert(const AssignedValue& a) {
        N zero = n_from(0), one = n_from(1), neg_one = n_neg(one);
        N b;
        if (!n_inv(a.val, b)) b = zero;
        N cv = n_sub(one, n_mul(a.val, b));
        auto cells = one_line({Pair(&a, zero), Pair(ValueSchema(cv), zero)}, nullptr, {one}, nullptr);
        AssignedValue c1 = cells[1];
        auto r = one_line_with_last({Pair(&a, zero), Pair(ValueSchema(b), zero)}, Pair(&c1, one), &neg_one, {one}, nullptr);
        return {AssignedCondition(r.second), r.first[1]};
    }
    AssignedCondition is_zero(const AssignedValue& a) { return invert(a).first; }
    // base_chip.rs:327-342
    AssignedValue div_unsafe(const AssignedValue& a, const AssignedValue& b) {
        N bi;
        ORC_ASSERT(n_inv(b.val, bi));
        N cv = n_mul(bi, a.val);
        N one = n_from(1), zero = n_from(0);
        return one_line_with_last({Pair(&b, zero), Pair(ValueSchema(cv), zero)}, Pair(&a, n_neg(one)), nullptr, {one}, nullptr)
            .first[1];
    }
    // base_chip.rs:344-355
    AssignedValue assign_constant(const N& v) { return one_line_add({Pair(ValueSchema(v), n_neg(n_from(1)))}, &v)[0]; }
    AssignedValue assign(const N& v) { return one_line_add({Pair(ValueSchema(v), n_from(0))}, nullptr)[0]; }
    // base_chip.rs:357-367
    AssignedCondition assign_bit(const N& a) {
        N zero = n_from(0), one = n_from(1);
        return AssignedCondition(one_line({Pair(ValueSchema(a), one), Pair(ValueSchema(a), zero)}, nullptr, {n_neg(one)}, nullptr)[0]);
    }
    // base_chip.rs:369-390
    void assert_equal(const AssignedValue& a, const AssignedValue& b) {
        N one = n_from(1);
        one_line_add({Pair(&a, n_neg(one)), Pair(&b, one)}, nullptr);
    }
    void assert_constant(const AssignedValue& a, const N& b) {
        ORC_ASSERT(a.val == b);
        one_line_add({Pair(&a, n_neg(n_from(1)))}, &b);
    }
    void assert_bit(const AssignedValue& a) {
        N zero = n_from(0), one = n_from(1);
        one_line({Pair(&a, one), Pair(&a, zero)}, nullptr, {n_neg(one)}, nullptr);
    }
    // base_chip.rs:392-467
    AssignedCondition and_(const AssignedCondition& a, const AssignedCondition& b) { return AssignedCondition(mul(a.v, b.v)); }
    AssignedCondition not_(const AssignedCondition& a) {
        N one = n_from(1);
        return AssignedCondition(sum_with_constant({Elem(&a.v, n_neg(one))}, &one));
    }
    AssignedCondition not_and(const AssignedCondition& a, const AssignedCondition& b) {
        N one = n_from(1), zero = n_from(0);
        N cv = n_sub(b.v.val, n_mul(a.v.val, b.v.val));
        return AssignedCondition(
            one_line_with_last({Pair(&a.v, zero), Pair(&b.v, one)}, Pair(ValueSchema(cv), n_neg(one)), nullptr, {n_neg(one)}, nullptr)
                .second);
    }
    AssignedCondition or_(const AssignedCondition& a, const AssignedCondition& b) {
        N one = n_from(1);
        N cv = n_sub(n_add(a.v.val, b.v.val), n_mul(a.v.val, b.v.val));
        return AssignedCondition(
            one_line_with_last({Pair(&a.v, one), Pair(&b.v, one)}, Pair(ValueSchema(cv), n_neg(one)), nullptr, {n_neg(one)}, nullptr)
                .second);
    }
    AssignedCondition xor_(const AssignedCondition& a, const AssignedCondition& b) {
        N one = n_from(1), two = n_from(2);
        N cv = n_sub(n_add(a.v.val, b.v.val), n_mul(n_mul(two, a.v.val), b.v.val));
        return AssignedCondition(
            one_line_with_last({Pair(&a.v, one), Pair(&b.v, one)}, Pair(ValueSchema(cv), n_neg(one)), nullptr, {n_neg(two)}, nullptr)
                .second);
    }
    AssignedCondition xnor(const AssignedCondition& a, const AssignedCondition& b) {
        N one = n_from(1), two = n_from(2);
        N cv = n_add(n_sub(n_sub(one, a.v.val), b.v.val), n_mul(n_mul(two, a.v.val), b.v.val));
        return AssignedCondition(
            one_line_with_last({Pair(&a.v, n_neg(one)), Pair(&b.v, n_neg(one))}, Pair(ValueSchema(cv), n_neg(one)), &one, {two}, nullptr)
                .second);
    }
    // base_chip.rs:574-604 (VAR_COLUMNS >= 5 branch)
    AssignedValue bisec(const AssignedCondition& cond, const AssignedValue& a, const AssignedValue& b) {
        N zero = n_from(0), one = n_from(1);
        AssignedValue cond_v = cond.v;
        N cv = n_add(n_mul(cond.v.val, a.val), n_mul(n_sub(one, cond.v.val), b.val));
        return one_line_with_last({Pair(&cond_v, zero), Pair(&a, zero), Pair(&cond_v, zero), Pair(&b, one)},
                                  Pair(ValueSchema(cv), n_neg(one)), nullptr, {one, n_neg(one)}, nullptr)
            .second;
    }
    AssignedCondition bisec_cond(const AssignedCondition& cond, const AssignedCondition& a, const AssignedCondition& b) {
        return AssignedCondition(bisec(cond, a.v, b.v));
    }
    // base_chip.rs:487-500
    void assert_true(const AssignedCondition& a) {
        ORC_ASSERT(a.v.val == n_from(1));
        assert_constant(a.v, n_from(1));
    }
    void assert_false(const AssignedCondition& a) {
        ORC_ASSERT(a.v.val == n_from(0));
        assert_constant(a.v, n_from(0));
    }
    bool try_assert_false(const AssignedCondition& a) {
        // NOTE: the reference's assert_constant asserts a.val == 0 before writing the row
        // (base_chip.rs:375-379), so try_assert_false panics rather than returning false when the
        // value is non-zero. The oracle keeps the row and reports the failure through the return.
        N zero = n_from(0);
        bool ok = a.v.val == zero;
        one_line_add({Pair(&a.v, n_neg(n_from(1)))}, &zero);
        return ok;
    }
};

// ------------------------------- IntegerContext ------------------------------------------------
// src/context.rs:161-188: shares one Context between chips; info = RangeInfo<W,N>.
struct IntegerContext {
    std::shared_ptr<Context> ctx;
    std::shared_ptr<RangeInfo> info;
    BN w_mod;

    IntegerContext(std::shared_ptr<Context> c, const BN& w) : ctx(c), info(std::make_shared<RangeInfo>(w)), w_mod(w) {}
    IntegerContext(std::shared_ptr<Context> c, std::shared_ptr<RangeInfo> i) : ctx(c), info(i), w_mod(i->w_modulus) {}

    BaseOps base() { return BaseOps(ctx.get()); }

    // ---- RangeChipOps (range_chip.rs:270-347) ----
    static void decompose_bn(const BN& bn, uint64_t decompose, const BN& mask, N& v, std::vector<N>& out) {
        v = bn_to_n(bn);
        out.clear();
        for (uint64_t i = 0; i < decompose; i++) out.push_back(bn_to_n((bn >> (i * COMMON_RANGE_BITS)) & mask));
    }
    AssignedValue assign_common(const BN& bn) {
        N v = bn_to_n(bn);
        size_t offset = ctx->range_offset;
        AssignedValue res = ctx->records.assign_one_line_range_value(offset, {v}, v, COMMON_RANGE_BITS);
        ctx->range_offset += 1;
        return res;
    }
    AssignedValue assign_range(const BN& bn, uint64_t decompose, uint64_t bits) {
        N v;
        std::vector<N> dv;
        decompose_bn(bn, decompose, info->common_range_mask, v, dv);
        size_t offset = ctx->range_offset;
        auto r = ctx->records.assign_range_value(offset, dv, v, bits);
        ctx->range_offset += r.second;
        return r.first;
    }
    AssignedValue assign_nonleading_limb(const BN& bn) { return assign_range(bn, MAX_CHUNKS * RANGE_CHIP_RANGE_COLUMNS, info->limb_bits); }
    AssignedValue assign_w_ceil_leading_limb(const BN& bn) {
        return assign_range(bn, info->w_ceil_leading_decompose, info->w_ceil_bits % info->limb_bits);
    }
    AssignedValue assign_d_leading_limb(const BN& bn) {
        return assign_range(bn, info->d_leading_decompose, info->d_bits % info->limb_bits);
    }

    // ---- SelectChipOps (select_chip.rs:118-161) ----
    static N encode_offset(size_t g, size_t offset, size_t limb_offset) {
        return bn_to_n((BN(offset) << 128) + (BN(g) << 64) + BN(limb_offset));
    }
    void assign_cache_value(const AssignedValue& v, size_t offset, size_t group_index, size_t selector) {
        size_t so = ctx->select_offset;
        ctx->records.assign_cache_value(so, v, encode_offset(group_index, selector, offset));
        ctx->select_offset += 1;
    }
    AssignedValue assign_selected_value(const AssignedValue& v, size_t offset, size_t group_index, const AssignedValue& selector) {
        size_t so = ctx->select_offset;
        AssignedValue r = ctx->records.assign_select_value(so, v, encode_offset(group_index, 0, offset), selector);
        ctx->select_offset += 1;
        return r;
    }

    // ---- IntegerChipOps ----
    // integer_chip.rs:217-224
    BN get_w_bn(const AssignedInteger& a) const {
        BN res;
        for (int i = (int)info->limbs - 1; i >= 0; i--) {
            res = res << info->limb_bits;
            res = res + a.limbs_le[i].val;
        }
        return res;
    }

    AssignedValue native_sum(const std::vector<AssignedValue>& limbs) {
        std::vector<BaseOps::Elem> schemas;
        for (size_t i = 0; i < limbs.size(); i++) schemas.push_back(BaseOps::Elem(&limbs[i], info->limb_coeffs[i]));
        return base().sum_with_constant(schemas, nullptr);
    }

    // integer_chip.rs:236-258
    AssignedInteger assign_w(const BN& w) {
        std::vector<AssignedValue> limbs;
        for (uint64_t i = 0; i + 1 < info->limbs; i++) limbs.push_back(assign_nonleading_limb((w >> (i * info->limb_bits)) & info->limb_mask));
        limbs.push_back(assign_w_ceil_leading_limb((w >> ((info->limbs - 1) * info->limb_bits)) & info->limb_mask));
        AssignedValue native = native_sum(limbs);
        return AssignedInteger(limbs, native, 1);
    }
    // integer_chip.rs:260-281
    std::pair<std::vector<AssignedValue>, AssignedValue> assign_d(const BN& d) {
        std::vector<AssignedValue> limbs;
        for (uint64_t i = 0; i + 1 < info->limbs; i++) limbs.push_back(assign_nonleading_limb((d >> (i * info->limb_bits)) & info->limb_mask));
        limbs.push_back(assign_d_leading_limb((d >> ((info->limbs - 1) * info->limb_bits)) & info->limb_mask));
        AssignedValue native = native_sum(limbs);
        return {limbs, native};
    }

    // integer_chip.rs:73-193
    void add_constraints_for_mul_equation_on_limbs(const AssignedInteger& a, const AssignedInteger& b, const std::vector<AssignedValue>& d,
                                                   const AssignedInteger& rem) {
        ORC_ASSERT(a.times < info->overflow_limit);
        ORC_ASSERT(b.times < info->overflow_limit);
        ORC_ASSERT(rem.times == 1);
        N one = n_from(1), neg_one = n_neg(one);
        size_t L = info->limbs;
        std::vector<AssignedValue> limbs;
        for (size_t pos = 0; pos < info->mul_check_limbs; pos++) {
            size_t r_bound = std::min(pos + 1, L);
            size_t l_bound = pos >= L - 1 ? pos - (L - 1) : 0;
            std::vector<BaseOps::MulAddTerm> terms;
            for (size_t i = l_bound; i < r_bound; i++)
                terms.push_back({&a.limbs_le[i], &b.limbs_le[pos - i], &d[i], n_neg(info->w_modulus_limbs_le[pos - i])});
            limbs.push_back(base().mul_add_with_next_line(terms));
        }
        N borrow = n_add(n_mul(n_from(L), info->limb_modulus_n), n_from(2));
        N c0 = n_mul(info->limb_modulus_n, borrow);
        AssignedValue u = base().sum_with_constant({BaseOps::Elem(&limbs[0], one), BaseOps::Elem(&rem.limbs_le[0], neg_one)}, &c0);
        BN v, r;
        BN::div_rem(u.val, info->limb_modulus, v, r);
        ORC_ASSERT(r.is_zero());
        BN v_h_bn, v_l_bn;
        BN::div_rem(v, info->limb_modulus, v_h_bn, v_l_bn);
        AssignedValue v_h = assign_common(v_h_bn);
        AssignedValue v_l = assign_nonleading_limb(v_l_bn);
        base().one_line_with_last({Pair(&v_h, info->limb_coeffs[2]), Pair(&v_l, info->limb_coeffs[1])}, Pair(&u, neg_one), nullptr, {},
                                  nullptr);
        N c1 = n_sub(n_mul(info->limb_modulus_n, borrow), borrow);
        for (size_t i = 1; i < info->mul_check_limbs; i++) {
            std::vector<BaseOps::Elem> elems;
            elems.push_back(BaseOps::Elem(&limbs[i], one));
            if (i < L) elems.push_back(BaseOps::Elem(&rem.limbs_le[i], neg_one));  // integer_chip.rs:136-145 vs 167-175
            elems.push_back(BaseOps::Elem(&v_h, info->limb_coeffs[1]));
            elems.push_back(BaseOps::Elem(&v_l, info->limb_coeffs[0]));
            AssignedValue ui = base().sum_with_constant(elems, &c1);
            BN::div_rem(ui.val, info->limb_modulus, v, r);
            ORC_ASSERT(r.is_zero());
            BN::div_rem(v, info->limb_modulus, v_h_bn, v_l_bn);
            v_h = assign_common(v_h_bn);
            v_l = assign_nonleading_limb(v_l_bn);
            base().one_line_with_last({Pair(&v_h, info->limb_coeffs[2]), Pair(&v_l, info->limb_coeffs[1])}, Pair(&ui, neg_one), nullptr, {},
                                      nullptr);
        }
        ORC_ASSERT(info->limbs <= info->mul_check_limbs);
    }

    // integer_chip.rs:195-215
    void add_constraints_for_mul_equation_on_native(const AssignedInteger& a, const AssignedInteger& b, const AssignedValue& d_native,
                                                    const AssignedInteger& rem) {
        N zero = n_from(0), one = n_from(1);
        base().one_line({Pair(&a.native, zero), Pair(&b.native, zero), Pair(&d_native, info->w_native), Pair(&rem.native, one)}, nullptr,
                        {n_neg(one)}, nullptr);
    }

    // integer_chip.rs:283-373
    AssignedInteger reduce(const AssignedInteger& a) {
        if (a.times == 1) return a;
        N zero = n_from(0), one = n_from(1), neg_one = n_neg(one);
        uint64_t overflow_limit = info->overflow_limit;
        ORC_ASSERT(a.times < overflow_limit);
        BN a_bn = get_w_bn(a);
        BN d, rem;
        BN::div_rem(a_bn, info->w_modulus, d, rem);
        AssignedInteger assigned_rem = assign_w(rem);
        AssignedValue assigned_d = assign_common(d);
        base().one_line_with_last({Pair(&assigned_d, info->w_native), Pair(&assigned_rem.native, one)}, Pair(&a.native, neg_one), nullptr,
                                  {}, nullptr);
        bool have_last = false;
        AssignedValue last_v;
        std::vector<BN> rem_limbs = info->bn_to_limb_le(rem);
        for (size_t i = 0; i < info->reduce_check_limbs; i++) {
            uint64_t last_borrow = i != 0 ? overflow_limit : 0;
            BN carry = have_last ? last_v.val : BN(0);
            BN u = d * info->w_modulus_limbs_le_bn[i] + rem_limbs[i] + info->limb_modulus * BN(overflow_limit) - a.limbs_le[i].val + carry -
                   BN(last_borrow);
            BN v, v_rem;
            BN::div_rem(u, info->limb_modulus, v, v_rem);
            ORC_ASSERT(v_rem.is_zero());
            AssignedValue vv = assign_nonleading_limb(v);
            N kconst = bn_to_n(info->limb_modulus * BN(overflow_limit) - BN(i == 0 ? 0 : overflow_limit));
            base().one_line_with_last({Pair(&assigned_d, info->w_modulus_limbs_le[i]), Pair(&assigned_rem.limbs_le[i], one),
                                       Pair(&a.limbs_le[i], neg_one), have_last ? Pair(&last_v, one) : Pair(ValueSchema(zero), zero)},
                                      Pair(&vv, n_neg(bn_to_n(info->limb_modulus))), &kconst, {}, nullptr);
            last_v = vv;
            have_last = true;
        }
        return assigned_rem;
    }

    // integer_chip.rs:375-382
    AssignedInteger conditionally_reduce(const AssignedInteger& a) {
        uint64_t threshold = 1ull << (info->overflow_bits - 2);
        return a.times > threshold ? reduce(a) : a;
    }

    // integer_chip.rs:384-406
    AssignedInteger int_add(const AssignedInteger& a, const AssignedInteger& b) {
        std::vector<AssignedValue> limbs;
        for (size_t i = 0; i < info->limbs; i++) limbs.push_back(base().add(a.limbs_le[i], b.limbs_le[i]));
        AssignedValue native = native_sum(limbs);
        return conditionally_reduce(AssignedInteger(limbs, native, a.times + b.times));
    }
    // integer_chip.rs:408-437
    AssignedInteger int_sub(const AssignedInteger& a, const AssignedInteger& b) {
        ORC_ASSERT(b.times >= 1 && b.times < info->overflow_limit);
        const std::vector<N>& upper = info->w_modulus_of_ceil_times[b.times];
        N one = n_from(1), neg_one = n_neg(one);
        std::vector<AssignedValue> limbs;
        for (size_t i = 0; i < info->limbs; i++)
            limbs.push_back(base().sum_with_constant({BaseOps::Elem(&a.limbs_le[i], one), BaseOps::Elem(&b.limbs_le[i], neg_one)}, &upper[i]));
        AssignedValue native = native_sum(limbs);
        return conditionally_reduce(AssignedInteger(limbs, native, a.times + b.times + 1));
    }
    // integer_chip.rs:439-464
    AssignedInteger int_neg(const AssignedInteger& a) {
        ORC_ASSERT(a.times >= 1 && a.times < info->overflow_limit);
        const std::vector<N>& upper = info->w_modulus_of_ceil_times[a.times];
        N neg_one = n_neg(n_from(1));
        std::vector<AssignedValue> limbs;
        for (size_t i = 0; i < info->limbs; i++) limbs.push_back(base().sum_with_constant({BaseOps::Elem(&a.limbs_le[i], neg_one)}, &upper[i]));
        AssignedValue native = native_sum(limbs);
        return conditionally_reduce(AssignedInteger(limbs, native, a.times + 1));
    }
    // integer_chip.rs:466-483
    AssignedInteger int_mul(const AssignedInteger& a, const AssignedInteger& b) {
        BN a_bn = get_w_bn(a), b_bn = get_w_bn(b);
        BN d, rem;
        BN::div_rem(a_bn * b_bn, info->w_modulus, d, rem);
        AssignedInteger rem_a = assign_w(rem);
        auto d_a = assign_d(d);
        add_constraints_for_mul_equation_on_limbs(a, b, d_a.first, rem_a);
        add_constraints_for_mul_equation_on_native(a, b, d_a.second, rem_a);
        return rem_a;
    }
    // integer_chip.rs:485-491
    AssignedInteger int_unsafe_invert(const AssignedInteger& x) {
        AssignedInteger one = assign_int_constant(BN(1));
        auto r = int_div(one, x);
        base().assert_false(r.first);
        return r.second;
    }
    // integer_chip.rs:493-538
    std::pair<AssignedCondition, AssignedInteger> int_div(const AssignedInteger& a_in, const AssignedInteger& b_in) {
        AssignedInteger b = reduce(b_in);
        AssignedCondition is_b_zero = is_int_zero(b);
        AssignedCondition a_coeff = base().not_(is_b_zero);
        AssignedInteger a;
        {
            AssignedInteger ar = reduce(a_in);
            std::vector<AssignedValue> limbs_le;
            for (size_t i = 0; i < info->limbs; i++) limbs_le.push_back(base().mul(ar.limbs_le[i], a_coeff.v));
            AssignedValue native = base().mul(ar.native, a_coeff.v);
            a = AssignedInteger(limbs_le, native, ar.times);
        }
        BN a_bn = get_w_bn(a), b_bn = get_w_bn(b);
        BN c_bn;
        {
            BN binv;
            if (bn_modinv(b_bn % w_mod, w_mod, binv))
                c_bn = ((a_bn % w_mod) * binv) % w_mod;
            else
                c_bn = BN(0);
        }
        BN d_bn = (b_bn * c_bn - a_bn) / info->w_modulus;
        AssignedInteger c = assign_w(c_bn);
        auto d = assign_d(d_bn);
        add_constraints_for_mul_equation_on_limbs(b, c, d.first, a);
        add_constraints_for_mul_equation_on_native(b, c, d.second, a);
        return {is_b_zero, c};
    }
    // integer_chip.rs:540-548
    AssignedCondition is_pure_zero(const AssignedInteger& a) {
        std::vector<BaseOps::Elem> e;
        for (auto& v : a.limbs_le) e.push_back(BaseOps::Elem(&v, n_from(1)));
        AssignedValue sum = base().sum_with_constant(e, nullptr);
        return base().is_zero(sum);
    }
    // integer_chip.rs:550-570
    AssignedCondition is_pure_w_modulus(const AssignedInteger& a) {
        ORC_ASSERT(a.times == 1);
        AssignedValue native_diff = base().add_constant(a.native, n_neg(info->w_native));
        AssignedCondition is_eq = base().is_zero(native_diff);
        for (size_t i = 0; i < info->pure_w_check_limbs; i++) {
            AssignedValue limb_diff = base().add_constant(a.limbs_le[i], n_neg(info->w_modulus_limbs_le[i]));
            AssignedCondition is_limb_eq = base().is_zero(limb_diff);
            is_eq = base().and_(is_eq, is_limb_eq);
        }
        return is_eq;
    }
    // integer_chip.rs:572-578
    AssignedCondition is_int_zero(const AssignedInteger& a_in) {
        AssignedInteger a = reduce(a_in);
        AssignedCondition is_zero = is_pure_zero(a);
        AssignedCondition is_w_modulus = is_pure_w_modulus(a);
        return base().or_(is_zero, is_w_modulus);
    }
    // integer_chip.rs:47-54
    AssignedCondition is_int_equal(const AssignedInteger& a, const AssignedInteger& b) {
        AssignedInteger diff = int_sub(a, b);
        return is_int_zero(diff);
    }
    // integer_chip.rs:580-598 (w is the canonical value of the W element)
    AssignedInteger assign_int_constant(const BN& w) {
        std::vector<N> limbs_value = info->bn_to_limb_le_n(w);
        std::vector<AssignedValue> limbs;
        for (auto& l : limbs_value) limbs.push_back(base().assign_constant(l));
        AssignedValue native = base().assign_constant(bn_to_n(w % info->n_modulus));
        return AssignedInteger(limbs, native, 1);
    }
    // integer_chip.rs:600-612
    void assert_int_equal(const AssignedInteger& a, const AssignedInteger& b) {
        AssignedInteger diff = int_sub(a, b);
        diff = reduce(diff);
        std::vector<BaseOps::Elem> e;
        for (auto& v : diff.limbs_le) e.push_back(BaseOps::Elem(&v, n_from(1)));
        AssignedValue sum = base().sum_with_constant(e, nullptr);
        base().assert_constant(sum, n_from(0));
    }
    // integer_chip.rs:614-616
    AssignedInteger int_square(const AssignedInteger& a) { return int_mul(a, a); }
    // integer_chip.rs:618-658
    AssignedInteger int_mul_small_constant(const AssignedInteger& a_in, uint64_t b) {
        uint64_t threshold = 1ull << (info->overflow_bits - 2);
        ORC_ASSERT(b < threshold);
        AssignedInteger a = (a_in.times * b >= info->overflow_limit) ? reduce(a_in) : a_in;
        std::vector<AssignedValue> limbs;
        for (size_t i = 0; i < info->limbs; i++) limbs.push_back(base().sum_with_constant({BaseOps::Elem(&a.limbs_le[i], n_from(b))}, nullptr));
        AssignedValue native = native_sum(limbs);
        return conditionally_reduce(AssignedInteger(limbs, native, a.times * b));
    }
    // integer_chip.rs:660-681
    AssignedInteger bisec_int(const AssignedCondition& cond, const AssignedInteger& a, const AssignedInteger& b) {
        std::vector<AssignedValue> limbs;
        for (size_t i = 0; i < info->limbs; i++) limbs.push_back(base().bisec(cond, a.limbs_le[i], b.limbs_le[i]));
        AssignedValue native = base().bisec(cond, a.native, b.native);
        return AssignedInteger(limbs, native, std::max(a.times, b.times));
    }
    // integer_chip.rs:683-685
    BN get_w(const AssignedInteger& a) const { return get_w_bn(a) % w_mod; }
};

}  // namespace orc
