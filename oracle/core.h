// ORACLE (test infrastructure, NOT product code). Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may call it; the product library never does.
// Pinning: the reference crate cannot be built in this image (no cargo / nightly toolchain / git
// dependencies) and its tests hold no known-answer vectors (all inputs are OsRng), so no output of the
// Rust implementation itself is available: parity against the reference BINARY is unpinned. What pins
// this restatement instead: (1) the reference's embedded constant tables (tests/golden/), (2) the
// structural row / permutation counts of SURVEY Appendix B, (3) a restatement of the reference's gate
// constraints that every emitted record must satisfy (gate_check.h, the MockProver equivalent), and
// (4) independent plain-math MSM / pairing values. See DESIGN.md section 4.
// CPU restatement of the reference's data model: field<->bigint helpers (src/utils.rs:4-17),
// RangeInfo (src/range_info.rs:14-359), Assigned* handles (src/assign.rs:5-229) and the record
// store write side (src/context.rs:36-46,135-158,241-301,590-997).
#pragma once
#include <memory>
#include <vector>

#include "bn.h"

namespace orc {

// ---------------------------------------------------------------------------------------------
// Field moduli. The arithmetic of these fields lives in third-party crates that are not under
// /root/reference (pairing_bn256 0.1.1 @5ab0806, bls12_381 0.7.0 @31fcd53, re-exported by
// halo2_proofs @9a81f60; Cargo.lock:94-103,434-447,693-706). Prime-field results are unique, so
// they are restated as plain modular arithmetic on canonical values.
// ---------------------------------------------------------------------------------------------
inline const BN& BN256_FR() {
    static BN v = BN::from_hex("30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001");
    return v;
}
inline const BN& BN256_FQ() {
    static BN v = BN::from_hex("30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47");
    return v;
}
inline const BN& BLS12_381_FQ() {
    static BN v = BN::from_hex(
        "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab");
    return v;
}
inline const BN& BLS12_381_FR() {
    static BN v = BN::from_hex("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001");
    return v;
}

// N is always bn256 Fr on this path (all reference tests run "over bn256 Fr").
typedef BN N;
inline const BN& NMOD() { return BN256_FR(); }
inline N n_from(uint64_t v) { return BN(v); }
inline N n_add(const N& a, const N& b) {
    BN s = a + b;
    return s >= NMOD() ? s - NMOD() : s;
}
inline N n_sub(const N& a, const N& b) { return a >= b ? a - b : a + NMOD() - b; }
inline N n_neg(const N& a) { return a.is_zero() ? a : NMOD() - a; }
inline N n_mul(const N& a, const N& b) { return (a * b) % NMOD(); }
inline bool n_inv(const N& a, N& out) { return bn_modinv(a, NMOD(), out); }

// src/utils.rs:10-17 (bn_to_field reduces mod the field modulus)
inline N bn_to_n(const BN& bn) { return bn % NMOD(); }

// ---------------------------------------------------------------------------------------------
// chip constants (src/circuit/base_chip.rs:14-16, range_chip.rs:22-33, select_chip.rs:18,
// context.rs:36-38, ecc_chip.rs:20-21)
// ---------------------------------------------------------------------------------------------
static const int VAR_COLUMNS = 5;
static const int MUL_COLUMNS = 2;
static const int FIXED_COLUMNS = VAR_COLUMNS + MUL_COLUMNS + 2;
static const uint64_t MAX_CHUNKS = 3;
static const uint64_t COMMON_RANGE_BITS = 18;
static const int RANGE_CHIP_RANGE_COLUMNS = 2;
static const int RANGE_CHIP_ADV_COLUMNS = 3;
static const int RANGE_CHIP_FIX_COLUMNS = 2;
static const uint64_t RANGE_VALUE_DECOMPOSE = MAX_CHUNKS * RANGE_CHIP_RANGE_COLUMNS;
static const uint64_t OVERFLOW_BITS = 6;
static const size_t MSM_PREFIX_OFFSET = 1u << 20;
static const size_t MSM_LIMIT = (1u << 8) * MSM_PREFIX_OFFSET;
static const int SELECTOR_ENCODE_OFFSET = 128;

enum RangeAdvCol { ValueAccCol = 0, TaggedRangeCol = 1, CommonRangeCol = 2 };
enum RangeFixCol { AccLinesCol = 0, TagCol = 1 };
enum SelectAdvCol { SelValueCol = 0, SelSelectCol = 1 };
enum SelectFixCol { EncodeCol = 0, IsLookupCol = 1 };

// ---------------------------------------------------------------------------------------------
// RangeInfo (src/range_info.rs)
// ---------------------------------------------------------------------------------------------
struct RangeInfo {
    uint64_t limbs, limb_bits;
    uint64_t w_ceil_leading_decompose, n_floor_leading_decompose, d_leading_decompose;
    uint64_t w_ceil_bits, d_bits, n_floor_bits;
    uint64_t d_leading_bits, w_ceil_leading_bits, n_floor_leading_bits;
    BN w_ceil, n_modulus, w_modulus, common_range_mask, limb_mask, limb_modulus, max_d;
    std::vector<BN> w_modulus_limbs_le_bn;
    std::vector<N> w_modulus_limbs_le;
    std::vector<N> limb_coeffs;
    N limb_modulus_n;
    uint64_t overflow_bits, overflow_limit;
    N w_native;
    uint64_t pure_w_check_limbs, reduce_check_limbs, mul_check_limbs;
    std::vector<std::vector<N>> w_modulus_of_ceil_times;  // index 0 unused (None)

    // range_info.rs:57-75
    static void bits_to_leading_bits_and_decompose(uint64_t bits, uint64_t common_bits, uint64_t& lead, uint64_t& dec) {
        uint64_t common_limb_bits = RANGE_VALUE_DECOMPOSE * common_bits;
        uint64_t leading_bits = (bits % common_limb_bits == 0) ? common_limb_bits : bits % common_limb_bits;
        ORC_ASSERT(leading_bits >= 2 * common_bits);
        ORC_ASSERT(leading_bits <= RANGE_VALUE_DECOMPOSE * common_bits);
        uint64_t leading_chunk_bits = leading_bits % common_bits;
        if (leading_chunk_bits == 0) {
            lead = common_bits;
            dec = leading_bits / common_bits;
        } else {
            lead = leading_chunk_bits;
            dec = leading_bits / common_bits + 1;
        }
    }

    // range_info.rs:299-314
    static uint64_t calc_d_bits(const BN& w, uint64_t overflow_bits) {
        BN w_max = w - BN(1);
        uint64_t w_ceil_bits = w_max.bits();
        uint64_t d_bits = w_ceil_bits + overflow_bits * 2 + 1;
        BN max_a = bn_pow2(w_ceil_bits + overflow_bits);
        ORC_ASSERT(bn_pow2(d_bits) * w >= max_a * max_a);
        return d_bits;
    }

    // range_info.rs:77-184
    RangeInfo(const BN& w_mod, uint64_t common_bits = COMMON_RANGE_BITS, uint64_t overflow_bits_ = OVERFLOW_BITS) {
        ORC_ASSERT(common_bits == COMMON_RANGE_BITS);
        ORC_ASSERT(overflow_bits_ == OVERFLOW_BITS);
        BN w_max = w_mod - BN(1);
        w_ceil_bits = w_max.bits();
        bits_to_leading_bits_and_decompose(w_ceil_bits, common_bits, w_ceil_leading_bits, w_ceil_leading_decompose);
        BN n_max = NMOD() - BN(1);
        n_floor_bits = n_max.bits() - 1;
        bits_to_leading_bits_and_decompose(n_floor_bits, common_bits, n_floor_leading_bits, n_floor_leading_decompose);
        d_bits = calc_d_bits(w_mod, overflow_bits_);
        bits_to_leading_bits_and_decompose(d_bits, common_bits, d_leading_bits, d_leading_decompose);

        limb_bits = common_bits * RANGE_VALUE_DECOMPOSE;
        limbs = (w_ceil_bits + limb_bits - 1) / limb_bits;
        max_d = bn_pow2(d_bits);
        limb_mask = bn_pow2(limb_bits) - BN(1);
        n_modulus = n_max + BN(1);
        w_modulus = w_max + BN(1);
        BN w_native_bn = w_modulus % n_modulus;
        for (uint64_t i = 0; i < limbs; i++) {
            w_modulus_limbs_le_bn.push_back((w_modulus >> (i * limb_bits)) & limb_mask);
            w_modulus_limbs_le.push_back(bn_to_n(w_modulus_limbs_le_bn.back()));
        }
        limb_modulus = bn_pow2(limb_bits);
        limb_modulus_n = bn_to_n(limb_modulus);
        overflow_bits = overflow_bits_;
        overflow_limit = 1ull << overflow_bits;
        w_ceil = bn_pow2(w_ceil_bits);
        common_range_mask = BN((1ull << common_bits) - 1);
        for (uint64_t i = 0; i < limbs; i++) limb_coeffs.push_back(bn_to_n(bn_pow2(i * limb_bits)));
        w_native = bn_to_n(w_native_bn);
        pure_w_check_limbs = (w_ceil_bits - n_floor_bits + limb_bits - 1) / limb_bits;
        mul_check_limbs =
            (std::max(w_ceil_bits * 2 + overflow_bits * 2, d_bits + w_ceil_bits) - n_floor_bits + limb_bits - 1) / limb_bits;
        reduce_check_limbs =
            (std::max(w_ceil_bits + overflow_bits, common_bits + w_ceil_bits) - n_floor_bits + limb_bits - 1) / limb_bits;
        w_modulus_of_ceil_times.resize(overflow_limit);
        for (uint64_t i = 1; i < overflow_limit; i++) w_modulus_of_ceil_times[i] = find_w_modulus_of_ceil_times(i);
        pre_check();
    }

    // range_info.rs:186-297
    void pre_check() const {
        uint64_t common_modulus = 1ull << COMMON_RANGE_BITS;
        {
            BN limb_check_modulus = bn_pow2(limb_bits * pure_w_check_limbs);
            ORC_ASSERT(bn_lcm(n_modulus, limb_check_modulus) >= w_ceil);
        }
        BN max_wi = w_modulus_limbs_le_bn[0];
        for (auto& x : w_modulus_limbs_le_bn) max_wi = bn_max(max_wi, x);
        {
            BN max_a = w_ceil * BN(overflow_limit - 1) - BN(1);
            BN max_d_ = bn_pow2(COMMON_RANGE_BITS) - BN(1);
            ORC_ASSERT(max_a <= max_d_ * w_modulus);
            BN lm = bn_pow2(limb_bits * reduce_check_limbs);
            ORC_ASSERT(bn_lcm(n_modulus, lm) >= max_d_ * w_modulus + w_ceil);
            BN max_v = limb_modulus - BN(1);
            BN max_rem = limb_modulus - BN(1);
            ORC_ASSERT(max_v * limb_modulus >= max_d_ * max_wi + max_rem + max_v + BN(overflow_limit) * limb_modulus);
            ORC_ASSERT(max_v * limb_modulus < n_modulus);
            ORC_ASSERT(max_d_ * max_wi + max_rem + max_v + BN(overflow_limit) * limb_modulus < n_modulus);
            BN max_ai = limb_modulus * BN(overflow_limit - 1) - BN(1);
            ORC_ASSERT(BN(overflow_limit) * limb_modulus - BN(overflow_limit) >= max_ai);
        }
        {
            BN max_a = w_ceil * BN(overflow_limit - 1) - BN(1);
            BN max_d_ = bn_pow2(d_bits) - BN(1);
            ORC_ASSERT(max_a * max_a <= max_d_ * w_modulus);
            BN lcm = bn_lcm(n_modulus, bn_pow2(limb_bits * mul_check_limbs));
            BN max_rem = w_ceil - BN(1);
            ORC_ASSERT(lcm > max_a * max_a);
            ORC_ASSERT(lcm > max_d_ * w_modulus + max_rem);
            BN borrow = BN(limbs) * limb_modulus + BN(2);
            BN max_d_j = limb_modulus - BN(1);
            BN max_rem_i = limb_modulus - BN(1);
            ORC_ASSERT(borrow * limb_modulus - borrow >= BN(limbs) * max_d_j * max_wi + max_rem_i);
            BN max_v = limb_modulus * BN(common_modulus) - BN(1);
            BN max_a_j = limb_modulus * BN(overflow_limit - 1);
            ORC_ASSERT(max_v * limb_modulus >= max_a_j * max_a_j * BN(limbs) + limb_modulus * borrow);
            ORC_ASSERT(max_v * limb_modulus < n_modulus);
        }
        ORC_ASSERT(limbs >= 3);
    }

    // range_info.rs:316-332
    std::vector<N> bn_to_limb_le_n(const BN& w) const {
        std::vector<N> r;
        for (uint64_t i = 0; i < limbs; i++) r.push_back(bn_to_n((w >> (i * limb_bits)) & limb_mask));
        return r;
    }
    std::vector<BN> bn_to_limb_le(const BN& w) const {
        std::vector<BN> r;
        for (uint64_t i = 0; i < limbs; i++) r.push_back((w >> (i * limb_bits)) & limb_mask);
        return r;
    }

    // range_info.rs:334-359
    std::vector<N> find_w_modulus_of_ceil_times(uint64_t times) const {
        BN max = w_ceil * BN(times);
        BN n, rem;
        BN::div_rem(max, w_modulus, n, rem);
        if (rem > BN(0)) n = n + BN(1);
        BN upper = w_modulus * n;
        std::vector<N> out;
        for (uint64_t i = 0; i + 1 < limbs; i++) {
            BN r = (upper & limb_mask) + limb_modulus * BN(times);
            upper = (upper - r) >> limb_bits;
            out.push_back(bn_to_n(r));
            ORC_ASSERT(r >= limb_modulus * BN(times) - BN(1));
            ORC_ASSERT(r < limb_modulus * BN(times + 1));
        }
        ORC_ASSERT(upper >= bn_pow2(w_ceil_bits % limb_bits) * BN(times));
        ORC_ASSERT(upper < bn_pow2(w_ceil_bits % limb_bits) * BN(times + 1));
        out.push_back(bn_to_n(upper));
        return out;
    }
};

// ---------------------------------------------------------------------------------------------
// Assigned handles (src/assign.rs)
// ---------------------------------------------------------------------------------------------
enum Chip { BaseChip = 0, RangeChip = 1, SelectChip = 2 };

struct Cell {
    Chip region;
    uint32_t col;
    uint32_t row;
    bool operator==(const Cell& o) const { return region == o.region && col == o.col && row == o.row; }
};

struct AssignedValue {
    Cell cell;
    N val;
    AssignedValue() : cell{BaseChip, 0, 0} {}
    AssignedValue(Chip region, uint32_t col, uint32_t row, const N& v) : cell{region, col, row}, val(v) {}
};

struct AssignedCondition {
    AssignedValue v;  // tuple field .0 in the reference
    AssignedCondition() {}
    explicit AssignedCondition(const AssignedValue& a) : v(a) {}
};

struct AssignedInteger {
    std::vector<AssignedValue> limbs_le;
    AssignedValue native;
    uint64_t times;
    AssignedInteger() : times(0) {}
    AssignedInteger(const std::vector<AssignedValue>& l, const AssignedValue& n, uint64_t t) : limbs_le(l), native(n), times(t) {}
};

// ValueSchema (assign.rs:123-146): either a reference to an assigned cell or a raw value.
struct ValueSchema {
    bool assigned;
    Cell cell;
    N val;
    ValueSchema(const AssignedValue& a) : assigned(true), cell(a.cell), val(a.val) {}
    ValueSchema(const AssignedValue* a) : assigned(true), cell(a->cell), val(a->val) {}
    ValueSchema(const N& v) : assigned(false), cell{BaseChip, 0, 0}, val(v) {}
};
typedef std::pair<ValueSchema, N> Pair;  // pair!(x, y)

// ---------------------------------------------------------------------------------------------
// Records (src/context.rs:241-301, 590-997). Rows are stored compactly (32-byte values).
// ---------------------------------------------------------------------------------------------
struct V256 {
    uint64_t w[4];
};
inline V256 to_v256(const BN& b) {
    V256 v;
    for (int i = 4; i < BN::NW; i++) ORC_ASSERT(b.w[i] == 0);
    memcpy(v.w, b.w, 32);
    return v;
}
inline BN from_v256(const V256& v) {
    BN b;
    memcpy(b.w, v.w, 32);
    return b;
}

struct AdvCell {
    V256 v;
    uint8_t some;
    uint8_t permute;
};
struct FixCell {
    V256 v;
    uint8_t some;
};

template <int ADV, int FIX>
struct RegionStore {
    std::vector<AdvCell> adv;  // [row][ADV]
    std::vector<FixCell> fix;  // [row][FIX]
    void ensure(size_t row) {
        if ((row + 1) * ADV > adv.size()) {
            size_t nrows = std::max((row + 1) * 2, (size_t)1024);
            adv.resize(nrows * ADV, AdvCell{{{0, 0, 0, 0}}, 0, 0});
            fix.resize(nrows * FIX, FixCell{{{0, 0, 0, 0}}, 0});
        }
    }
    AdvCell& a(size_t row, int col) {
        ensure(row);
        return adv[row * ADV + col];
    }
    FixCell& f(size_t row, int col) {
        ensure(row);
        return fix[row * FIX + col];
    }
};

struct RecordsInner {
    RegionStore<VAR_COLUMNS, FIXED_COLUMNS> base;
    RegionStore<RANGE_CHIP_ADV_COLUMNS, RANGE_CHIP_FIX_COLUMNS> range;
    RegionStore<2, 2> select;
};

struct Records {
    std::shared_ptr<RecordsInner> inner;
    size_t base_height = 0, range_height = 0, select_height = 0;
    std::vector<std::pair<Cell, Cell>> permutations;

    Records() : inner(std::make_shared<RecordsInner>()) {}

    // context.rs:590-608
    void enable_permute(const Cell& cell) {
        switch (cell.region) {
            case BaseChip: inner->base.a(cell.row, cell.col).permute = 1; break;
            case RangeChip: inner->range.a(cell.row, cell.col).permute = 1; break;
            case SelectChip: inner->select.a(cell.row, cell.col).permute = 1; break;
        }
    }
    void assign_adv_cell_in_base_chip(size_t offset, int col, const N& val) {
        AdvCell& c = inner->base.a(offset, col);
        c.v = to_v256(val);
        c.some = 1;
    }
    void assign_fix_cell_in_base_chip(size_t offset, int col, const N& val) {
        FixCell& c = inner->base.f(offset, col);
        c.v = to_v256(val);
        c.some = 1;
    }

    // context.rs:634-683
    void one_line(size_t offset, const std::vector<Pair>& base_coeff_pairs, const N* constant, const std::vector<N>& mul_coeffs,
                  const N* next) {
        ORC_ASSERT(base_coeff_pairs.size() <= (size_t)VAR_COLUMNS);
        if (offset >= base_height) base_height = offset + 1;
        for (size_t i = 0; i < base_coeff_pairs.size(); i++) {
            const ValueSchema& base = base_coeff_pairs[i].first;
            if (base.assigned) {
                Cell new_cell{BaseChip, (uint32_t)i, (uint32_t)offset};
                enable_permute(new_cell);
                enable_permute(base.cell);
                permutations.push_back({base.cell, new_cell});
            }
            assign_adv_cell_in_base_chip(offset, i, base.val);
            assign_fix_cell_in_base_chip(offset, i, base_coeff_pairs[i].second);
        }
        for (size_t i = 0; i < mul_coeffs.size(); i++) assign_fix_cell_in_base_chip(offset, VAR_COLUMNS + i, mul_coeffs[i]);
        if (next) {
            assign_fix_cell_in_base_chip(offset, VAR_COLUMNS + MUL_COLUMNS, *next);
        } else {
            ORC_ASSERT(!inner->base.f(offset, VAR_COLUMNS + MUL_COLUMNS).some);
        }
        if (constant) {
            assign_fix_cell_in_base_chip(offset, VAR_COLUMNS + MUL_COLUMNS + 1, *constant);
        } else {
            ORC_ASSERT(!inner->base.f(offset, VAR_COLUMNS + MUL_COLUMNS + 1).some);
        }
    }

    // context.rs:685-714
    void one_line_with_last(size_t offset, const std::vector<Pair>& base_coeff_pairs, const Pair& tail, const N* constant,
                            const std::vector<N>& mul_coeffs, const N* next) {
        ORC_ASSERT(base_coeff_pairs.size() <= (size_t)VAR_COLUMNS - 1);
        one_line(offset, base_coeff_pairs, constant, mul_coeffs, next);
        const ValueSchema& base = tail.first;
        int i = VAR_COLUMNS - 1;
        if (base.assigned) {
            Cell new_cell{BaseChip, (uint32_t)i, (uint32_t)offset};
            enable_permute(new_cell);
            enable_permute(base.cell);
            permutations.push_back({base.cell, new_cell});
        }
        assign_adv_cell_in_base_chip(offset, i, base.val);
        assign_fix_cell_in_base_chip(offset, i, tail.second);
    }

    // context.rs:716-720
    void ensure_range_record_size(size_t offset) {
        if (offset >= range_height) range_height = offset + 1;
    }

    void assign_adv_cell_in_select_chip(size_t offset, int col, const N& val) {
        AdvCell& c = inner->select.a(offset, col);
        c.v = to_v256(val);
        c.some = 1;
    }
    void assign_fix_cell_in_select_chip(size_t offset, int col, const N& val) {
        FixCell& c = inner->select.f(offset, col);
        c.v = to_v256(val);
        c.some = 1;
    }

    // context.rs:749-767
    void assign_cache_value(size_t offset, const AssignedValue& v, const N& encode) {
        if (offset >= select_height) select_height = offset + 1;
        assign_adv_cell_in_select_chip(offset, SelValueCol, v.val);
        Cell idx{SelectChip, SelValueCol, (uint32_t)offset};
        permutations.push_back({idx, v.cell});
        enable_permute(idx);
        enable_permute(v.cell);
        assign_fix_cell_in_select_chip(offset, EncodeCol, encode);
        assign_fix_cell_in_select_chip(offset, IsLookupCol, n_from(0));
    }

    // context.rs:769-801
    AssignedValue assign_select_value(size_t offset, const AssignedValue& v, const N& encode, const AssignedValue& selector) {
        if (offset >= select_height) select_height = offset + 1;
        assign_adv_cell_in_select_chip(offset, SelValueCol, v.val);
        assign_adv_cell_in_select_chip(offset, SelSelectCol, selector.val);
        Cell selector_cell{SelectChip, SelSelectCol, (uint32_t)offset};
        permutations.push_back({selector_cell, selector.cell});
        enable_permute(selector_cell);
        enable_permute(selector.cell);
        assign_fix_cell_in_select_chip(offset, EncodeCol, encode);
        assign_fix_cell_in_select_chip(offset, IsLookupCol, n_from(1));
        return AssignedValue(SelectChip, SelValueCol, offset, v.val);
    }

    void assign_adv_cell_in_range_chip(size_t offset, int col, const N& val) {
        AdvCell& c = inner->range.a(offset, col);
        c.v = to_v256(val);
        c.some = 1;
    }
    void assign_fix_cell_in_range_chip(size_t offset, int col, const N& val) {
        FixCell& c = inner->range.f(offset, col);
        c.v = to_v256(val);
        c.some = 1;
    }

    // context.rs:835-857
    AssignedValue assign_one_line_range_value(size_t offset, const std::vector<N>& v, const N& v_acc, uint64_t bits) {
        ORC_ASSERT(bits <= COMMON_RANGE_BITS);
        ensure_range_record_size(offset + 1);
        assign_fix_cell_in_range_chip(offset, AccLinesCol, n_from(1));
        assign_fix_cell_in_range_chip(offset, TagCol, n_from(bits));
        assign_adv_cell_in_range_chip(offset, TaggedRangeCol, v[0]);
        assign_adv_cell_in_range_chip(offset, ValueAccCol, v_acc);
        return AssignedValue(RangeChip, ValueAccCol, offset, v_acc);
    }

    // context.rs:859-907
    AssignedValue assign_two_line_range_value(size_t offset, const std::vector<N>& v, const N& v_acc, uint64_t bits) {
        ORC_ASSERT(bits >= COMMON_RANGE_BITS * 2);
        ORC_ASSERT(bits <= COMMON_RANGE_BITS * 4);
        ensure_range_record_size(offset + 2);
        assign_fix_cell_in_range_chip(offset, AccLinesCol, n_from(2));
        assign_adv_cell_in_range_chip(offset, CommonRangeCol, v[0]);
        assign_adv_cell_in_range_chip(offset + 1, CommonRangeCol, v[1]);
        uint64_t cell_bits = bits >= 3 * COMMON_RANGE_BITS ? COMMON_RANGE_BITS : bits % COMMON_RANGE_BITS;
        assign_fix_cell_in_range_chip(offset, TagCol, n_from(cell_bits));
        assign_adv_cell_in_range_chip(offset, TaggedRangeCol, v[2]);
        cell_bits = bits > 3 * COMMON_RANGE_BITS ? bits - 3 * COMMON_RANGE_BITS : 0;
        assign_fix_cell_in_range_chip(offset + 1, TagCol, n_from(cell_bits));
        assign_adv_cell_in_range_chip(offset + 1, TaggedRangeCol, v[3]);
        assign_adv_cell_in_range_chip(offset, ValueAccCol, v_acc);
        return AssignedValue(RangeChip, ValueAccCol, offset, v_acc);
    }

    // context.rs:909-972
    AssignedValue assign_three_line_range_value(size_t offset, const std::vector<N>& v, const N& v_acc, uint64_t bits) {
        ORC_ASSERT(bits >= COMMON_RANGE_BITS * 3);
        ORC_ASSERT(bits <= COMMON_RANGE_BITS * 6);
        ensure_range_record_size(offset + 3);
        assign_fix_cell_in_range_chip(offset, AccLinesCol, n_from(3));
        assign_adv_cell_in_range_chip(offset, CommonRangeCol, v[0]);
        assign_adv_cell_in_range_chip(offset + 1, CommonRangeCol, v[1]);
        assign_adv_cell_in_range_chip(offset + 2, CommonRangeCol, v[2]);
        uint64_t cell_bits = bits >= 4 * COMMON_RANGE_BITS ? COMMON_RANGE_BITS : bits % COMMON_RANGE_BITS;
        assign_fix_cell_in_range_chip(offset, TagCol, n_from(cell_bits));
        assign_adv_cell_in_range_chip(offset, TaggedRangeCol, v[3]);
        cell_bits = bits >= 5 * COMMON_RANGE_BITS ? COMMON_RANGE_BITS : (bits > 4 * COMMON_RANGE_BITS ? bits % COMMON_RANGE_BITS : 0);
        assign_fix_cell_in_range_chip(offset + 1, TagCol, n_from(cell_bits));
        assign_adv_cell_in_range_chip(offset + 1, TaggedRangeCol, v[4]);
        cell_bits = bits > 5 * COMMON_RANGE_BITS ? bits - 5 * COMMON_RANGE_BITS : 0;
        assign_fix_cell_in_range_chip(offset + 2, TagCol, n_from(cell_bits));
        assign_adv_cell_in_range_chip(offset + 2, TaggedRangeCol, v[5]);
        assign_adv_cell_in_range_chip(offset, ValueAccCol, v_acc);
        return AssignedValue(RangeChip, ValueAccCol, offset, v_acc);
    }

    // context.rs:974-997
    std::pair<AssignedValue, size_t> assign_range_value(size_t offset, std::vector<N> v, const N& v_acc, uint64_t bits) {
        if (bits <= COMMON_RANGE_BITS) {
            return {assign_one_line_range_value(offset, v, v_acc, bits), 1};
        } else if (bits < 2 * COMMON_RANGE_BITS) {
            ORC_ASSERT(!"unreachable");
        } else if (bits <= 4 * COMMON_RANGE_BITS) {
            v.resize(4, n_from(0));
            return {assign_two_line_range_value(offset, v, v_acc, bits), 2};
        } else if (bits <= 6 * COMMON_RANGE_BITS) {
            v.resize(6, n_from(0));
            return {assign_three_line_range_value(offset, v, v_acc, bits), 3};
        }
        ORC_ASSERT(!"unreachable");
        return {AssignedValue(), 0};
    }
};

// src/context.rs:40-46, 135-158
struct Context {
    Records records;
    size_t base_offset = 0, range_offset = 0, select_offset = 0;

    Context clone_without_permutation() const {
        Context c;
        c.records.inner = records.inner;
        c.records.base_height = records.base_height;
        c.records.range_height = records.range_height;
        c.records.select_height = records.select_height;
        c.base_offset = base_offset;
        c.range_offset = range_offset;
        c.select_offset = select_offset;
        return c;
    }
};

}  // namespace orc
