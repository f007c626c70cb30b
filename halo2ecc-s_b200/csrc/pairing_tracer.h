// Symbolic mirror of the reference's extension-field tower and pairing chips (see tracer.h).
//   Fq2/Fq6/Fq12ChipOps + *BnSpecificOps   src/circuit/fq12.rs:10-459
//   PairingChipOps                          src/circuit/pairing_chip.rs:10-177
//   bn256                                   src/circuit/bn256_pairing_chip.rs:29-350, bn256_constants.rs
//   bls12_381                               src/circuit/bls12_381_pairing_chip.rs:29-287
// Every method only sequences IntegerChipOps calls, so the GPU program is the same integer
// macro-ops as everywhere else; "wasteful" calls of the reference (unused sums, squaring of one,
// fresh zero constants) are kept because they emit rows.
#pragma once
#include <array>

#include "ecc_tracer.h"

namespace h2e {

struct AssignedFq2 {
    AssignedInteger c0, c1;
};
struct AssignedFq6 {
    AssignedFq2 c0, c1, c2;
};
struct AssignedFq12 {
    AssignedFq6 c0, c1;
};
struct AssignedG2Affine {
    AssignedFq2 x, y;
    AssignedCondition z;
};
struct AssignedG2 {
    AssignedFq2 x, y, z;
};
typedef std::array<AssignedFq2, 3> LineCoeffs;
typedef std::vector<LineCoeffs> AssignedG2Prepared;

// host Fq2 = Fq[u]/(u^2+1) for shape-level constants
struct HostFq2 {
    Big c0, c1;
};
struct HostFq2Ops {
    Big p;
    HostFq2 mul(const HostFq2& a, const HostFq2& b) const {
        Big t0 = (a.c0 * b.c0) % p, t1 = (a.c1 * b.c1) % p;
        return HostFq2{(t0 + p - t1) % p, ((a.c0 * b.c1) % p + (a.c1 * b.c0) % p) % p};
    }
    HostFq2 conj(const HostFq2& a) const { return HostFq2{a.c0, a.c1.is_zero() ? a.c1 : p - a.c1}; }
    HostFq2 pow(const HostFq2& a, const Big& e) const {
        HostFq2 r{Big(1), Big(0)};
        for (unsigned i = e.bits(); i-- > 0;) {
            r = mul(r, r);
            if (e.bit(i)) r = mul(r, a);
        }
        return r;
    }
};

// Frobenius / twist coefficients, derived from their definitions:
//   FROBENIUS_COEFF_FQ6_C1[i] = xi^((p^i-1)/3), _FQ6_C2[i] = xi^(2(p^i-1)/3), _FQ12_C1[i] = xi^((p^i-1)/6),
//   XI_TO_Q_MINUS_1_OVER_2 = xi^((p-1)/2)   (bn256_constants.rs:14-383; xi = 9+u)
//   bls12_381 (xi = 1+u): the three from_raw_unchecked coefficients of bls12_381_pairing_chip.rs:58-107
struct TowerConsts {
    bool is_bn;
    HostFq2 fq2_c1[2], fq6_c1[6], fq6_c2[6], fq12_c1[12], xi_to_q_minus_1_over_2;
    explicit TowerConsts(bool bn) : is_bn(bn) {
        Big p = modulus_of(bn ? F_BN256_FQ : F_BLS12_381_FQ);
        HostFq2Ops m{p};
        HostFq2 xi = bn ? HostFq2{Big(9), Big(1)} : HostFq2{Big(1), Big(1)};
        Big pm1 = p - Big(1);
        HostFq2 one{Big(1), Big(0)};
        fq2_c1[0] = one;
        fq2_c1[1] = HostFq2{pm1, Big(0)};
        HostFq2 g6 = m.pow(xi, pm1 / Big(6)), g3 = m.mul(g6, g6);
        fq12_c1[0] = one;
        fq6_c1[0] = one;
        for (int i = 1; i < 12; i++) fq12_c1[i] = m.mul(m.conj(fq12_c1[i - 1]), g6);
        for (int i = 1; i < 6; i++) fq6_c1[i] = m.mul(m.conj(fq6_c1[i - 1]), g3);
        for (int i = 0; i < 6; i++) fq6_c2[i] = m.mul(fq6_c1[i], fq6_c1[i]);
        xi_to_q_minus_1_over_2 = m.mul(g3, g6);
    }
};

static const int8_t SIX_U_PLUS_2_NAF[65] = {0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0, 1, 1, 1,
                                            0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, 1, 1};
static const uint64_t BN_X = 4965661367192848881ull;
static const uint64_t BLS_X = 0xd201000000010000ull;

class PairingOps {
   public:
    EccContext& e;
    IntegerContext& ic;
    bool is_bn;
    TowerConsts k;
    PairingOps(EccContext& ec, bool bn) : e(ec), ic(ec.base), is_bn(bn), k(bn) {}

    // ---------------- Fq2ChipOps (fq12.rs:24-104) ----------------
    AssignedFq2 fq2_reduce(const AssignedFq2& x) {
        AssignedInteger a = ic.reduce(x.c0);
        return AssignedFq2{a, ic.reduce(x.c1)};
    }
    void fq2_assert_equal(const AssignedFq2& x, const AssignedFq2& y) {
        ic.assert_int_equal(x.c0, y.c0);
        ic.assert_int_equal(x.c1, y.c1);
    }
    AssignedFq2 fq2_assign_zero() {
        AssignedInteger z = ic.assign_int_constant(Big(0));
        return AssignedFq2{z, z};
    }
    AssignedFq2 fq2_assign_one() {
        AssignedInteger a = ic.assign_int_constant(Big(1));
        return AssignedFq2{a, ic.assign_int_constant(Big(0))};
    }
    AssignedFq2 fq2_assign_constant(const HostFq2& c) {
        AssignedInteger a = ic.assign_int_constant(c.c0);
        return AssignedFq2{a, ic.assign_int_constant(c.c1)};
    }
    // per-instance constant (the tests pass G2 points as circuit constants); cells: 2 logical inputs
    AssignedFq2 fq2_assign_constant_input(uint32_t c0_cell, uint32_t c1_cell) {
        AssignedInteger a = ic.assign_int_constant_input(c0_cell);
        return AssignedFq2{a, ic.assign_int_constant_input(c1_cell)};
    }
    AssignedFq2 fq2_add(const AssignedFq2& a, const AssignedFq2& b) {
        AssignedInteger x = ic.int_add(a.c0, b.c0);
        return AssignedFq2{x, ic.int_add(a.c1, b.c1)};
    }
    AssignedFq2 fq2_sub(const AssignedFq2& a, const AssignedFq2& b) {
        AssignedInteger x = ic.int_sub(a.c0, b.c0);
        return AssignedFq2{x, ic.int_sub(a.c1, b.c1)};
    }
    AssignedFq2 fq2_double(const AssignedFq2& a) {
        AssignedInteger x = ic.int_add(a.c0, a.c0);
        return AssignedFq2{x, ic.int_add(a.c1, a.c1)};
    }
    AssignedFq2 fq2_neg(const AssignedFq2& a) {
        AssignedInteger x = ic.int_neg(a.c0);
        return AssignedFq2{x, ic.int_neg(a.c1)};
    }
    AssignedFq2 fq2_conjugate(const AssignedFq2& a) { return AssignedFq2{a.c0, ic.int_neg(a.c1)}; }
    AssignedFq2 fq2_mul(const AssignedFq2& a, const AssignedFq2& b) {
        AssignedInteger ab00 = ic.int_mul(a.c0, b.c0);
        AssignedInteger ab11 = ic.int_mul(a.c1, b.c1);
        AssignedInteger c0 = ic.int_sub(ab00, ab11);
        AssignedInteger a01 = ic.int_add(a.c0, a.c1);
        AssignedInteger b01 = ic.int_add(b.c0, b.c1);
        AssignedInteger c1 = ic.int_mul(a01, b01);
        c1 = ic.int_sub(c1, ab00);
        c1 = ic.int_sub(c1, ab11);
        return AssignedFq2{c0, c1};
    }
    AssignedFq2 fq2_square(const AssignedFq2& a) { return fq2_mul(a, a); }
    AssignedFq2 fq2_unsafe_invert(const AssignedFq2& x) {
        AssignedInteger t0 = ic.int_square(x.c0);
        AssignedInteger t1 = ic.int_square(x.c1);
        t0 = ic.int_add(t0, t1);
        AssignedInteger t = ic.int_unsafe_invert(t0);
        AssignedInteger c0 = ic.int_mul(x.c0, t);
        AssignedInteger c1 = ic.int_mul(x.c1, t);
        return AssignedFq2{c0, ic.int_neg(c1)};
    }

    // ---------------- curve-specific pieces ----------------
    // bn256_pairing_chip.rs:32-45 (xi = 9+u) / bls12_381_pairing_chip.rs:32-37 (xi = 1+u)
    AssignedFq2 fq2_mul_by_nonresidue(const AssignedFq2& a) {
        if (is_bn) {
            AssignedFq2 a2 = fq2_double(a);
            AssignedFq2 a4 = fq2_double(a2);
            AssignedFq2 a8 = fq2_double(a4);
            AssignedInteger t = ic.int_add(a8.c0, a.c0);
            AssignedInteger c0 = ic.int_sub(t, a.c1);
            t = ic.int_add(a8.c1, a.c0);
            return AssignedFq2{c0, ic.int_add(t, a.c1)};
        }
        AssignedInteger c0 = ic.int_sub(a.c0, a.c1);
        return AssignedFq2{c0, ic.int_add(a.c0, a.c1)};
    }
    AssignedFq2 fq2_frobenius_map(const AssignedFq2& x, size_t power) {
        if (!is_bn) return fq2_conjugate(x);
        AssignedInteger v = ic.assign_int_constant(k.fq2_c1[power % 2].c0);
        return AssignedFq2{x.c0, ic.int_mul(x.c1, v)};
    }
    AssignedFq6 fq6_mul_by_nonresidue(const AssignedFq6& a) { return AssignedFq6{fq2_mul_by_nonresidue(a.c2), a.c0, a.c1}; }
    AssignedFq6 fq6_frobenius_map(const AssignedFq6& x, size_t power) {
        AssignedFq2 c0 = fq2_frobenius_map(x.c0, power);
        AssignedFq2 c1 = fq2_frobenius_map(x.c1, power);
        AssignedFq2 c2 = fq2_frobenius_map(x.c2, power);
        // bls12_381 ignores `power` and always uses the p^1 coefficients (bls12_381_pairing_chip.rs:52-83)
        size_t i = is_bn ? power % 6 : 1;
        AssignedFq2 coeff_c1 = fq2_assign_constant(k.fq6_c1[i]);
        c1 = fq2_mul(c1, coeff_c1);
        AssignedFq2 coeff_c2 = fq2_assign_constant(k.fq6_c2[i]);
        c2 = fq2_mul(c2, coeff_c2);
        return AssignedFq6{c0, c1, c2};
    }
    AssignedFq12 fq12_frobenius_map(const AssignedFq12& x, size_t power) {
        AssignedFq6 c0 = fq6_frobenius_map(x.c0, power);
        AssignedFq6 c1 = fq6_frobenius_map(x.c1, power);
        AssignedFq2 coeff = fq2_assign_constant(k.fq12_c1[is_bn ? power % 12 : 1]);
        AssignedFq2 a = fq2_mul(c1.c0, coeff);
        AssignedFq2 b = fq2_mul(c1.c1, coeff);
        AssignedFq2 c = fq2_mul(c1.c2, coeff);
        return AssignedFq12{c0, AssignedFq6{a, b, c}};
    }

    // ---------------- Fq6ChipOps (fq12.rs:106-287) ----------------
    template <class F>
    AssignedFq6 map3(const AssignedFq6& a, const AssignedFq6& b, F f) {
        AssignedFq2 x = f(a.c0, b.c0);
        AssignedFq2 y = f(a.c1, b.c1);
        AssignedFq2 z = f(a.c2, b.c2);
        return AssignedFq6{x, y, z};
    }
    AssignedFq6 fq6_add(const AssignedFq6& a, const AssignedFq6& b) {
        return map3(a, b, [&](const AssignedFq2& p, const AssignedFq2& q) { return fq2_add(p, q); });
    }
    AssignedFq6 fq6_sub(const AssignedFq6& a, const AssignedFq6& b) {
        return map3(a, b, [&](const AssignedFq2& p, const AssignedFq2& q) { return fq2_sub(p, q); });
    }
    AssignedFq6 fq6_neg(const AssignedFq6& a) {
        AssignedFq2 x = fq2_neg(a.c0);
        AssignedFq2 y = fq2_neg(a.c1);
        return AssignedFq6{x, y, fq2_neg(a.c2)};
    }
    void fq6_assert_equal(const AssignedFq6& x, const AssignedFq6& y) {
        fq2_assert_equal(x.c0, y.c0);
        fq2_assert_equal(x.c1, y.c1);
        fq2_assert_equal(x.c2, y.c2);
    }
    AssignedFq6 fq6_assign_zero() {
        AssignedFq2 z = fq2_assign_zero();
        return AssignedFq6{z, z, z};
    }
    AssignedFq6 fq6_assign_one() {
        AssignedFq2 one = fq2_assign_one();
        AssignedFq2 z = fq2_assign_zero();
        return AssignedFq6{one, z, z};
    }
    AssignedFq6 fq6_mul(const AssignedFq6& a, const AssignedFq6& b) {
        AssignedFq2 ab00 = fq2_mul(a.c0, b.c0);
        AssignedFq2 ab11 = fq2_mul(a.c1, b.c1);
        AssignedFq2 ab22 = fq2_mul(a.c2, b.c2);
        AssignedFq2 b12 = fq2_add(b.c1, b.c2);
        AssignedFq2 a12 = fq2_add(a.c1, a.c2);
        AssignedFq2 t = fq2_mul(a12, b12);
        t = fq2_sub(t, ab11);
        t = fq2_sub(t, ab22);
        t = fq2_mul_by_nonresidue(t);
        AssignedFq2 c0 = fq2_add(t, ab00);
        AssignedFq2 b01 = fq2_add(b.c0, b.c1);
        AssignedFq2 a01 = fq2_add(a.c0, a.c1);
        t = fq2_mul(a01, b01);
        t = fq2_sub(t, ab00);
        t = fq2_sub(t, ab11);
        AssignedFq2 ab22n = fq2_mul_by_nonresidue(ab22);
        AssignedFq2 c1 = fq2_add(t, ab22n);
        AssignedFq2 b02 = fq2_add(b.c0, b.c2);
        AssignedFq2 a02 = fq2_add(a.c0, a.c2);
        t = fq2_mul(a02, b02);
        t = fq2_sub(t, ab00);
        t = fq2_add(t, ab11);
        AssignedFq2 c2 = fq2_sub(t, ab22);
        return AssignedFq6{c0, c1, c2};
    }
    AssignedFq6 fq6_square(const AssignedFq6& a) { return fq6_mul(a, a); }
    AssignedFq6 fq6_mul_by_1(const AssignedFq6& a, const AssignedFq2& b1) {
        AssignedFq2 ab11 = fq2_mul(a.c1, b1);
        AssignedFq2 a12 = fq2_add(a.c1, a.c2);
        AssignedFq2 t = fq2_mul(a12, b1);
        t = fq2_sub(t, ab11);
        AssignedFq2 c0 = fq2_mul_by_nonresidue(t);
        AssignedFq2 a01 = fq2_add(a.c0, a.c1);
        t = fq2_mul(a01, b1);
        AssignedFq2 c1 = fq2_sub(t, ab11);
        return AssignedFq6{c0, c1, ab11};
    }
    AssignedFq6 fq6_mul_by_01(const AssignedFq6& a, const AssignedFq2& b0, const AssignedFq2& b1) {
        AssignedFq2 ab00 = fq2_mul(a.c0, b0);
        AssignedFq2 ab11 = fq2_mul(a.c1, b1);
        AssignedFq2 a12 = fq2_add(a.c1, a.c2);
        AssignedFq2 t = fq2_mul(a12, b1);
        t = fq2_sub(t, ab11);
        t = fq2_mul_by_nonresidue(t);
        AssignedFq2 c0 = fq2_add(t, ab00);
        AssignedFq2 b01 = fq2_add(b0, b1);
        AssignedFq2 a01 = fq2_add(a.c0, a.c1);
        t = fq2_mul(a01, b01);
        t = fq2_sub(t, ab00);
        AssignedFq2 c1 = fq2_sub(t, ab11);
        AssignedFq2 a02 = fq2_add(a.c0, a.c2);
        t = fq2_mul(a02, b0);
        t = fq2_sub(t, ab00);
        AssignedFq2 c2 = fq2_add(t, ab11);
        return AssignedFq6{c0, c1, c2};
    }
    AssignedFq6 fq6_unsafe_invert(const AssignedFq6& x) {
        AssignedFq2 c0 = fq2_mul_by_nonresidue(x.c2);
        c0 = fq2_mul(c0, x.c1);
        c0 = fq2_neg(c0);
        AssignedFq2 x0s = fq2_square(x.c0);
        c0 = fq2_add(c0, x0s);
        AssignedFq2 c1 = fq2_square(x.c2);
        c1 = fq2_mul_by_nonresidue(c1);
        AssignedFq2 x01 = fq2_mul(x.c0, x.c1);
        c1 = fq2_sub(c1, x01);
        AssignedFq2 c2 = fq2_square(x.c1);
        AssignedFq2 x02 = fq2_mul(x.c0, x.c2);
        c2 = fq2_sub(c2, x02);
        AssignedFq2 c0x0 = fq2_mul(c0, x.c0);
        AssignedFq2 c1x2 = fq2_mul(c1, x.c2);
        AssignedFq2 c2x1 = fq2_mul(c2, x.c1);
        AssignedFq2 t = fq2_add(c1x2, c2x1);
        t = fq2_mul_by_nonresidue(t);
        t = fq2_add(t, c0x0);
        t = fq2_unsafe_invert(t);
        AssignedFq2 r0 = fq2_mul(t, c0);
        AssignedFq2 r1 = fq2_mul(t, c1);
        return AssignedFq6{r0, r1, fq2_mul(t, c2)};
    }

    // ---------------- Fq12ChipOps (fq12.rs:289-459) ----------------
    AssignedFq12 fq12_assign_one() {
        AssignedFq6 one = fq6_assign_one();
        return AssignedFq12{one, fq6_assign_zero()};
    }
    void fq12_assert_eq(const AssignedFq12& x, const AssignedFq12& y) {
        fq6_assert_equal(x.c0, y.c0);
        fq6_assert_equal(x.c1, y.c1);
    }
    void fq12_assert_one(const AssignedFq12& x) {
        AssignedFq12 one = fq12_assign_one();
        fq12_assert_eq(x, one);
    }
    AssignedFq12 fq12_mul(const AssignedFq12& a, const AssignedFq12& b) {
        AssignedFq6 ab00 = fq6_mul(a.c0, b.c0);
        AssignedFq6 ab11 = fq6_mul(a.c1, b.c1);
        AssignedFq6 a01 = fq6_add(a.c0, a.c1);
        AssignedFq6 b01 = fq6_add(b.c0, b.c1);
        AssignedFq6 c1 = fq6_mul(a01, b01);
        c1 = fq6_sub(c1, ab00);
        c1 = fq6_sub(c1, ab11);
        AssignedFq6 ab11n = fq6_mul_by_nonresidue(ab11);
        return AssignedFq12{fq6_add(ab00, ab11n), c1};
    }
    AssignedFq12 fq12_square(const AssignedFq12& a) { return fq12_mul(a, a); }
    AssignedFq12 fq12_conjugate(const AssignedFq12& x) { return AssignedFq12{x.c0, fq6_neg(x.c1)}; }
    AssignedFq12 fq12_mul_by_014(const AssignedFq12& x, const AssignedFq2& c0, const AssignedFq2& c1, const AssignedFq2& c4) {
        AssignedFq6 t0 = fq6_mul_by_01(x.c0, c0, c1);
        AssignedFq6 t1 = fq6_mul_by_1(x.c1, c4);
        AssignedFq2 o = fq2_add(c1, c4);
        AssignedFq6 x0 = fq6_mul_by_nonresidue(t1);
        x0 = fq6_add(x0, t0);
        AssignedFq6 x1 = fq6_add(x.c0, x.c1);
        x1 = fq6_mul_by_01(x1, c0, o);
        x1 = fq6_sub(x1, t0);
        x1 = fq6_sub(x1, t1);
        return AssignedFq12{x0, x1};
    }
    AssignedFq12 fq12_mul_by_034(const AssignedFq12& x, const AssignedFq2& c0, const AssignedFq2& c3, const AssignedFq2& c4) {
        AssignedFq2 t00 = fq2_mul(x.c0.c0, c0);
        AssignedFq2 t01 = fq2_mul(x.c0.c1, c0);
        AssignedFq2 t02 = fq2_mul(x.c0.c2, c0);
        AssignedFq6 t0{t00, t01, t02};
        AssignedFq6 t1 = fq6_mul_by_01(x.c1, c3, c4);
        AssignedFq6 t2 = fq6_add(x.c0, x.c1);
        AssignedFq2 o = fq2_add(c0, c3);
        t2 = fq6_mul_by_01(t2, o, c4);
        t2 = fq6_sub(t2, t0);
        AssignedFq6 x1 = fq6_sub(t2, t1);
        t1 = fq6_mul_by_nonresidue(t1);
        return AssignedFq12{fq6_add(t0, t1), x1};
    }
    void fp4_square(AssignedFq2& c0, AssignedFq2& c1, const AssignedFq2& a0, const AssignedFq2& a1) {
        AssignedFq2 t0 = fq2_square(a0);
        AssignedFq2 t1 = fq2_square(a1);
        AssignedFq2 t2 = fq2_mul_by_nonresidue(t1);
        c0 = fq2_add(t2, t0);
        t2 = fq2_add(a0, a1);
        t2 = fq2_square(t2);
        t2 = fq2_sub(t2, t0);
        c1 = fq2_sub(t2, t1);
    }
    AssignedFq12 fq12_cyclotomic_square(const AssignedFq12& x) {
        AssignedFq2 zero = fq2_assign_zero();  // fq12.rs:406: a fresh zero constant per call
        AssignedFq2 t3 = zero, t4 = zero, t5 = zero, t6 = zero;
        fp4_square(t3, t4, x.c0.c0, x.c1.c1);
        AssignedFq2 t2 = fq2_sub(t3, x.c0.c0);
        t2 = fq2_double(t2);
        AssignedFq2 c00 = fq2_add(t2, t3);
        t2 = fq2_add(t4, x.c1.c1);
        t2 = fq2_double(t2);
        AssignedFq2 c11 = fq2_add(t2, t4);
        fp4_square(t3, t4, x.c1.c0, x.c0.c2);
        fp4_square(t5, t6, x.c0.c1, x.c1.c2);
        t2 = fq2_sub(t3, x.c0.c1);
        t2 = fq2_double(t2);
        AssignedFq2 c01 = fq2_add(t2, t3);
        t2 = fq2_add(t4, x.c1.c2);
        t2 = fq2_double(t2);
        AssignedFq2 c12 = fq2_add(t2, t4);
        t3 = fq2_mul_by_nonresidue(t6);
        t2 = fq2_add(t3, x.c1.c0);
        t2 = fq2_double(t2);
        AssignedFq2 c10 = fq2_add(t2, t3);
        t2 = fq2_sub(t5, x.c0.c2);
        t2 = fq2_double(t2);
        AssignedFq2 c02 = fq2_add(t2, t5);
        return AssignedFq12{AssignedFq6{c00, c01, c02}, AssignedFq6{c10, c11, c12}};
    }
    AssignedFq12 fq12_unsafe_invert(const AssignedFq12& x) {
        AssignedFq6 x0s = fq6_square(x.c0);
        AssignedFq6 x1s = fq6_square(x.c1);
        AssignedFq6 t = fq6_mul_by_nonresidue(x1s);
        t = fq6_sub(x0s, t);
        t = fq6_unsafe_invert(t);
        AssignedFq6 c0 = fq6_mul(t, x.c0);
        AssignedFq6 c1 = fq6_mul(t, x.c1);
        return AssignedFq12{c0, fq6_neg(c1)};
    }

    // ---------------- PairingChipOps (pairing_chip.rs:13-133) ----------------
    LineCoeffs doubling_step(AssignedG2& pt) {
        AssignedFq2 x2 = fq2_square(pt.x);
        AssignedFq2 y2 = fq2_square(pt.y);
        AssignedFq2 _2y2 = fq2_double(y2);
        AssignedFq2 _4y2 = fq2_double(_2y2);
        AssignedFq2 _4y4 = fq2_square(_2y2);
        AssignedFq2 _8y4 = fq2_double(_4y4);
        AssignedFq2 z2 = fq2_square(pt.z);
        AssignedFq2 _4xy2 = fq2_mul(y2, pt.x);
        _4xy2 = fq2_double(_4xy2);
        _4xy2 = fq2_double(_4xy2);
        AssignedFq2 _3x2 = fq2_double(x2);
        _3x2 = fq2_add(_3x2, x2);
        AssignedFq2 _6x2 = fq2_double(_3x2);
        AssignedFq2 _9x4 = fq2_square(_3x2);
        fq2_add(_3x2, pt.x);  // `_3x2_x`: computed and never used (pairing_chip.rs:38) -- the rows exist
        AssignedFq2 rx = fq2_sub(_9x4, _4xy2);
        rx = fq2_sub(rx, _4xy2);
        AssignedFq2 ry = fq2_sub(_4xy2, rx);
        ry = fq2_mul(ry, _3x2);
        ry = fq2_sub(ry, _8y4);
        AssignedFq2 rz = fq2_mul(pt.y, pt.z);
        rz = fq2_double(rz);
        AssignedFq2 c0 = fq2_mul(z2, rz);
        c0 = fq2_double(c0);
        AssignedFq2 c1 = fq2_mul(z2, _6x2);
        c1 = fq2_neg(c1);
        AssignedFq2 c2 = fq2_mul(_6x2, pt.x);
        c2 = fq2_sub(c2, _4y2);
        pt = AssignedG2{rx, ry, rz};
        return {c0, c1, c2};
    }
    LineCoeffs addition_step(AssignedG2& pt, const AssignedG2Affine& pq) {
        AssignedFq2 zt2 = fq2_square(pt.z);
        AssignedFq2 yqzt = fq2_mul(pq.y, pt.z);
        AssignedFq2 yqzt3 = fq2_mul(yqzt, zt2);
        AssignedFq2 yqzt3_yt = fq2_sub(yqzt3, pt.y);
        AssignedFq2 _2yqzt3_2yt = fq2_double(yqzt3_yt);
        AssignedFq2 xqzt2 = fq2_mul(pq.x, zt2);
        AssignedFq2 xqzt2_xt = fq2_sub(xqzt2, pt.x);
        AssignedFq2 _2_xqzt2_xt = fq2_double(xqzt2_xt);
        AssignedFq2 _4_xqzt2_xt_2 = fq2_square(_2_xqzt2_xt);
        AssignedFq2 t0 = fq2_mul(_4_xqzt2_xt_2, xqzt2_xt);
        AssignedFq2 t1 = fq2_double(_4_xqzt2_xt_2);
        AssignedFq2 t2 = fq2_mul(t1, pt.x);
        AssignedFq2 t = fq2_square(_2yqzt3_2yt);
        t = fq2_sub(t, t0);
        AssignedFq2 rx = fq2_sub(t, t2);
        t0 = fq2_mul(_4_xqzt2_xt_2, pt.x);
        t0 = fq2_sub(t0, rx);
        t0 = fq2_mul(_2yqzt3_2yt, t0);
        t1 = fq2_mul(_2_xqzt2_xt, _4_xqzt2_xt_2);
        t1 = fq2_mul(t1, pt.y);
        AssignedFq2 ry = fq2_sub(t0, t1);
        AssignedFq2 rz = fq2_mul(pt.z, _2_xqzt2_xt);
        AssignedFq2 c0 = fq2_double(rz);
        AssignedFq2 c1 = fq2_double(_2yqzt3_2yt);
        c1 = fq2_neg(c1);
        t0 = fq2_double(_2yqzt3_2yt);
        t0 = fq2_mul(t0, pq.x);
        t1 = fq2_mul(pq.y, rz);
        t1 = fq2_double(t1);
        AssignedFq2 c2 = fq2_sub(t0, t1);
        pt = AssignedG2{rx, ry, rz};
        return {c0, c1, c2};
    }
    AssignedG2 g2affine_to_g2(const AssignedG2Affine& g2) {
        e.ctx->assert_false(g2.z);
        return AssignedG2{g2.x, g2.y, fq2_assign_one()};
    }
    AssignedG2Affine g2_neg(const AssignedG2Affine& g2) { return AssignedG2Affine{g2.x, fq2_neg(g2.y), g2.z}; }

    // ---------------- prepare_g2 ----------------
    AssignedG2Prepared prepare_g2(const AssignedG2Affine& g2) {
        AssignedG2Prepared coeffs;
        if (is_bn) {  // bn256_pairing_chip.rs:104-155
            AssignedG2Affine neg_g2 = g2_neg(g2);
            AssignedG2 r = g2affine_to_g2(g2);
            for (int i = 64; i >= 1; i--) {
                coeffs.push_back(doubling_step(r));
                int x = SIX_U_PLUS_2_NAF[i - 1];
                if (x == 1) coeffs.push_back(addition_step(r, g2));
                if (x == -1) coeffs.push_back(addition_step(r, neg_g2));
            }
            AssignedG2Affine q1 = g2;
            AssignedFq2 c11 = fq2_assign_constant(k.fq6_c1[1]);
            AssignedFq2 c12 = fq2_assign_constant(k.fq6_c1[2]);
            AssignedFq2 xi = fq2_assign_constant(k.xi_to_q_minus_1_over_2);
            q1.x.c1 = ic.int_neg(q1.x.c1);
            q1.x = fq2_mul(q1.x, c11);
            q1.y.c1 = ic.int_neg(q1.y.c1);
            q1.y = fq2_mul(q1.y, xi);
            coeffs.push_back(addition_step(r, q1));
            AssignedG2Affine minusq2 = g2;
            minusq2.x = fq2_mul(minusq2.x, c12);
            coeffs.push_back(addition_step(r, minusq2));
        } else {  // bls12_381_pairing_chip.rs:165-189
            AssignedG2 f = g2affine_to_g2(g2);
            bool found_one = false;
            for (int b = 63; b >= 0; b--) {
                bool i = ((BLS_X >> 1) >> b) & 1;
                if (!found_one) {
                    found_one = i;
                    continue;
                }
                coeffs.push_back(doubling_step(f));
                if (i) coeffs.push_back(addition_step(f, g2));
            }
            coeffs.push_back(doubling_step(f));
        }
        return coeffs;
    }
    // bn256_pairing_chip.rs:157-174 / bls12_381_pairing_chip.rs:123-140
    AssignedFq12 ell(const AssignedFq12& f, const LineCoeffs& c, const AssignedPoint& p) {
        AssignedInteger c00 = ic.int_mul(c[0].c0, p.y);
        AssignedInteger c01 = ic.int_mul(c[0].c1, p.y);
        AssignedInteger c10 = ic.int_mul(c[1].c0, p.x);
        AssignedInteger c11 = ic.int_mul(c[1].c1, p.x);
        if (is_bn) return fq12_mul_by_034(f, AssignedFq2{c00, c01}, AssignedFq2{c10, c11}, c[2]);
        return fq12_mul_by_014(f, c[2], AssignedFq2{c10, c11}, AssignedFq2{c00, c01});
    }
    typedef std::vector<std::pair<const AssignedPoint*, const AssignedG2Prepared*>> Terms;
    AssignedFq12 multi_miller_loop(const Terms& terms) {
        std::vector<size_t> it(terms.size(), 0);
        for (auto& t : terms) e.ctx->assert_false(t.first->z);
        AssignedFq12 f = fq12_assign_one();
        auto ell_all = [&]() {
            for (size_t j = 0; j < terms.size(); j++) f = ell(f, terms[j].second->at(it[j]++), *terms[j].first);
        };
        if (is_bn) {  // bn256_pairing_chip.rs:176-228
            for (int i = 64; i >= 1; i--) {
                if (i != 64) f = fq12_square(f);
                ell_all();
                if (SIX_U_PLUS_2_NAF[i - 1] != 0) ell_all();
            }
            ell_all();
            ell_all();
        } else {  // bls12_381_pairing_chip.rs:191-234
            bool found_one = false;
            for (int b = 63; b >= 0; b--) {
                bool i = ((BLS_X >> 1) >> b) & 1;
                if (!found_one) {
                    found_one = i;
                    continue;
                }
                ell_all();
                if (i) ell_all();
                f = fq12_square(f);
            }
            ell_all();
            f = fq12_conjugate(f);
        }
        for (size_t j = 0; j < terms.size(); j++)
            if (it[j] != terms[j].second->size()) throw std::logic_error("unused line coefficients");
        return f;
    }
    // bn256_pairing_chip.rs:230-240
    AssignedFq12 exp_by_x(const AssignedFq12& f) {
        AssignedFq12 res = fq12_assign_one();
        for (int i = 63; i >= 0; i--) {
            res = fq12_cyclotomic_square(res);
            if ((BN_X >> i) & 1) res = fq12_mul(res, f);
        }
        return res;
    }
    // bls12_381_pairing_chip.rs:142-159
    AssignedFq12 cyclotomic_exp(const AssignedFq12& f) {
        AssignedFq12 tmp = fq12_assign_one();
        bool found_one = false;
        for (int b = 63; b >= 0; b--) {
            bool i = (BLS_X >> b) & 1;
            if (found_one)
                tmp = fq12_cyclotomic_square(tmp);
            else
                found_one = i;
            if (i) tmp = fq12_mul(tmp, f);
        }
        return fq12_conjugate(tmp);
    }
    AssignedFq12 final_exponentiation(const AssignedFq12& f) {
        if (is_bn) {  // bn256_pairing_chip.rs:242-323
            AssignedFq12 f1 = fq12_conjugate(f);
            AssignedFq12 f2 = fq12_unsafe_invert(f);
            AssignedFq12 r = fq12_mul(f1, f2);
            f2 = r;
            r = fq12_frobenius_map(r, 2);
            r = fq12_mul(r, f2);
            AssignedFq12 fp = fq12_frobenius_map(r, 1);
            AssignedFq12 fp2 = fq12_frobenius_map(r, 2);
            AssignedFq12 fp3 = fq12_frobenius_map(fp2, 1);
            AssignedFq12 fu = exp_by_x(r);
            AssignedFq12 fu2 = exp_by_x(fu);
            AssignedFq12 fu3 = exp_by_x(fu2);
            AssignedFq12 y3 = fq12_frobenius_map(fu, 1);
            AssignedFq12 fu2p = fq12_frobenius_map(fu2, 1);
            AssignedFq12 fu3p = fq12_frobenius_map(fu3, 1);
            AssignedFq12 y2 = fq12_frobenius_map(fu2, 2);
            AssignedFq12 y0 = fq12_mul(fp, fp2);
            y0 = fq12_mul(y0, fp3);
            AssignedFq12 y1 = fq12_conjugate(r);
            AssignedFq12 y5 = fq12_conjugate(fu2);
            y3 = fq12_conjugate(y3);
            AssignedFq12 y4 = fq12_mul(fu, fu2p);
            y4 = fq12_conjugate(y4);
            AssignedFq12 y6 = fq12_mul(fu3, fu3p);
            y6 = fq12_conjugate(y6);
            y6 = fq12_cyclotomic_square(y6);
            y6 = fq12_mul(y6, y4);
            y6 = fq12_mul(y6, y5);
            AssignedFq12 t1 = fq12_mul(y3, y5);
            t1 = fq12_mul(t1, y6);
            y6 = fq12_mul(y6, y2);
            t1 = fq12_cyclotomic_square(t1);
            t1 = fq12_mul(t1, y6);
            t1 = fq12_cyclotomic_square(t1);
            AssignedFq12 t0 = fq12_mul(t1, y1);
            t1 = fq12_mul(t1, y0);
            t0 = fq12_cyclotomic_square(t0);
            return fq12_mul(t0, t1);
        }
        // bls12_381_pairing_chip.rs:236-286
        AssignedFq12 t0 = fq12_frobenius_map(f, 1);
        for (int i = 0; i < 5; i++) t0 = fq12_frobenius_map(t0, 1);
        AssignedFq12 t1 = fq12_unsafe_invert(f);
        AssignedFq12 t2 = fq12_mul(t0, t1);
        t1 = t2;
        t2 = fq12_frobenius_map(t2, 1);
        t2 = fq12_frobenius_map(t2, 1);
        t2 = fq12_mul(t2, t1);
        t1 = fq12_cyclotomic_square(t2);
        t1 = fq12_conjugate(t1);
        AssignedFq12 t3 = cyclotomic_exp(t2);
        AssignedFq12 t4 = fq12_cyclotomic_square(t3);
        AssignedFq12 t5 = fq12_mul(t1, t3);
        t1 = cyclotomic_exp(t5);
        t0 = cyclotomic_exp(t1);
        AssignedFq12 t6 = cyclotomic_exp(t0);
        t6 = fq12_mul(t6, t4);
        t4 = cyclotomic_exp(t6);
        t5 = fq12_conjugate(t5);
        AssignedFq12 t = fq12_mul(t5, t2);
        t4 = fq12_mul(t4, t);
        t5 = fq12_conjugate(t2);
        t1 = fq12_mul(t1, t2);
        for (int i = 0; i < 3; i++) t1 = fq12_frobenius_map(t1, 1);
        t6 = fq12_mul(t6, t5);
        t6 = fq12_frobenius_map(t6, 1);
        t3 = fq12_mul(t3, t0);
        for (int i = 0; i < 2; i++) t3 = fq12_frobenius_map(t3, 1);
        t3 = fq12_mul(t3, t1);
        t3 = fq12_mul(t3, t6);
        return fq12_mul(t3, t4);
    }
    // pairing_chip.rs:157-176
    AssignedFq12 pairing(const std::vector<std::pair<const AssignedPoint*, const AssignedG2Affine*>>& terms) {
        std::vector<AssignedG2Prepared> prepared;
        for (auto& t : terms) prepared.push_back(prepare_g2(*t.second));
        Terms pt;
        for (size_t i = 0; i < terms.size(); i++) pt.push_back({terms[i].first, &prepared[i]});
        AssignedFq12 res = multi_miller_loop(pt);
        return final_exponentiation(res);
    }
    void check_pairing(const std::vector<std::pair<const AssignedPoint*, const AssignedG2Affine*>>& terms) {
        AssignedFq12 res = pairing(terms);
        fq12_assert_one(res);
    }
};

}  // namespace h2e
