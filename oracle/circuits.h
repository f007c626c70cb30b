// ORACLE (test infrastructure, NOT product code). Whole-circuit harness shapes = the bodies of the
// reference's own tests:
//   kind 0/1  bn256 G1 MSM with / without select chip   src/tests/native_scalar_ecc_chip.rs:13-110
//   kind 2    bn256 check_pairing (second block)         src/tests/native_scalar_pairing_chip.rs:67-97
//   kind 3    bls12_381 check_pairing (second block)     src/tests/general_scalar_pairing_chip.rs:74-105
//   kind 4    bls12_381 G1 MSM, general scalar           src/tests/general_scalar_ecc_chip.rs:14-49
//   kind 5/6  bn256 / bls12_381 pairing(a,b) alone (first blocks, value returned for KATs)
#pragma once
#include "pairing.h"

namespace orc {

// input layout for MSM kinds: [x_i, y_i, z_i] * n, [s_i] * n, r1.x, r1.y, r2.x, r2.y, exp.x, exp.y, exp.z  (z = identity flag)
inline int run_msm(int kind, size_t n, const std::vector<BN>& in, std::shared_ptr<Context> ctx) {
    bool bls = (kind == 4);
    EccContext e = bls ? EccContext::general(ctx, bls12_381_g1(), BLS12_381_FQ()) : EccContext::native(ctx, bn256_g1(), BN256_FQ(), kind == 0);
    ORC_ASSERT(in.size() == 4 * n + 7);
    try {
        std::vector<AssignedPoint> points;
        for (size_t i = 0; i < n; i++) points.push_back(e.assign_point(HostPoint{in[3 * i], in[3 * i + 1], !in[3 * i + 2].is_zero()}));
        std::vector<AssignedScalar> scalars;
        for (size_t i = 0; i < n; i++) {
            AssignedScalar s;
            if (bls)
                s.i = e.scalar->assign_w(in[3 * n + i]);
            else
                s.v = e.bc().assign(bn_to_n(in[3 * n + i]));
            scalars.push_back(s);
        }
        const BN* t = &in[4 * n];
        HostPoint r1{t[0], t[1], false}, r2{t[2], t[3], false};
        AssignedPoint res = e.msm_unsafe(points, scalars, r1, r2);
        AssignedPoint res_expect = e.assign_point(HostPoint{t[4], t[5], !t[6].is_zero()});
        e.ecc_assert_equal(res, res_expect);
    } catch (UnsafeError& u) {
        return u.code;
    }
    return 0;
}

// inputs kind 2: b.x.c0, b.x.c1, b.y.c0, b.y.c1, nega.(x,y,z), a.(x,y,z)
// inputs kind 3: b (4), bc (4), nega.(x,y,z), ac.(x,y,z)
inline int run_check_pairing(int kind, const std::vector<BN>& in, std::shared_ptr<Context> ctx) {
    bool bn = (kind == 2);
    EccContext e = bn ? EccContext::native(ctx, bn256_g1(), BN256_FQ(), true) : EccContext::general(ctx, bls12_381_g1(), BLS12_381_FQ());
    PairingContext pc(e, bn);
    auto g2_const = [&](const BN* v) {
        AssignedFq2 x = pc.fq2_assign_constant(HFq2{v[0], v[1]});
        AssignedFq2 y = pc.fq2_assign_constant(HFq2{v[2], v[3]});
        AssignedValue z = e.bc().assign_constant(n_from(0));
        return AssignedG2Affine{x, y, AssignedCondition(z)};
    };
    if (bn) {
        ORC_ASSERT(in.size() == 10);
        AssignedG2Affine b = g2_const(&in[0]);
        AssignedPoint neg_a = e.assign_point(HostPoint{in[4], in[5], !in[6].is_zero()});
        AssignedPoint a = e.assign_point(HostPoint{in[7], in[8], !in[9].is_zero()});
        pc.check_pairing({{&a, &b}, {&neg_a, &b}});
    } else {
        ORC_ASSERT(in.size() == 14);
        AssignedG2Affine b = g2_const(&in[0]);
        AssignedG2Affine bc = g2_const(&in[4]);
        AssignedPoint neg_a = e.assign_point(HostPoint{in[8], in[9], !in[10].is_zero()});
        AssignedPoint ac = e.assign_point(HostPoint{in[11], in[12], !in[13].is_zero()});
        pc.check_pairing({{&ac, &b}, {&neg_a, &bc}});
    }
    return 0;
}

// kind 5 (bn256) / 6 (bls12_381): pairing([(a, b)]) with b constant; the 12 Fq coefficients of the
// result are left in `result` (test-only KAT against independent plain-math pairings).
inline int run_single_pairing(int kind, const std::vector<BN>& in, std::shared_ptr<Context> ctx, std::vector<BN>* result) {
    bool bn = (kind == 5);
    EccContext e = bn ? EccContext::native(ctx, bn256_g1(), BN256_FQ(), true) : EccContext::general(ctx, bls12_381_g1(), BLS12_381_FQ());
    PairingContext pc(e, bn);
    ORC_ASSERT(in.size() == 7);
    AssignedFq2 x = pc.fq2_assign_constant(HFq2{in[0], in[1]});
    AssignedFq2 y = pc.fq2_assign_constant(HFq2{in[2], in[3]});
    AssignedValue z = e.bc().assign_constant(n_from(0));
    AssignedG2Affine b{x, y, AssignedCondition(z)};
    AssignedPoint a = e.assign_point(HostPoint{in[4], in[5], !in[6].is_zero()});
    AssignedFq12 r = pc.pairing({{&a, &b}});
    if (result) {
        const AssignedFq6* h[2] = {&r.c0, &r.c1};
        for (int i = 0; i < 2; i++)
            for (const AssignedFq2* q : {&h[i]->c0, &h[i]->c1, &h[i]->c2}) {
                result->push_back(e.base.get_w(q->first));
                result->push_back(e.base.get_w(q->second));
            }
    }
    return 0;
}

inline int run_circuit(int kind, const uint64_t* params, size_t n_params, const std::vector<BN>& in, std::shared_ptr<Context> ctx,
                       std::vector<BN>* result = nullptr) {
    switch (kind) {
        case 0:
        case 1:
        case 4: ORC_ASSERT(n_params >= 1); return run_msm(kind, params[0], in, ctx);
        case 2:
        case 3: return run_check_pairing(kind, in, ctx);
        case 5:
        case 6: return run_single_pairing(kind, in, ctx, result);
    }
    throw OraclePanic{"circuit kind not implemented"};
}

}  // namespace orc
