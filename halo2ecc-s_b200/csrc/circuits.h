// Shapes of the reference's own test circuits, traced through the chip mirror. Input layouts
// (logical 64-byte inputs; logical input i = input cells 2i, 2i+1):
//   MSM kinds 0/1/4 (src/tests/native_scalar_ecc_chip.rs:13-110, general_scalar_ecc_chip.rs:14-49):
//       [x_i, y_i, z_i] * n, [s_i] * n, r1.x, r1.y, r2.x, r2.y, expected.(x, y, z)     (z = identity flag)
//   bn256 check_pairing (native_scalar_pairing_chip.rs:67-97):  b.x.c0, b.x.c1, b.y.c0, b.y.c1, -a.(x,y,z), a.(x,y,z)
//   bls12_381 check_pairing (general_scalar_pairing_chip.rs:74-105): b (4), bc (4), -a.(x,y,z), ac.(x,y,z)
#pragma once
#include "pairing_tracer.h"

namespace h2e {

inline uint32_t cell_of(size_t logical) { return (uint32_t)(2 * logical); }

inline void build_msm(Context& ctx, int kind, size_t n) {
    bool bls = (kind == 4);
    EccContext e(&ctx, bls ? curve_bls12_381_g1() : curve_bn256_g1(), !bls, kind != 1);
    std::vector<AssignedPoint> points;
    for (size_t i = 0; i < n; i++) points.push_back(e.assign_point(PointInput{cell_of(3 * i), cell_of(3 * i + 1)}, cell_of(3 * i + 2)));
    std::vector<AssignedScalar> scalars;
    for (size_t i = 0; i < n; i++) {
        AssignedScalar s;
        if (bls)
            s.i = e.scalar.assign_w(cell_of(3 * n + i));
        else
            s.v = ctx.assign(cell_of(3 * n + i));
        scalars.push_back(s);
    }
    size_t t = 4 * n;
    AssignedPoint res = e.msm_unsafe(points, scalars, PointInput{cell_of(t), cell_of(t + 1)}, PointInput{cell_of(t + 2), cell_of(t + 3)});
    AssignedPoint res_expect = e.assign_point(PointInput{cell_of(t + 4), cell_of(t + 5)}, cell_of(t + 6));
    e.ecc_assert_equal(res, res_expect);
}

inline AssignedG2Affine g2_constant_input(Context& ctx, PairingOps& pc, size_t logical) {
    AssignedFq2 x = pc.fq2_assign_constant_input(cell_of(logical), cell_of(logical + 1));
    AssignedFq2 y = pc.fq2_assign_constant_input(cell_of(logical + 2), cell_of(logical + 3));
    AssignedValue z = ctx.assign_constant(Big(0));
    return AssignedG2Affine{x, y, AssignedCondition{z}};
}

inline void build_check_pairing(Context& ctx, int kind) {
    bool bn = (kind == 2);
    EccContext e(&ctx, bn ? curve_bn256_g1() : curve_bls12_381_g1(), bn, true);
    PairingOps pc(e, bn);
    if (bn) {
        AssignedG2Affine b = g2_constant_input(ctx, pc, 0);
        AssignedPoint neg_a = e.assign_point(PointInput{cell_of(4), cell_of(5)}, cell_of(6));
        AssignedPoint a = e.assign_point(PointInput{cell_of(7), cell_of(8)}, cell_of(9));
        pc.check_pairing({{&a, &b}, {&neg_a, &b}});
    } else {
        AssignedG2Affine b = g2_constant_input(ctx, pc, 0);
        AssignedG2Affine bc = g2_constant_input(ctx, pc, 4);
        AssignedPoint neg_a = e.assign_point(PointInput{cell_of(8), cell_of(9)}, cell_of(10));
        AssignedPoint ac = e.assign_point(PointInput{cell_of(11), cell_of(12)}, cell_of(13));
        pc.check_pairing({{&ac, &b}, {&neg_a, &bc}});
    }
}

inline void build_circuit(Context& ctx, int kind, const uint64_t* params, size_t n_params) {
    switch (kind) {
        case 0:
        case 1:
        case 4:
            if (n_params < 1 || params[0] == 0) throw std::runtime_error("MSM shape needs params[0] = number of points");
            build_msm(ctx, kind, params[0]);
            break;
        case 2:
        case 3: build_check_pairing(ctx, kind); break;
        default: throw std::runtime_error("circuit kind not implemented");
    }
    ctx.finish();
}

}  // namespace h2e
