"""Fq2 / Fq6 / Fq12ChipOps (src/circuit/fq12.rs:10-459), PairingChipOps::pairing / multi_miller_loop /
final_exponentiation (src/circuit/pairing_chip.rs:13-176), ecc_mul, ecc_reduce_with_curvature and the general-scalar
msm_unsafe (src/circuit/general_scalar_ecc_chip.rs:96-147) through the generic op-script entry of the C ABI: every
trait method is reachable without a built-in shape, and its records are bit-exact with the oracle's."""
import random

import pytest

import circuits_util as cu
import ecmath as em
import helpers


def _tower_script(h2e, bn):
    sb = h2e.ScriptBuilder()
    ints = [sb.assign_w(i) for i in range(12)]
    f2 = [sb.fq2_from_ints(ints[2 * i], ints[2 * i + 1]) for i in range(6)]
    # Fq2ChipOps
    m = sb.fq2_mul(f2[0], f2[1])
    s = sb.fq2_sub(sb.fq2_add(m, f2[2]), f2[3])
    t = sb.fq2_mul_by_nonresidue(sb.fq2_double(sb.fq2_neg(s)))
    inv = sb.fq2_unsafe_invert(f2[4])
    sb.fq2_assert_equal(sb.fq2_reduce(sb.fq2_mul(inv, f2[4])), sb.fq2_reduce(sb.fq2_mul(f2[4], inv)))
    fr = sb.fq2_frobenius_map(t, 1)
    c0, c1 = sb.fq2_parts(fr)
    sb.assert_int_equal(c0, c0)
    # Fq6ChipOps
    a6 = sb.fq6_from_fq2s(f2[0], f2[1], f2[2])
    b6 = sb.fq6_from_fq2s(f2[3], f2[4], f2[5])
    p6 = sb.fq6_mul(a6, b6)
    q6 = sb.fq6_sub(sb.fq6_add(p6, a6), sb.fq6_neg(b6))
    r6 = sb.fq6_mul_by_01(sb.fq6_mul_by_1(q6, f2[1]), f2[2], f2[3])
    i6 = sb.fq6_unsafe_invert(b6)
    sb.fq6_assert_equal(sb.fq6_mul(i6, b6), sb.fq6_mul(b6, i6))
    k6 = sb.fq6_frobenius_map(r6, 2)
    # Fq12ChipOps
    x = sb.fq12_from_fq6s(p6, k6)
    y = sb.fq12_from_fq6s(b6, a6)
    z = sb.fq12_mul(x, y)
    z = sb.fq12_mul_by_034(z, f2[0], f2[1], f2[2]) if bn else sb.fq12_mul_by_014(z, f2[0], f2[1], f2[2])
    z = sb.fq12_cyclotomic_square(z)
    zi = sb.fq12_unsafe_invert(y)
    sb.fq12_assert_eq(sb.fq12_mul(zi, y), sb.fq12_mul(y, zi))
    w = sb.fq12_frobenius_map(z, 1)
    h0, h1 = sb.fq12_parts(w)
    sb.fq6_assert_equal(h0, h0)
    return sb


@pytest.mark.parametrize("field", [0, 1])
def test_tower_ops_script(h2e, oracle, field):
    rng = random.Random(100 + field)
    p = oracle.FIELD_MODULUS[field]
    sb = _tower_script(h2e, field == 0)
    inputs = [[rng.randrange(1, p) for _ in range(12)] for _ in range(2)]
    helpers.check_script(h2e, oracle, field, sb.words, inputs)


def _pt(P):
    return [0, 0, 1] if P is None else [P[0], P[1], 0]


def test_ecc_mul_and_reduce_with_curvature_script(h2e, oracle):
    C = em.BN256
    sb = h2e.ScriptBuilder()
    a = sb.assign_point(0)
    k = sb.assign(3)
    want = sb.assign_point(8)
    res = sb.ecc_mul(a, k, 4, 6)  # blinding points r1, r2 at inputs 4..7
    sb.ecc_assert_equal(res, want)
    pwc = sb.ecc_reduce_with_curvature(res)
    sb.ecc_assert_equal(sb.ecc_double(pwc), sb.assign_point(11))
    rows = []
    for i in range(2):
        A, kk = C.mul(C.g1, 12345 + i, 1), 987654321987654321 + i
        r1, r2 = C.mul(C.g1, 777 + i, 1), C.mul(C.g1, 999 + i, 1)
        R = C.mul(A, kk, 1)
        rows.append(_pt(A) + [kk, r1[0], r1[1], r2[0], r2[1]] + _pt(R) + _pt(C.add(R, R, 1)))
    helpers.check_script(h2e, oracle, 0, sb.words, rows)


def _msm_general_script(h2e, n):
    sb = h2e.ScriptBuilder()
    pts = [sb.assign_point(3 * i) for i in range(n)]
    scs = [sb.assign_scalar_w(3 * n + i) for i in range(n)]
    res = sb.msm_general(pts, scs, 4 * n, 4 * n + 2)
    sb.ecc_assert_equal(res, sb.assign_point(4 * n + 4))
    return sb


def test_general_scalar_msm_script_is_the_builtin_shape(h2e, oracle):
    """msm_unsafe of the general-scalar context (bls12_381 G1, scalars as integers of its scalar field) from a script: the same
    chip calls as the reference's test shape (src/tests/general_scalar_ecc_chip.rs:14-49), so the traced shape must be
    identical to the built-in one, and the records bit-exact with the oracle's."""
    import numpy as np

    n = 1
    sb = _msm_general_script(h2e, n)
    rows = [cu.msm_inputs(em.BLS12_381, n, 31 + i) for i in range(1)]
    shape = helpers.check_script(h2e, oracle, 1, sb.words, rows)
    builtin = h2e.Shape.build(h2e.CIRCUIT_MSM_BLS12_381, [n])
    assert (shape.n_slots, shape.n_perms, shape.base_offset, shape.range_offset, shape.select_offset) == (
        builtin.n_slots, builtin.n_perms, builtin.base_offset, builtin.range_offset, builtin.select_offset)
    assert np.array_equal(shape.program(), builtin.program())


def test_script_rejects_wrong_arity_and_truncation(h2e):
    """run_script validates every record against a per-opcode arity table before it reads an argument."""
    sb = h2e.ScriptBuilder()
    a = sb.assign_w(0)
    words = list(sb.words) + [h2e.OPS["INT_ADD"], 1, a]  # int_add with one argument
    with pytest.raises(h2e.H2EError, match="takes 2 arguments"):
        h2e.Shape.from_script(0, words)
    with pytest.raises(h2e.H2EError, match="truncated"):
        h2e.Shape.from_script(0, list(sb.words) + [h2e.OPS["INT_ADD"], 2, a])  # record runs past the end
    with pytest.raises(h2e.H2EError, match="count argument"):
        h2e.Shape.from_script(0, list(sb.words) + [h2e.OPS["MSM"], 0])
    with pytest.raises(h2e.H2EError, match="unknown script op"):
        h2e.Shape.from_script(0, [9999, 0])


@pytest.mark.gpu
@pytest.mark.parametrize("field", [0, 1])
def test_tower_ops_script_gpu(h2e, oracle, field):
    rng = random.Random(200 + field)
    p = oracle.FIELD_MODULUS[field]
    sb = _tower_script(h2e, field == 0)
    inputs = [[rng.randrange(1, p) for _ in range(12)] for _ in range(34)]
    shape = h2e.Shape.from_script(field, sb.words)
    vals, status = helpers.run_gpu(shape, h2e.pack_inputs(inputs))
    assert (status == 0).all()
    cells = None
    for i in (0, 33):
        rec = oracle.run_script(field, sb.words, inputs[i])
        assert rec.status == 0 and rec.gate_ok, (rec.error, rec.gate_msg)
        if cells is None:
            cells = helpers.compare_static(shape, rec)
        helpers.compare_instance(shape, cells, vals, i, rec)


@pytest.mark.gpu
def test_pairing_trait_methods_script_gpu(h2e, oracle):
    """pairing() + fq12_assert_one is check_pairing; multi_miller_loop + final_exponentiation + fq12_assert_one as well
    (pairing_chip.rs:157-176): both scripts must trace the built-in shape of the reference's bn256 pairing test, and the GPU
    records must be the oracle's."""
    import numpy as np

    def script(split):
        sb = h2e.ScriptBuilder()
        b = sb.assign_g2_constant(0)
        neg_a = sb.assign_point(4)
        a = sb.assign_point(7)
        terms = [(a, b), (neg_a, b)]
        f = sb.final_exponentiation(sb.multi_miller_loop(terms)) if split else sb.pairing(terms)
        sb.fq12_assert_one(f)
        return sb

    builtin = h2e.Shape.build(h2e.CIRCUIT_PAIRING_BN256, [])
    rows = [cu.bn_check_pairing_inputs(424243 + i, 171719 + i) for i in range(2)]
    for split in (False, True):
        sb = script(split)
        shape = h2e.Shape.from_script(0, sb.words)
        assert np.array_equal(shape.program(), builtin.program())
        assert (shape.n_slots, shape.n_perms) == (builtin.n_slots, builtin.n_perms)
    vals, status = helpers.run_gpu(shape, h2e.pack_inputs(rows))
    assert (status == 0).all()
    rec = oracle.run_script(0, sb.words, rows[1])
    assert rec.status == 0, rec.error
    cells = helpers.compare_static(shape, rec)
    helpers.compare_instance(shape, cells, vals, 1, rec)


@pytest.mark.gpu
def test_ecc_mul_and_general_msm_script_gpu(h2e, oracle):
    sb = _msm_general_script(h2e, 2)
    rows = [cu.msm_inputs(em.BLS12_381, 2, 51 + i) for i in range(3)]
    helpers.check_script(h2e, oracle, 1, sb.words, rows, runner=helpers.run_gpu)
