"""EccChipBaseOps / EccChipScalarOps through the generic op-script entry of the C ABI
(h2e_shape_from_script): the safe-point API of src/circuit/ecc_chip.rs (assign_point,
to_point_with_curvature, ecc_add, ecc_double, ecc_neg, ecc_reduce, ecc_assert_equal, ecc_encode) and
msm_unsafe with explicit blinding points, on bn256 G1 and bls12_381 G1. The product's records must be
bit-exact with the oracle's, and the in-circuit results must equal independent plain math."""
import pytest

import ecmath as em
import helpers


def _pt(P):
    return [0, 0, 1] if P is None else [P[0], P[1], 0]


def _group_law_script(h2e):
    sb = h2e.ScriptBuilder()
    a, b, want_sum, want_dbl, ident = (sb.assign_point(3 * i) for i in range(5))
    pa = sb.to_point_with_curvature(a)
    s = sb.ecc_add(pa, b)                      # a + b
    d = sb.ecc_double(pa)                      # 2a
    sb.ecc_assert_equal(s, want_sum)
    sb.ecc_assert_equal(d, want_dbl)
    z = sb.ecc_add(pa, sb.ecc_neg(a))          # a + (-a) = identity
    sb.ecc_assert_equal(z, ident)
    pi = sb.to_point_with_curvature(ident)
    sb.ecc_assert_equal(sb.ecc_add(pi, b), b)  # identity + b = b
    sb.ecc_assert_equal(sb.ecc_add(pa, a), d)  # a + a takes the tangent branch
    enc = sb.ecc_encode(sb.ecc_reduce(s))
    sb.assert_equal(enc[0], enc[0])
    return sb


def _group_law_inputs(C, k1, k2):
    A, B = C.mul(C.g1, k1, 1), C.mul(C.g1, k2, 1)
    return _pt(A) + _pt(B) + _pt(C.add(A, B, 1)) + _pt(C.add(A, A, 1)) + _pt(None)


@pytest.mark.parametrize("field,curve", [(0, "BN256"), (1, "BLS12_381")])
def test_safe_point_api_script(h2e, oracle, field, curve):
    C = getattr(em, curve)
    sb = _group_law_script(h2e)
    inputs = [_group_law_inputs(C, 5 + i, 11 + 7 * i) for i in range(2)]
    helpers.check_script(h2e, oracle, field, sb.words, inputs)


def _msm_script(h2e, n):
    sb = h2e.ScriptBuilder()
    pts = [sb.assign_point(3 * i) for i in range(n)]
    scs = [sb.assign(3 * n + i) for i in range(n)]
    res = sb.msm(pts, scs, 4 * n, 4 * n + 2)
    sb.ecc_assert_equal(res, sb.assign_point(4 * n + 4))
    return sb


def _msm_inputs(n, seed):
    C = em.BN256
    g = em.scalar_stream(seed, C.r)
    pts, scs, acc = [], [], None
    for _ in range(n):
        P, s = C.mul(C.g1, next(g) or 1, 1), next(g)
        pts += _pt(P)
        scs.append(s)
        acc = C.add(acc, C.mul(P, s, 1), 1)
    r1, r2 = C.mul(C.g1, next(g) or 1, 1), C.mul(C.g1, next(g) or 1, 1)
    return pts + scs + [r1[0], r1[1], r2[0], r2[1]] + _pt(acc)


def test_msm_script_matches_builtin_shape_semantics(h2e, oracle):
    """msm_unsafe through the script (2 points) -- same chip calls as the built-in MSM shape."""
    sb = _msm_script(h2e, 2)
    helpers.check_script(h2e, oracle, 0, sb.words, [_msm_inputs(2, 77)])


def test_wrong_sum_is_flagged(h2e, oracle):
    sb = _group_law_script(h2e)
    good = _group_law_inputs(em.BN256, 5, 11)
    bad = list(good)
    bad[6], bad[7] = good[9], good[10]  # claim a + b = 2a
    shape = h2e.Shape.from_script(0, sb.words)
    _, status = helpers.run_emulated(shape, h2e.pack_inputs([good, bad]))
    assert status[0] == 0 and status[1] & h2e.ST_ASSERT_VALUE
    assert oracle.run_script(0, sb.words, bad).status != 0


@pytest.mark.gpu
def test_safe_point_api_script_gpu(h2e, oracle):
    sb = _group_law_script(h2e)
    inputs = [_group_law_inputs(em.BN256, 5 + i, 11 + 7 * i) for i in range(3)]
    helpers.check_script(h2e, oracle, 0, sb.words, inputs, runner=helpers.run_gpu)


@pytest.mark.gpu
def test_msm_script_gpu(h2e, oracle):
    sb = _msm_script(h2e, 2)
    helpers.check_script(h2e, oracle, 0, sb.words, [_msm_inputs(2, 77), _msm_inputs(2, 78)], runner=helpers.run_gpu)


def _pairing_script(h2e):
    """The reference's bn256 check_pairing test body (native_scalar_pairing_chip.rs:67-97) as a script."""
    sb = h2e.ScriptBuilder()
    b = sb.assign_g2_constant(0)
    neg_a = sb.assign_point(4)
    a = sb.assign_point(7)
    sb.check_pairing([(a, b), (neg_a, b)])
    return sb


def test_check_pairing_script_equals_builtin_shape(h2e, oracle):
    """PairingChipOps through the script entry: same records as the built-in bn256 check_pairing shape."""
    import numpy as np

    sb = _pairing_script(h2e)
    s1 = h2e.Shape.from_script(0, sb.words)
    s2 = h2e.Shape.build(h2e.CIRCUIT_PAIRING_BN256)
    assert (s1.n_slots, s1.n_instr, s1.n_perms, s1.base_height, s1.range_height) == (s2.n_slots, s2.n_instr, s2.n_perms, s2.base_height, s2.range_height)
    assert np.array_equal(s1.program(), s2.program()) and np.array_equal(s1.perms(), s2.perms()) and np.array_equal(s1.fixed(), s2.fixed())
    import circuits_util as cu

    inp = cu.bn_check_pairing_inputs(1000003, 2000003)
    r1, r2 = oracle.run_script(0, sb.words, inp), oracle.run_circuit(2, [], inp)
    assert r1.status == 0 and r2.status == 0 and r1.n_adv == r2.n_adv == s1.n_slots
    for reg in range(3):
        assert np.array_equal(r1.adv[reg], r2.adv[reg])


@pytest.mark.gpu
def test_check_pairing_script_gpu(h2e, oracle):
    import circuits_util as cu

    sb = _pairing_script(h2e)
    inputs = [cu.bn_check_pairing_inputs(1000003 + i, 2000003 + i) for i in range(2)]
    shape = h2e.Shape.from_script(0, sb.words)
    vals, status = helpers.run_gpu(shape, h2e.pack_inputs(inputs))
    assert (status == 0).all()
    rec = oracle.run_script(0, sb.words, inputs[1])
    cells = helpers.compare_static(shape, rec)
    helpers.compare_instance(shape, cells, vals, 1, rec)
