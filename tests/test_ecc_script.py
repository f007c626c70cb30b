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
