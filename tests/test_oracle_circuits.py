"""Pins the oracle (the CPU restatement of the reference) on the ECC / MSM / pairing path:
  * the Frobenius / twist constants it derives equal the tables embedded in the reference
    (golden fixture extracted by tests/golden/gen_reference_constants.py),
  * row / permutation counts equal SURVEY Appendix B (an independent structural replay),
  * every gate, lookup and permutation of the reference's `configure` holds on its records,
  * circuit self-checks (MSM == expected, pairing product == 1) pass against values computed by
    independent plain math (tests/ecmath.py), and the in-circuit pairing value equals the plain
    optimal-ate pairing."""
import json
import os

import pytest

import circuits_util as cu
import ecmath as em

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_pairing_constants.json")))


def test_pairing_constants_match_reference_tables(oracle):
    k = oracle.pairing_constants()
    g = GOLD["bn256"]
    pairs = lambda flat: [(flat[2 * i], flat[2 * i + 1]) for i in range(len(flat) // 2)]
    assert [c[0] for c in k["fq2_c1"]] == g["FROBENIUS_COEFF_FQ2_C1"]
    assert k["fq6_c1"] == pairs(g["FROBENIUS_COEFF_FQ6_C1"])
    assert k["fq6_c2"] == pairs(g["FROBENIUS_COEFF_FQ6_C2"])
    assert k["fq12_c1"] == pairs(g["FROBENIUS_COEFF_FQ12_C1"])
    assert k["xi_to_q_minus_1_over_2"] == tuple(g["XI_TO_Q_MINUS_1_OVER_2"])
    assert sum(d << i for i, d in enumerate(g["SIX_U_PLUS_2_NAF"])) == 6 * g["BN_X"] + 2 == 6 * em.BN256.x + 2
    b = GOLD["bls12_381"]
    assert k["bls_fq6_c1"] == (0, b["FQ6_C1_c1"])
    assert k["bls_fq6_c2"] == (b["FQ6_C2_c0"], 0)
    assert k["bls_fq12_c1"] == tuple(b["FQ12_C1"])
    assert b["BLS_X"] == em.BLS12_381.x


def test_reference_constants_fixture_is_current():
    """When the reference is mounted (build container), the committed fixture must equal it."""
    if not os.path.exists("/root/reference/src/circuit/bn256_constants.rs"):
        pytest.skip("reference not mounted")
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen", os.path.join(HERE, "golden", "gen_reference_constants.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert {"bn256": gen.parse_bn(), "bls12_381": gen.parse_bls()} == GOLD


def test_generators_and_subgroups():
    for C in (em.BN256, em.BLS12_381):
        assert C.on_curve(C.g1, 1) and C.on_curve(C.g2, 2)
        assert C.mul(C.g1, C.r, 1) is None and C.mul(C.g2, C.r, 2) is None


# (base rows, range rows, select rows, permutations)
@pytest.mark.parametrize("kind,n,expect", [(0, 1, (106792, 123876, 2048, 266928)), (0, 7, None), (1, 3, None), (4, 2, None)])
def test_msm_oracle(oracle, kind, n, expect):
    C = em.BLS12_381 if kind == 4 else em.BN256
    rec = oracle.run_circuit(kind, [n], cu.msm_inputs(C, n, 20240601))
    assert rec.status == 0, rec.error
    assert rec.gate_ok, rec.gate_msg
    if expect:
        assert (rec.base_offset, rec.range_offset, rec.select_offset, len(rec.perms)) == expect


def test_msm_wrong_expected_result_is_rejected(oracle):
    inp = cu.msm_inputs(em.BN256, 2, 5)
    inp[-3] = (inp[-3] + 1) % em.BN256.p  # not even on the curve any more
    rec = oracle.run_circuit(0, [2], inp)
    assert rec.status != 0


def test_bn256_pairing_value_matches_plain_math(oracle):
    """first block of src/tests/native_scalar_pairing_chip.rs:20-65: in-circuit pairing(a,b) vs an
    independent optimal-ate pairing (affine lines, plain final power)."""
    C = em.BN256
    a, b = C.mul(C.g1, 1234567891234572, 1), C.mul(C.g2, 98765432198770, 2)
    rec, res = oracle.run_circuit_result(5, [], cu.g2_flat(b) + list(a) + [0])
    assert rec.status == 0 and rec.gate_ok, (rec.error, rec.gate_msg)
    assert res == [c for q in C.to_tower(C.pairing(a, b)) for c in q]


@pytest.mark.slow
def test_bls12_381_pairing_value_matches_plain_math(oracle):
    """The bls12_381 final exponentiation of the reference (ported from zkcrypto/bls12_381) yields
    the cube of the plain reduced pairing; the oracle must reproduce exactly that."""
    C = em.BLS12_381
    a, b = C.mul(C.g1, 1234567891234573, 1), C.mul(C.g2, 98765432198771, 2)
    rec, res = oracle.run_circuit_result(6, [], cu.g2_flat(b) + list(a) + [0])
    assert rec.status == 0 and rec.gate_ok, (rec.error, rec.gate_msg)
    assert res == [c for q in C.to_tower(C.pairing(a, b, 3)) for c in q]


def test_bn256_check_pairing_oracle(oracle):
    """second block of src/tests/native_scalar_pairing_chip.rs:67-97. Phase row counts match
    SURVEY (prepare_g2 187992, Miller 408752, final exp 452897, assert-one 124); the harness is
    181 rows because the test assigns the G2 constant once and uses it for both pairs."""
    rec = oracle.run_circuit(2, [], cu.bn_check_pairing_inputs(123456789123456789, 987654321987654321))
    assert rec.status == 0, rec.error
    assert rec.gate_ok, rec.gate_msg
    assert (rec.base_offset, rec.range_offset, rec.select_offset, len(rec.perms)) == (187992 + 408752 + 452897 + 124 + 181, 1103352, 0, 2690284)


@pytest.mark.slow
def test_bls12_381_check_pairing_oracle(oracle):
    """second block of src/tests/general_scalar_pairing_chip.rs:74-105; counts = SURVEY Appendix B."""
    rec = oracle.run_circuit(3, [], cu.bls_check_pairing_inputs(777777777, 5555555, 123456789012345678901234567890))
    assert rec.status == 0, rec.error
    assert rec.gate_ok, rec.gate_msg
    assert (rec.base_offset, rec.range_offset, rec.select_offset, len(rec.perms)) == (1300575, 1433618, 0, 3524865)


def test_check_pairing_rejects_unrelated_points(oracle):
    C = em.BN256
    inp = cu.bn_check_pairing_inputs(5, 7)
    a2 = C.mul(C.g1, 6, 1)
    inp[7], inp[8] = a2  # a != -(-a)
    rec = oracle.run_circuit(2, [], inp)
    assert rec.status != 0
