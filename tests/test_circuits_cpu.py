"""Whole-circuit parity on the CPU: the product's traced shapes (static half) and macro-op code
(run on the host emulator, test infrastructure) against the oracle's records, for the shapes of
the reference's MSM and pairing tests."""
import time

import pytest

import circuits_util as cu
import ecmath as em
import helpers


def check_circuit(h2e, oracle, kind, params, inputs_per_instance, runner=helpers.run_emulated):
    shape = h2e.Shape.build(kind, params)
    packed = h2e.pack_inputs(inputs_per_instance)
    vals, status = runner(shape, packed)
    cells = None
    for i, inp in enumerate(inputs_per_instance):
        rec = oracle.run_circuit(kind, params, inp)
        assert rec.status == 0, rec.error
        assert status[i] == 0, f"instance {i}: status {status[i]}"
        if cells is None:
            cells = helpers.compare_static(shape, rec)
            assert rec.n_adv == shape.n_slots
        helpers.compare_instance(shape, cells, vals, i, rec)
    return shape


@pytest.mark.parametrize("kind,n", [(0, 1), (0, 7), (1, 3), (4, 2)])
def test_msm_shapes(h2e, oracle, kind, n):
    C = em.BLS12_381 if kind == 4 else em.BN256
    inputs = [cu.msm_inputs(C, n, 20240601 + i) for i in range(2)]
    check_circuit(h2e, oracle, kind, [n], inputs)


def test_bn256_check_pairing_shape(h2e, oracle):
    inputs = [cu.bn_check_pairing_inputs(123456789 + i, 987654321 + 7 * i) for i in range(1)]
    shape = check_circuit(h2e, oracle, 2, [], inputs)
    assert (shape.base_offset, shape.range_offset) == (1049946, 1103352)


@pytest.mark.slow
def test_bls12_381_check_pairing_shape(h2e, oracle):
    inputs = [cu.bls_check_pairing_inputs(777777777, 5555555, 1234567890123456789012345)]
    shape = check_circuit(h2e, oracle, 3, [], inputs)
    assert (shape.base_offset, shape.range_offset) == (1300575, 1433618)


def test_unsafe_error_is_reported_per_instance(h2e, oracle):
    """UnsafeError::AddSameOrNegPoint (ecc_chip.rs:23-34, 840-858): when the blinding point equals an
    input point the reference returns Err and its tests retry with fresh randomness
    (native_scalar_ecc_chip.rs:52-57). Here the instance's status carries the code; the good instance
    beside it is unaffected."""
    import circuits_util as cu
    import ecmath as em

    good = cu.msm_inputs(em.BN256, 1, 5)
    bad = list(good)
    bad[6], bad[7] = good[0], good[1]  # r2 := P_0 -> the first candidate addition is P + P
    assert oracle.run_circuit(0, [1], bad).status == 1
    shape = h2e.Shape.build(0, [1])
    vals, status = helpers.run_emulated(shape, h2e.pack_inputs([good, bad]))
    assert status[0] == 0 and status[1] & h2e.ST_ADD_SAME_OR_NEG
    rec = oracle.run_circuit(0, [1], good)
    cells = helpers.compare_static(shape, rec)
    helpers.compare_instance(shape, cells, vals, 0, rec)
