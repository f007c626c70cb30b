"""The levelised schedule used by team mode: it must be a permutation of the program, respect every
data dependency (an instruction's operands are produced in strictly earlier levels), and executing
it in schedule order must produce exactly the records the program order produces."""
import random

import numpy as np
import pytest

import circuits_util as cu
import ecmath as em
import helpers


def _outs(prog, n_slots):
    out = prog[:, 4:8].copy().view(np.uint32).reshape(-1)
    order = np.argsort(out, kind="stable")
    ends = np.empty_like(out)
    ends[order] = np.append(out[order][1:], n_slots)
    return out, ends


def test_schedule_is_valid_and_equivalent(h2e, oracle):
    shape = h2e.Shape.build(0, [3])  # bn256 MSM with select chip, 3 points: 254 independent windows
    prog = shape.program()
    sprog, level_start = shape.schedule()
    OP_REDUCE, OP_INT_MUL, OP_DIV_CORE, OP_INT_ADD, OP_HEAD, OP_TAIL, OP_RHEAD, OP_RTAIL, OP_DINV, OP_DCORE_S = 8, 9, 10, 4, 29, 30, 31, 32, 33, 34
    OP_IS_INT_ZERO, OP_ZHEAD, OP_ZTAIL, OP_DHEAD, OP_DTAIL = 11, 35, 36, 37, 38
    pops = prog[:, 0:2].copy().view(np.uint16).reshape(-1)
    n_mul, n_red, n_div = int((pops == OP_INT_MUL).sum()), int((pops == OP_REDUCE).sum()), int((pops == OP_DIV_CORE).sum())
    n_iz = int((pops == OP_IS_INT_ZERO).sum())
    assert n_red > 0 and n_div > 0 and n_iz > 0
    n_ztail = int((sprog[:, 0:2].copy().view(np.uint16).reshape(-1) == OP_ZTAIL).sum())  # merged: up to 3 blocks per TAIL
    assert (n_iz + 2) // 3 <= n_ztail <= n_iz
    n_dinv = int((sprog[:, 0:2].copy().view(np.uint16).reshape(-1) == OP_DINV).sum())  # merged: up to 3 denominators per inversion
    assert (n_div + 2) // 3 <= n_dinv <= n_div  # (merging is a tuning option, off by default: H2E_DIVMERGE)
    assert sprog.shape[0] == prog.shape[0] + n_mul + n_red + n_div + n_dinv + n_ztail and level_start[0] == 0 and level_start[-1] == sprog.shape[0]
    # every int_mul appears as one HEAD and one TAIL, every reduce as one HEAD and one TAIL, with the same
    # operands; everything else is a permutation
    sops = sprog[:, 0:2].copy().view(np.uint16).reshape(-1)
    assert int((sops == OP_HEAD).sum()) == n_mul and int((sops == OP_TAIL).sum()) == n_mul and not (sops == OP_INT_MUL).any()
    assert int((sops == OP_RHEAD).sum()) == n_red and int((sops == OP_RTAIL).sum()) == n_red and not (sops == OP_REDUCE).any()
    # every div_core appears as the inversion (OP_DIV_INV, result in a scratch entry), the HEAD that turns it into the
    # quotient cells later ops read, and a deferred TAIL; every is_int_zero as a HEAD (condition cell only) and a TAIL
    assert int((sprog[sops == OP_DINV, 3] & 3).sum()) == n_div, "every int_div's inversion belongs to exactly one (merged) OP_DIV_INV"
    assert int((sops == OP_DHEAD).sum()) == n_div and int((sops == OP_DTAIL).sum()) == n_div
    assert not (sops == OP_DIV_CORE).any() and not (sops == OP_DCORE_S).any()
    assert int((sops == OP_ZHEAD).sum()) == n_iz and not (sops == OP_IS_INT_ZERO).any()
    assert int((sprog[sops == OP_ZTAIL, 3] & 3).sum()) == n_iz, "every is_int_zero block belongs to exactly one merged TAIL"
    flags = sprog[:, 3]
    assert (flags[(sops == OP_ZTAIL) | (sops == OP_DTAIL)] & 0x80).all(), "the TAILs must be deferred work"
    assert not (flags[(sops == OP_ZHEAD) | (sops == OP_DHEAD) | (sops == OP_DINV)] & 0x80).any()
    def rest(pr, ops, drop):  # flags bit 7 (set by the scheduler on deferred instructions) is not part of the program
        pr = pr.copy()
        pr[:, 3] &= 0x7F
        return sorted(bytes(x) for x, o in zip(pr, ops) if o not in drop)

    assert rest(sprog, sops, (OP_HEAD, OP_TAIL, OP_RHEAD, OP_RTAIL, OP_DINV, OP_ZHEAD, OP_ZTAIL, OP_DHEAD, OP_DTAIL)) == rest(
        prog, pops, (OP_INT_MUL, OP_REDUCE, OP_DIV_CORE, OP_IS_INT_ZERO)), "schedule is not a permutation of the program"
    assert sorted(bytes(x[2:]) for x, o in zip(sprog, sops) if o == OP_HEAD) == sorted(bytes(x[2:]) for x, o in zip(prog, pops) if o == OP_INT_MUL)
    n_levels = len(level_start) - 1
    assert n_levels < sprog.shape[0] * 0.6, "no parallelism found"
    # widths: the window phase must expose (at least) one op per window at some level
    assert (np.diff(level_start.astype(np.int64)) >= 254).any()
    # executing in schedule order gives the same bit-exact records as the oracle
    inputs = [cu.msm_inputs(em.BN256, 3, 99)]
    packed = h2e.pack_inputs(inputs)
    vals, status = helpers.run_emulated(shape, packed, program=sprog)
    assert status[0] == 0
    rec = oracle.run_circuit(0, [3], inputs[0])
    cells = helpers.compare_static(shape, rec)
    helpers.compare_instance(shape, cells, vals, 0, rec)
    # dependency check: producer level < consumer level for plain slot operands of int ops
    out, ends = _outs(sprog, shape.n_slots)
    level_of = np.zeros(sprog.shape[0], dtype=np.int64)
    for l in range(n_levels):
        level_of[level_start[l]:level_start[l + 1]] = l
    slot_level = np.full(shape.n_slots, -1, dtype=np.int64)
    ops = sops
    args = sprog[:, 8:64].copy().view(np.uint32).reshape(-1, 14)
    L = 3
    head_offsets = [6, 13, 18, 22, 23 + 6, 23 + 13, 23 + 18, 23 + 22]
    for i in range(sprog.shape[0]):
        if ops[i] != OP_HEAD:
            slot_level[out[i]:ends[i]] = np.maximum(slot_level[out[i]:ends[i]], -1)
    # blocks: non-HEAD instructions own [out, next out); HEAD cells are produced at the HEAD's level
    order = np.argsort(out, kind="stable")
    rhead_offsets = [6, 13, 18, 22, 23]
    for i in range(sprog.shape[0]):
        if ops[i] in (OP_HEAD, OP_RHEAD, OP_ZHEAD, OP_DHEAD, OP_DINV):
            continue
        slot_level[out[i]:ends[i]] = level_of[i]
    for i in np.nonzero(ops == OP_DHEAD)[0]:
        slot_level[out[i] + np.array(head_offsets[:4])] = level_of[i]
    for i in np.nonzero(ops == OP_ZHEAD)[0]:
        slot_level[args[i, 13]] = level_of[i]
    for i in np.nonzero(ops == OP_HEAD)[0]:
        slot_level[out[i] + np.array(head_offsets)] = level_of[i]
    for i in np.nonzero(ops == OP_RHEAD)[0]:
        slot_level[out[i] + np.array(rhead_offsets)] = level_of[i]
    for i in np.nonzero(ops == OP_DTAIL)[0][:1000]:
        assert (slot_level[args[i, :8]] < level_of[i]).all()
        assert (slot_level[out[i] + np.array(head_offsets[:4])] < level_of[i]).all()
    for i in np.nonzero((ops == OP_ZHEAD) | (ops == OP_ZTAIL))[0][:1000]:
        n_blocks = 1 if ops[i] == OP_ZHEAD else int(sprog[i, 3] & 3)
        assert (slot_level[args[i, :4 * n_blocks]] < level_of[i]).all()
    for i in np.nonzero(ops == OP_RTAIL)[0][:1000]:
        assert (slot_level[args[i, :4]] < level_of[i]).all()
        assert (slot_level[out[i] + np.array(rhead_offsets)] < level_of[i]).all()
    for i in np.nonzero((ops == OP_HEAD) | (ops == OP_INT_ADD))[0][:3000]:
        idx = list(range(6)) if ops[i] == OP_INT_ADD else [0, 1, 2, 4, 5, 6]
        assert (slot_level[args[i, idx]] < level_of[i]).all()
    for i in np.nonzero(ops == OP_TAIL)[0][:1000]:
        assert (slot_level[args[i, :8]] < level_of[i]).all()
        assert (slot_level[out[i] + np.array(head_offsets)] < level_of[i]).all()


def test_pairing_schedule_stats(h2e):
    shape = h2e.Shape.build(2, [])
    _, level_start = shape.schedule()
    n_levels = len(level_start) - 1
    assert shape.n_instr == 174806 and 8000 < n_levels < 9100


@pytest.mark.parametrize("ctas", [1, 3, 37])
def test_team_streams_execute_to_the_same_records(h2e, oracle, ctas):
    """Team mode's dataflow streams (per-warp instruction streams + (warp, count) dependencies): a host
    model of the execution must not deadlock, must start every instruction exactly once, and the
    order in which it starts them must reproduce the oracle's records bit-exactly."""
    shape = h2e.Shape.build(0, [3])
    sprog, _ = shape.schedule()
    order, est = shape.team_order(ctas)
    assert est > 0 and order.shape == sprog.shape
    assert sorted(bytes(x) for x in order) == sorted(bytes(x) for x in sprog)
    inputs = [cu.msm_inputs(em.BN256, 3, 4242)]
    vals, status = helpers.run_emulated(shape, h2e.pack_inputs(inputs), program=order)
    assert status[0] == 0
    rec = oracle.run_circuit(0, [3], inputs[0])
    cells = helpers.compare_static(shape, rec)
    helpers.compare_instance(shape, cells, vals, 0, rec)


def test_team_streams_with_dedicated_inversion_ctas(h2e, oracle, monkeypatch):
    """H2E_INV_CTAS=1 (tuning option, off by default): OP_DIV_INV only on the first critical CTAs, is_int_zero TAILs
    only on the first tail CTAs. The streams must still run to completion in the host model and reproduce the oracle."""
    monkeypatch.setenv("H2E_INV_CTAS", "1")
    shape = h2e.Shape.build(0, [3])
    sprog, _ = shape.schedule()
    order, est = shape.team_order(37)
    assert est > 0 and sorted(bytes(x) for x in order) == sorted(bytes(x) for x in sprog)
    inputs = [cu.msm_inputs(em.BN256, 3, 777)]
    vals, status = helpers.run_emulated(shape, h2e.pack_inputs(inputs), program=order)
    assert status[0] == 0
    rec = oracle.run_circuit(0, [3], inputs[0])
    cells = helpers.compare_static(shape, rec)
    helpers.compare_instance(shape, cells, vals, 0, rec)


def test_team_streams_scale_with_ctas(h2e):
    """More CTAs per tile must shorten the modelled makespan (the serial tail of the MSM bounds the gain)."""
    shape = h2e.Shape.build(0, [3])
    _, e1 = shape.team_order(1)
    _, e8 = shape.team_order(8)
    assert e8 < e1
