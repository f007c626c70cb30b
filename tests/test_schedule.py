"""The levelised schedule used by team mode: it must be a permutation of the program, respect every
data dependency (an instruction's operands are produced in strictly earlier levels), and executing
it in schedule order must produce exactly the records the program order produces."""
import random

import numpy as np

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
    assert sprog.shape == prog.shape and level_start[0] == 0 and level_start[-1] == shape.n_instr
    assert sorted(map(bytes, sprog)) == sorted(map(bytes, prog)), "schedule is not a permutation of the program"
    n_levels = len(level_start) - 1
    assert n_levels < shape.n_instr * 0.6, "no parallelism found"
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
    level_of = np.zeros(shape.n_instr, dtype=np.int64)
    for l in range(n_levels):
        level_of[level_start[l]:level_start[l + 1]] = l
    slot_level = np.full(shape.n_slots, -1, dtype=np.int64)
    for i in range(shape.n_instr):
        slot_level[out[i]:ends[i]] = level_of[i]
    ops = sprog[:, 0:2].copy().view(np.uint16).reshape(-1)
    args = sprog[:, 8:64].copy().view(np.uint32).reshape(-1, 14)
    OP_INT_MUL, OP_INT_ADD = 9, 4
    for i in np.nonzero((ops == OP_INT_MUL) | (ops == OP_INT_ADD))[0][:2000]:
        n_ops = 8 if ops[i] == OP_INT_MUL else 6
        assert (slot_level[args[i, :n_ops]] < level_of[i]).all()


def test_pairing_schedule_stats(h2e):
    shape = h2e.Shape.build(2, [])
    _, level_start = shape.schedule()
    n_levels = len(level_start) - 1
    assert shape.n_instr == 174806 and 8000 < n_levels < 9000
