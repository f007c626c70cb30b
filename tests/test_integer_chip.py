"""Integer-chip parity on the CPU: oracle vs the survey's structural table, and the product's
tracer + macro-op code (run on the host emulator) vs the oracle. Shapes follow the reference's own
tests (src/tests/integer_chip.rs:11-99)."""
import random

import numpy as np
import pytest

import helpers

FIELDS = [0, 1, 2]
# (base rows, range rows, perms, advice cells) per op, by limb count
EXPECT = {
    3: dict(assign_w=(1, 8, 3, 23), int_add=(4, 0, 9, 13), reduce=(3, 12, 10, 40), int_mul=(17, 28, 47, 125),
            int_div=(33, 28, 74, 168), assert_int_equal=(9, 12, 23, 58)),
    4: dict(assign_w=(1, 11, 4, 31), int_add=(5, 0, 12, 17), reduce=(4, 18, 16, 60), int_mul=(30, 42, 83, 204),
            int_div=(51, 42, 119, 261), assert_int_equal=(11, 18, 33, 83)),
}


def _delta(oracle, field, words_before, words_after, inputs):
    r0 = oracle.run_script(field, words_before, inputs)
    r1 = oracle.run_script(field, words_after, inputs)
    assert r1.gate_ok, r1.gate_msg
    return (r1.base_offset - r0.base_offset, r1.range_offset - r0.range_offset, len(r1.perms) - len(r0.perms), r1.n_adv - r0.n_adv)


@pytest.mark.parametrize("field", FIELDS)
def test_oracle_structure_matches_survey(oracle, field):
    p = oracle.FIELD_MODULUS[field]
    L = 4 if field == 1 else 3
    rng = random.Random(field)
    a, b = rng.randrange(p), rng.randrange(1, p)
    sb = oracle.ScriptBuilder()
    ia, ib = sb.assign_w(0), sb.assign_w(1)
    base = list(sb.words)
    exp = EXPECT[L]

    def one(fn):
        s2 = oracle.ScriptBuilder()
        s2.words, s2.n_int, s2.n_val = list(base), 2, 0
        fn(s2)
        return _delta(oracle, field, base, s2.words, [a, b])

    r0 = oracle.run_script(field, base, [a, b])
    # range_height is one past the rows used (context.rs:716-720 with offset + lines)
    assert (r0.base_height, r0.range_offset, len(r0.perms), r0.n_adv) == tuple(2 * x for x in exp["assign_w"])
    assert one(lambda s: s.int_add(ia, ib)) == exp["int_add"]
    assert one(lambda s: s.int_sub(ia, ib)) == exp["int_add"]
    assert one(lambda s: s.int_mul(ia, ib)) == exp["int_mul"]
    # int_div on reduced operands, SURVEY counts exclude the (no-op) reduces
    assert one(lambda s: s.int_div(ia, ib)) == exp["int_div"]
    # reduce of an overflowed integer (times>1)
    s3 = oracle.ScriptBuilder()
    x = s3.load_int(3, 0)
    before = list(s3.words)
    s3.reduce(x)
    lead = oracle.FIELD_MODULUS[field].bit_length() % 108
    limbs = [rng.randrange(3 << 108) for _ in range(L - 1)] + [rng.randrange(3 << lead)]
    assert _delta(oracle, field, before, s3.words, limbs) == exp["reduce"]


@pytest.mark.parametrize("field", FIELDS)
def test_reference_integer_chip_test_shape(h2e, oracle, field):
    """src/tests/integer_chip.rs:11-55: add/sub/mul/div + div by zero flag, 40 random instances."""
    p = oracle.FIELD_MODULUS[field]
    rng = random.Random(100 + field)
    sb = h2e.ScriptBuilder()
    a, b = sb.assign_w(0), sb.assign_w(1)
    c1 = sb.assign_w(2)
    sb.assert_int_equal(c1, sb.int_add(a, b))
    d1 = sb.assign_w(3)
    sb.assert_int_equal(d1, sb.int_sub(a, b))
    e1 = sb.assign_w(4)
    sb.assert_int_equal(e1, sb.int_mul(a, b))
    f1 = sb.assign_w(5)
    sb.assert_int_equal(f1, sb.int_div(a, b)[1])
    zero = sb.int_sub(a, a)
    g1, _ = sb.int_div(a, zero)
    sb.assert_true(g1)
    inputs = []
    for i in range(40):
        av, bv = rng.randrange(p), rng.randrange(1, p)
        if i == 0:
            av = 0
        if i == 1:
            av, bv = p - 1, p - 1
        inputs.append([av, bv, (av + bv) % p, (av - bv) % p, av * bv % p, av * pow(bv, -1, p) % p])
    helpers.check_script(h2e, oracle, field, sb.words, inputs)


@pytest.mark.parametrize("field", FIELDS)
def test_overflowed_operands(h2e, oracle, field):
    """config 2 shape: operands with times in [2,16] feed reduce then int_mul; also raw int_mul on
    overflowed operands and the linear ops on them."""
    p = oracle.FIELD_MODULUS[field]
    L = 4 if field == 1 else 3
    lead_bits = p.bit_length() % 108
    rng = random.Random(200 + field)
    for ta, tb in [(2, 16), (7, 3), (16, 16), (63, 63), (1, 5)]:
        sb = h2e.ScriptBuilder()
        a = sb.load_int(ta, 0)
        b = sb.load_int(tb, L)
        ra, rb = sb.reduce(a), sb.reduce(b)
        sb.int_mul(ra, rb)
        sb.int_mul(a, b)
        if ta + tb < 60:
            s = sb.int_add(a, b)
            d = sb.int_sub(a, b)
            n = sb.int_neg(b)
            sb.int_mul(s, d)
            sb.int_mul(n, n)
        inputs = []
        for i in range(33):
            la = [rng.randrange(ta << 108) for _ in range(L - 1)] + [rng.randrange(ta << lead_bits)]
            lb = [rng.randrange(tb << 108) for _ in range(L - 1)] + [rng.randrange(tb << lead_bits)]
            if i == 0:  # maximal limbs
                la = [(ta << 108) - 1] * (L - 1) + [(ta << lead_bits) - 1]
                lb = [(tb << 108) - 1] * (L - 1) + [(tb << lead_bits) - 1]
            if i == 1:
                la = [0] * L
            inputs.append(la + lb)
        helpers.check_script(h2e, oracle, field, sb.words, inputs)


@pytest.mark.parametrize("field", FIELDS)
def test_misc_integer_and_base_ops(h2e, oracle, field):
    p = oracle.FIELD_MODULUS[field]
    r = oracle.MODULI["bn256_fr"]
    rng = random.Random(300 + field)
    sb = h2e.ScriptBuilder()
    a, b = sb.assign_w(0), sb.assign_w(1)
    k = sb.assign_int_constant(1, 0)
    kin = sb.assign_int_constant(0, 2)
    m3 = sb.mul_small_const(a, 3)
    m2 = sb.mul_small_const(b, 2)
    sb.is_int_zero(a)
    ce = sb.is_int_equal(a, b)
    ce2 = sb.is_int_equal(a, a)
    bi = sb.bisec_int(ce2, m3, m2)
    sb.bisec_int(ce, k, kin)
    inv = sb.int_unsafe_invert(b)
    sq = sb.int_square(bi)
    sb.int_mul(sq, inv)
    x, y = sb.assign(3), sb.assign(4)
    bit0, bit1 = sb.assign_bit(5), sb.assign_bit(6)
    cst = sb.assign_constant(1, 1)
    cin = sb.assign_constant(0, 3)
    s = sb.add(x, y)
    d = sb.sub(x, y)
    m = sb.mul(s, d)
    sb.is_zero(m)
    sb.is_zero(sb.sub(x, x))
    for f in (sb.and_, sb.or_, sb.xor, sb.xnor, sb.not_and):
        f(bit0, bit1)
    for f in (sb.or_, sb.xor, sb.xnor, sb.not_and):  # arbitrary field elements: the rows are computed in Fr (no bit fast path)
        f(x, y)
        f(bit0, y)
    nb = sb.not_(bit0)
    sb.bisec(nb, x, cst)
    sb.bisec(bit1, cin, y)
    sb.assert_equal(x, x)
    statics = [rng.randrange(p), rng.randrange(r)]
    inputs = []
    for i in range(34):
        inputs.append([rng.randrange(p) if i else 0, rng.randrange(1, p), rng.randrange(p), rng.randrange(r), rng.randrange(r),
                       rng.randrange(2), rng.randrange(2)])
    helpers.check_script(h2e, oracle, field, sb.words, inputs, statics)


def _run_in_schedule_order(shape, packed):
    """The team-mode program (HEAD/TAIL splits, merged is_int_zero TAILs, OP_DIV_INV + scratch) in level order."""
    sprog, _ = shape.schedule()
    return helpers.run_emulated(shape, packed, program=sprog)


@pytest.mark.parametrize("field", FIELDS)
def test_zero_denominators_and_zero_tests_through_the_team_schedule(h2e, oracle, field):
    """is_int_zero / int_div in their team-mode form: the condition cell comes from the HEAD (no inversion), the
    quotient cells from OP_DIV_HEAD_S, everything else from the deferred TAILs -- up to three is_int_zero blocks
    behind one Fr inversion, with zeros among the inverted values (value 0, value w, a == b) and a zero
    denominator (int_div by 0 yields c = 0 and the condition 1, integer_chip.rs:493-538)."""
    p = oracle.FIELD_MODULUS[field]
    rng = random.Random(900 + field)
    sb = h2e.ScriptBuilder()
    a, b, c = sb.assign_w(0), sb.assign_w(1), sb.assign_w(2)
    sb.is_int_zero(a)
    sb.is_int_zero(b)
    sb.is_int_equal(a, b)
    sb.is_int_equal(c, c)
    sb.is_int_zero(sb.int_sub(c, c))  # limbs hold a multiple of w before the reduce
    _, q = sb.int_div(a, b)           # seven zero tests in all: merged TAILs of 3 + 3 + 1 (L = 3) or 2 + 2 + 2 + 1
    sb.int_mul(q, c)
    sb.int_div(c, sb.int_add(a, b))
    inputs = [[0, 0, 5], [0, 7, 1], [9, 0, 2], [p - 1, p - 1, p - 1], [3, p - 3, 4], [1, 1, 0]]
    for _ in range(27):
        inputs.append([rng.randrange(p), rng.randrange(1, p), rng.randrange(p)])
    helpers.check_script(h2e, oracle, field, sb.words, inputs, runner=_run_in_schedule_order)


@pytest.mark.parametrize("field", FIELDS)
def test_merged_inversions_option_with_zero_denominators(h2e, oracle, field, monkeypatch):
    """H2E_DIVMERGE=3 (tuning option, off by default): the W inversions of one dependency level share ONE inversion
    (up to three denominators per OP_DIV_INV, two for the 4-limb field). A zero denominator among them must get the
    inverse 0 (int_div by 0 yields c = 0) without disturbing the other members."""
    monkeypatch.setenv("H2E_DIVMERGE", "3")
    p = oracle.FIELD_MODULUS[field]
    rng = random.Random(950 + field)
    sb = h2e.ScriptBuilder()
    v = [sb.assign_w(i) for i in range(6)]
    for i in range(0, 6, 2):  # three independent int_divs: the same dependency level
        _, q = sb.int_div(v[i], v[i + 1])
        sb.int_mul(q, v[i])
    shape = h2e.Shape.from_script(field, sb.words)
    sprog, _ = shape.schedule()
    ops = sprog[:, 0:2].copy().view(np.uint16).reshape(-1)
    k = sprog[ops == 33, 3] & 3
    assert int(k.sum()) == 3 and len(k) == (1 if field != 1 else 2), "three int_divs -> one merged inversion (L = 3) or 2 + 1 (L = 4)"
    inputs = [[1, 0, 2, 3, 4, 5], [1, 2, 3, 0, 5, 0], [0, 0, 0, 0, 0, 0], [p - 1, 1, 1, p - 1, 2, p - 2]]
    for _ in range(12):
        inputs.append([rng.randrange(p) for _ in range(6)])
    helpers.check_script(h2e, oracle, field, sb.words, inputs, runner=_run_in_schedule_order)


@pytest.mark.gpu
@pytest.mark.parametrize("field", [0, 1])
def test_zero_denominators_and_zero_tests_team_mode_gpu(h2e, oracle, field):
    """The same script on the GPU with team mode forced (2 CTAs per tile), bit-exact against the oracle."""
    p = oracle.FIELD_MODULUS[field]
    rng = random.Random(900 + field)
    sb = h2e.ScriptBuilder()
    a, b, c = sb.assign_w(0), sb.assign_w(1), sb.assign_w(2)
    sb.is_int_zero(a)
    sb.is_int_zero(b)
    sb.is_int_equal(a, b)
    sb.is_int_equal(c, c)
    sb.is_int_zero(sb.int_sub(c, c))
    _, q = sb.int_div(a, b)
    sb.int_mul(q, c)
    sb.int_div(c, sb.int_add(a, b))
    inputs = [[0, 0, 5], [0, 7, 1], [9, 0, 2], [p - 1, p - 1, p - 1], [3, p - 3, 4], [1, 1, 0]]
    for _ in range(27):
        inputs.append([rng.randrange(p), rng.randrange(1, p), rng.randrange(p)])

    def run_team(shape, packed):
        shape.set_mode(2, 2)
        return helpers.run_gpu(shape, packed)

    helpers.check_script(h2e, oracle, field, sb.words, inputs, runner=run_team)


def test_value_asserts_become_status_bits(h2e, oracle):
    """assert_int_equal on unequal values: the reference panics (base_chip.rs:375-379); the batch
    API reports it per instance instead."""
    p = oracle.FIELD_MODULUS[0]
    sb = h2e.ScriptBuilder()
    a, b = sb.assign_w(0), sb.assign_w(1)
    sb.assert_int_equal(a, b)
    inputs = [[5, 5], [5, 6], [p - 1, p - 1], [0, p - 1]]
    shape = h2e.Shape.from_script(0, sb.words)
    _, status = helpers.run_emulated(shape, h2e.pack_inputs(inputs))
    assert list(status) == [0, h2e.ST_ASSERT_VALUE, 0, h2e.ST_ASSERT_VALUE]
    assert oracle.run_script(0, sb.words, inputs[1]).status != 0  # the oracle "panics" like the reference


def test_load_int_packed_is_load_int(h2e, oracle):
    """load_int_packed (the L limbs of an operand in one 64-byte logical input, 16 bytes each) produces the same records as
    load_int (one logical input per limb) -- the bench's input layout."""
    import random

    import numpy as np

    rng = random.Random(8)
    L = 3
    sa = h2e.ScriptBuilder()
    sa.int_mul(sa.reduce(sa.load_int(5, 0)), sa.load_int(1, L))
    sp = h2e.ScriptBuilder()
    sp.int_mul(sp.reduce(sp.load_int_packed(5, 0)), sp.load_int_packed(1, 1))
    rows_a, rows_p = [], []
    for _ in range(3):
        la = [rng.randrange(5 << 108) for _ in range(L - 1)] + [rng.randrange(5 << 38)]
        lb = [rng.randrange(1 << 108) for _ in range(L - 1)] + [rng.randrange(1 << 38)]
        rows_a.append(la + lb)
        rows_p.append([sum(v << (128 * i) for i, v in enumerate(la)), sum(v << (128 * i) for i, v in enumerate(lb))])
    shape_p = helpers.check_script(h2e, oracle, 0, sp.words, rows_p)
    shape_a = h2e.Shape.from_script(0, sa.words)
    va, _ = helpers.run_emulated(shape_a, h2e.pack_inputs(rows_a))
    vp, _ = helpers.run_emulated(shape_p, h2e.pack_inputs(rows_p))
    assert np.array_equal(va[0][:, :3], vp[0][:, :3])
