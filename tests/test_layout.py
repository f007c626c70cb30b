"""Record layouts (csrc/layout.h) on the CPU: the static width table against the macro-op code's own stores (host
emulator built with -DH2E_WIDTH_PROBE), the copy classes of the UNIQUE form against actual values, and the
consumer-side expansion (h2e_records_expand) round trip. Plus the structural counts of the BASELINE-size shapes
(SURVEY Appendix B), which pin the traced shapes at the sizes the benchmark runs."""
import numpy as np
import pytest

import circuits_util as cu
import ecmath as em
import helpers


def _pack(vals, off, width, root, fmt, h2e, der_src=None):
    """numpy restatement of the export kernel: WIDE tiles -> COMPACT / UNIQUE / PRIMARY words [tiles, words_per_lane * 32]"""
    tiles, n_slots = vals.shape[0], vals.shape[1]
    v = vals.view(np.uint32).reshape(tiles, n_slots, 32, 8)
    out = np.zeros((tiles, int(off[-1]) * 32), dtype=np.uint32)
    for s in range(n_slots):
        if fmt in (h2e.REC_UNIQUE, h2e.REC_PRIMARY) and root[s] != s:
            continue
        if fmt == h2e.REC_PRIMARY and der_src[s] != 0xFFFFFFFF:
            continue
        w = int(width[s])
        out[:, int(off[s]) * 32:(int(off[s]) + w) * 32] = v[:, s, :, :w].reshape(tiles, 32 * w)
    return out


def _script_shape(h2e):
    sb = h2e.ScriptBuilder()
    a = sb.assign_w(0)
    b = sb.assign_w(1)
    c = sb.int_mul(a, b)
    d = sb.int_sub(sb.int_add(c, a), b)
    z, e = sb.int_div(d, b)
    sb.assert_int_equal(sb.int_mul(e, b), d)
    f = sb.bisec_int(z, sb.int_neg(a), sb.mul_small_const(b, 5))
    sb.is_int_equal(sb.reduce(f), a)
    x = sb.assign(2)
    y = sb.assign_bit(3)
    sb.bisec(y, sb.mul(x, x), sb.add(x, x))
    sb.is_zero(sb.sub(x, x))
    sb.xor(y, sb.or_(y, sb.not_(y)))
    return sb


@pytest.mark.parametrize("field", [0, 1, 2])
def test_width_table_matches_macro_op_code(h2e, oracle, field):
    sb = _script_shape(h2e)
    shape = h2e.Shape.from_script(field, sb.words)
    _, width, _ = shape.layout(h2e.REC_COMPACT)
    assert np.array_equal(width, helpers.emu_probe_widths(shape))


@pytest.mark.parametrize("kind,params", [(0, [3]), (1, [2]), (4, [1]), (2, [])])
def test_width_table_matches_macro_op_code_circuits(h2e, kind, params):
    shape = h2e.Shape.build(kind, params)
    _, width, _ = shape.layout(h2e.REC_COMPACT)
    assert np.array_equal(width, helpers.emu_probe_widths(shape))


def test_unique_and_compact_are_lossless(h2e, oracle):
    """Every copy holds its root's value, every value fits its width class, and the host-side expansion rebuilds the WIDE
    tiles bit for bit from both forms; the dense per-instance arrays are the oracle's advice records."""
    import random

    rng = random.Random(11)
    p = oracle.FIELD_MODULUS[0]
    sb = _script_shape(h2e)
    inputs = [[rng.randrange(p), rng.randrange(1, p), rng.randrange(1 << 200), rng.randrange(2)] for _ in range(37)]
    shape = h2e.Shape.from_script(0, sb.words)
    vals, status = helpers.run_emulated(shape, h2e.pack_inputs(inputs))
    assert (status == 0).all()
    n = len(inputs)
    der_src, der_shift = shape.layout_derived()
    assert (der_src != 0xFFFFFFFF).sum() > 0.3 * shape.n_slots  # the range chunks: 60 of the 125 cells of an int_mul block
    for fmt in (h2e.REC_COMPACT, h2e.REC_UNIQUE, h2e.REC_PRIMARY):
        off, width, root = shape.layout(fmt)
        v = vals.view(np.uint32).reshape(vals.shape[0], shape.n_slots, 32, 8)
        lanes = np.arange(vals.shape[0] * 32) < n
        for s in range(shape.n_slots):
            assert root[s] <= s and root[root[s]] == root[s]
            live = v[:, s].reshape(-1, 8)[lanes]
            assert not live[:, int(width[s]):].any(), f"slot {s} exceeds its width class"
            assert np.array_equal(live, v[:, root[s]].reshape(-1, 8)[lanes]), f"slot {s} differs from its root {root[s]}"
            if der_src[s] != 0xFFFFFFFF:
                # a derived cell is an 18-bit field of a stored cell (a root that is not itself derived)
                src = int(der_src[s])
                assert root[s] == s and root[src] == src and der_src[src] == 0xFFFFFFFF and width[s] == 1
                big = [int.from_bytes(x.tobytes(), "little") for x in v[:, src].reshape(-1, 8)[lanes]]
                want = [0 if der_shift[s] == 255 else (b >> int(der_shift[s])) & 0x3FFFF for b in big]
                assert [int(x) for x in live[:, 0]] == want, f"slot {s} is not bits {der_shift[s]}.. of slot {src}"
        rec = _pack(vals, off, width, root, fmt, h2e, der_src)
        assert rec.nbytes == shape.records_bytes(fmt, n)
        back = shape.records_expand(rec.view(np.uint8).reshape(-1), fmt, n, threads=3)
        # (padding lanes of the last tile are not emulated: compare the live lanes)
        assert np.array_equal(back.reshape(-1, shape.n_slots, 32, 32)[:, :, :, :].reshape(vals.shape)[0], vals[0])
        assert np.array_equal(back[1][:, : n - 32], vals[1][:, : n - 32])
        cells = shape.slot_cells()
        for mode in (h2e.EXPAND_COLUMNS, h2e.EXPAND_ROWS):
            dense = shape.records_expand(rec.view(np.uint8).reshape(-1), fmt, n, mode=mode, threads=2)
            assert dense.shape == (n, shape.dense_cells(), 32)
            heights = [shape.base_height, shape.range_height, shape.select_height]
            for inst in (0, 36):
                r = oracle.run_script(0, sb.words, inputs[inst])
                base = 0
                for reg in range(3):
                    cols = h2e.ADV_COLS[reg]
                    block = dense[inst, base:base + cols * heights[reg]]
                    got = block.reshape(cols, heights[reg], 32).transpose(1, 0, 2) if mode == h2e.EXPAND_COLUMNS else block.reshape(heights[reg], cols, 32)
                    want = r.adv[reg][: heights[reg]] * (r.advf[reg][: heights[reg]] & 1)[:, :, None]
                    assert np.array_equal(got, want), (fmt, mode, inst, reg)
                    base += cols * heights[reg]
    assert shape.records_bytes(h2e.REC_UNIQUE, n) < 0.25 * shape.vals_bytes(n)
    assert shape.records_bytes(h2e.REC_PRIMARY, n) < 0.8 * shape.records_bytes(h2e.REC_UNIQUE, n)
    assert shape.records_bytes(h2e.REC_COMPACT, n) < 0.5 * shape.vals_bytes(n)
    _ = cells


def test_unique_roots_on_msm_with_select_chip(h2e):
    """Select-chip rows record their permutation pair as (new cell, source cell): the root is still the older slot."""
    shape = h2e.Shape.build(0, [2])
    rows = [cu.msm_inputs(em.BN256, 2, 77)]
    vals, status = helpers.run_emulated(shape, h2e.pack_inputs(rows))
    assert status[0] == 0
    _, width, root = shape.layout(h2e.REC_UNIQUE)
    v = vals.view(np.uint32).reshape(1, shape.n_slots, 32, 8)[0, :, 0, :]
    assert (root <= np.arange(shape.n_slots)).all()
    assert np.array_equal(v, v[root])
    assert not np.where(np.arange(8)[None, :] >= width[:, None], v, 0).any()


# ---- SURVEY Appendix B at the BASELINE sizes (shape pass only: no values) -----------------------------------
# (base rows, range rows, select rows, permutation pairs, advice cells, int_mul, int_div, reduce)
APPENDIX_B = {
    (0, 1000): (6292311, 6634908, 457600, 15616155, None, 118285, 57510, 141382),   # configs[0]: native_scalar_ecc_chip.rs:13-61
    (1, 400): (6543111, 5828508, 0, 17889355, None, 105285, 51910, 118382),          # native_scalar_ecc_chip.rs:63-110
    (4, 50): (692860, 732076, 35600, 1795972, None, 8885, 4200, 10051),              # general_scalar_ecc_chip.rs:14-49
    (2, 0): (1049946, 1103352, 0, None, 6165013, 28872, 1, 24573),                   # configs[3] (17 rows less than SURVEY's table)
    (3, 0): (1300575, 1433618, 0, 3524865, 7952811, 25043, 1, 21207),                # configs[4]
}


@pytest.mark.parametrize("key", [(0, 1000), (1, 400), (4, 50), (2, 0), (3, 0)])
def test_appendix_b_counts_at_baseline_sizes(h2e, key):
    kind, n = key
    shape = h2e.Shape.build(kind, [n] if n else [])
    base, rng_rows, sel, perms, cells, n_mul, n_div, n_red = APPENDIX_B[key]
    assert (shape.base_offset, shape.range_offset, shape.select_offset) == (base, rng_rows, sel)
    if perms is not None:
        assert shape.n_perms == perms
    if cells is not None:
        assert shape.n_slots == cells
    ops = shape.program()[:, 0:2].copy().view(np.uint16).reshape(-1)
    assert (int((ops == 9).sum()), int((ops == 10).sum()), int((ops == 8).sum())) == (n_mul, n_div, n_red)
