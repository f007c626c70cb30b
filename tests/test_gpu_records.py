"""GPU tests of the record formats, the streaming host path and the device-side prover hand-off (through the C ABI)."""
import random

import numpy as np
import pytest

import circuits_util as cu
import ecmath as em
import helpers

pytestmark = pytest.mark.gpu


def _int_script(h2e):
    sb = h2e.ScriptBuilder()
    a = sb.assign_w(0)
    b = sb.assign_w(1)
    c = sb.int_mul(a, b)
    d = sb.int_sub(sb.int_add(c, a), b)
    z, e = sb.int_div(d, b)
    sb.assert_int_equal(sb.int_mul(e, b), d)
    sb.bisec_int(z, sb.int_neg(a), sb.mul_small_const(b, 5))
    return sb


def _inputs(oracle, n, seed=3):
    rng = random.Random(seed)
    p = oracle.FIELD_MODULUS[0]
    return [[rng.randrange(p), rng.randrange(1, p)] for _ in range(n)]


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_host_records_formats_expand_to_the_wide_records(h2e, oracle, fmt):
    sb = _int_script(h2e)
    inputs = _inputs(oracle, 70)
    shape = helpers.check_script(h2e, oracle, 0, sb.words, inputs[:3], runner=helpers.run_gpu)
    packed = h2e.pack_inputs(inputs)
    full, st = shape.run_host(packed)
    rec, st2 = shape.run_host_records(packed, fmt)
    assert (st == 0).all() and (st2 == 0).all()
    assert rec.nbytes == shape.records_bytes(fmt, len(inputs))
    assert np.array_equal(shape.records_expand(rec, fmt, len(inputs)), full)
    shape.compact_prepare()  # cross-check of the static width table against the device code's own stores
    assert rec.nbytes < (0.5 if fmt == 1 else 0.25) * full.nbytes


def test_stream_api_ring_of_chunks(h2e, oracle):
    """h2e_stream_*: many chunks through a ring of three pinned host buffers, tickets polled, every chunk bit-exact."""
    import torch

    sb = _int_script(h2e)
    shape = h2e.Shape.from_script(0, sb.words)
    n_chunks, ring = 7, 3
    st = shape.open_stream(h2e.REC_PRIMARY, chunk_bytes_hint=64 * shape.vals_bytes(32))  # 64 tiles per chunk
    assert st.chunk_instances == 64 * 32 and st.in_flight == 2 and not st.pieces
    inputs = [h2e.pack_inputs(_inputs(oracle, st.chunk_instances - (5 if c == n_chunks - 1 else 0), seed=100 + c)) for c in range(n_chunks)]
    bufs = [torch.empty((st.chunk_bytes,), dtype=torch.uint8).pin_memory().numpy() for _ in range(ring)]
    stat = [np.zeros(st.chunk_instances, dtype=np.uint32) for _ in range(ring)]
    pin_in = [torch.from_numpy(x).pin_memory().numpy() for x in inputs]
    tickets = {}
    checked = 0

    def consume(c):
        nonlocal checked
        st.wait(tickets.pop(c))
        n = inputs[c].shape[0]
        assert (stat[c % ring][:n] == 0).all()
        got = shape.records_expand(bufs[c % ring], h2e.REC_PRIMARY, n)
        want, s_w = shape.run(torch.from_numpy(inputs[c]).cuda())
        torch.cuda.synchronize()
        want = want.cpu().numpy()
        full_tiles = n // 32
        assert np.array_equal(got[:full_tiles], want[:full_tiles])
        if n % 32:
            assert np.array_equal(got[full_tiles][:, : n % 32], want[full_tiles][:, : n % 32])
        checked += 1

    for c in range(n_chunks):
        if c >= ring:
            consume(c - ring)  # the buffer is reused only after its ticket completed
        tickets[c] = st.submit(pin_in[c], bufs[c % ring], stat[c % ring])
    assert not st.poll(tickets[n_chunks - 1]) or True
    for c in sorted(list(tickets)):
        consume(c)
    assert checked == n_chunks
    st.close()


@pytest.mark.parametrize("fmt", [0, 2, 3])
def test_stream_slot_range_pieces(h2e, oracle, fmt, monkeypatch):
    """A tile whose records exceed the staging buffer (a 4096-point MSM) is exported in slot-range pieces with 2-D copies;
    H2E_STREAM_PIECE_WORDS forces that path on a small shape: several pieces per chunk, two chunks, every format that is
    derived from the VM's own records; the landed records must expand to the cells of the one-call device path."""
    import torch

    sb = _int_script(h2e)
    shape = h2e.Shape.from_script(0, sb.words)
    monkeypatch.setenv("H2E_STREAM_PIECE_WORDS", str(5 * 96))  # 5 tiles per chunk: <= 96 words per lane and tile per piece
    st = shape.open_stream(fmt, chunk_bytes_hint=5 * shape.vals_bytes(32))
    assert st.chunk_instances == 5 * 32 and st.pieces
    words_per_lane = st.tile_bytes // (32 * 4)
    assert words_per_lane > 3 * 96, "the shape must need several pieces"
    for c, n in enumerate((160, 131)):
        inputs = h2e.pack_inputs(_inputs(oracle, n, seed=300 + c))
        buf = torch.empty((st.chunk_bytes,), dtype=torch.uint8).pin_memory().numpy()
        stat = np.zeros(st.chunk_instances, dtype=np.uint32)
        st.wait(st.submit(torch.from_numpy(inputs).pin_memory().numpy(), buf, stat))
        assert (stat[:n] == 0).all()
        got = shape.records_expand(buf, fmt, n)
        want, _ = shape.run(torch.from_numpy(inputs).cuda())
        torch.cuda.synchronize()
        want = want.cpu().numpy()
        full = n // 32
        assert np.array_equal(got[:full], want[:full])
        if n % 32:
            assert np.array_equal(got[full][:, : n % 32], want[full][:, : n % 32])
    st.close()


@pytest.mark.parametrize("fmt", [0, 3])
def test_stream_shared_record_buffer_pipeline(h2e, oracle, fmt, monkeypatch):
    """When two chunks of record tiles do not fit next to their staging buffers (a 4096-point MSM tile), both pipeline slots
    compute into ONE record buffer and a chunk's VM waits for the previous chunk's export kernel; H2E_STREAM_SHARED forces that
    on a small shape. Five chunks in flight two deep, different inputs per chunk: every chunk must land its own records."""
    import torch

    sb = _int_script(h2e)
    shape = h2e.Shape.from_script(0, sb.words)
    monkeypatch.setenv("H2E_STREAM_SHARED", "1")
    st = shape.open_stream(fmt, chunk_bytes_hint=6 * shape.vals_bytes(32))
    assert st.chunk_instances == 6 * 32 and st.in_flight == 2
    n_chunks = 5
    inputs = [h2e.pack_inputs(_inputs(oracle, st.chunk_instances - 3 * c, seed=400 + c)) for c in range(n_chunks)]
    pins = [torch.from_numpy(x).pin_memory().numpy() for x in inputs]
    bufs = [torch.empty((st.chunk_bytes,), dtype=torch.uint8).pin_memory().numpy() for _ in range(n_chunks)]
    stat = [np.zeros(st.chunk_instances, dtype=np.uint32) for _ in range(n_chunks)]
    tickets = [st.submit(pins[c], bufs[c], stat[c]) for c in range(n_chunks)]
    for c in range(n_chunks):
        st.wait(tickets[c])
        n = inputs[c].shape[0]
        assert (stat[c][:n] == 0).all()
        got = shape.records_expand(bufs[c], fmt, n)
        want, _ = shape.run(torch.from_numpy(inputs[c]).cuda())
        torch.cuda.synchronize()
        want = want.cpu().numpy()
        full = n // 32
        assert np.array_equal(got[:full], want[:full]), c
        if n % 32:
            assert np.array_equal(got[full][:, : n % 32], want[full][:, : n % 32]), c
    st.close()


def test_stream_rejects_oversized_chunk(h2e):
    sb = _int_script(h2e)
    shape = h2e.Shape.from_script(0, sb.words)
    st = shape.open_stream(h2e.REC_COMPACT, chunk_bytes_hint=shape.vals_bytes(32))
    assert st.chunk_instances == 32
    inp = np.zeros((33, shape.n_input_cells, 32), dtype=np.uint8)
    with pytest.raises(h2e.H2EError):
        st.submit(inp, np.zeros(2 * st.tile_bytes, dtype=np.uint8), np.zeros(33, dtype=np.uint32))
    st.close()


def test_is_zero_with_zero_and_nonzero_values_in_one_tile(h2e, oracle):
    """BaseChipOps::is_zero / invert (base_chip.rs:298-325) where the 32 instances of a tile disagree on a == 0: the
    inversion's warp votes must see every lane (thread mode and team mode)."""
    sb = h2e.ScriptBuilder()
    x = sb.assign(0)
    y = sb.assign(1)
    z = sb.is_zero(sb.sub(x, y))
    sb.bisec(z, x, sb.mul(x, y))
    rng = random.Random(5)
    r = h2e.FR_MODULUS
    inputs = []
    for i in range(64):
        a = rng.randrange(r)
        inputs.append([a, a if (i * 7) % 3 == 0 else rng.randrange(r)])
    helpers.check_script(h2e, oracle, 0, sb.words, inputs, runner=helpers.run_gpu)

    def team(shape, packed):
        shape.set_mode(2, 2)
        return helpers.run_gpu(shape, packed)

    # long enough for team mode to make sense: repeat the block
    sb2 = h2e.ScriptBuilder()
    x = sb2.assign(0)
    y = sb2.assign(1)
    for _ in range(40):
        z = sb2.is_zero(sb2.sub(x, y))
        x = sb2.bisec(z, x, sb2.mul(x, y))
    helpers.check_script(h2e, oracle, 0, sb2.words, inputs[:40], runner=team)


def test_records_scatter_on_device(h2e, oracle):
    """Prover hand-off: dense per-instance advice arrays on the device, column-major and row-major, canonical and Montgomery."""
    import torch

    sb = _int_script(h2e)
    inputs = _inputs(oracle, 45, seed=9)
    shape = h2e.Shape.from_script(0, sb.words)
    vals, st = shape.run_records(torch.from_numpy(h2e.pack_inputs(inputs)).cuda(), h2e.REC_COMPACT)
    heights = [shape.base_height, shape.range_height, shape.select_height]
    R = h2e.FR_MODULUS
    for order in (h2e.EXPAND_COLUMNS, h2e.EXPAND_ROWS):
        for enc in (h2e.EXPORT_CANONICAL, h2e.EXPORT_MONTGOMERY):
            dense = shape.records_scatter(vals, len(inputs), order=order, encoding=enc)
            torch.cuda.synchronize()
            dense = dense.cpu().numpy()
            for inst in (0, 31, 44):
                r = oracle.run_script(0, sb.words, inputs[inst])
                base = 0
                for reg in range(3):
                    cols = h2e.ADV_COLS[reg]
                    block = dense[inst, base:base + cols * heights[reg]]
                    got = block.reshape(cols, heights[reg], 32).transpose(1, 0, 2) if order == h2e.EXPAND_COLUMNS else block.reshape(heights[reg], cols, 32)
                    want = r.adv[reg][: heights[reg]] * (r.advf[reg][: heights[reg]] & 1)[:, :, None]
                    if enc == h2e.EXPORT_MONTGOMERY:
                        flat = want.reshape(-1, 32)
                        conv = np.zeros_like(flat)
                        for i in range(flat.shape[0]):
                            conv[i] = np.frombuffer(((int.from_bytes(flat[i].tobytes(), "little") << 256) % R).to_bytes(32, "little"), dtype=np.uint8)
                        want = conv.reshape(want.shape)
                    assert np.array_equal(got, want), (order, enc, inst, reg)
                    base += cols * heights[reg]


def test_long_program_runs_in_team_groups_beyond_one_launch(h2e, oracle):
    """More tiles than one cooperative launch holds (SMs / 2): the batch entry splits into groups instead of falling back
    to one thread per instance; every group is compared with the oracle on sampled instances."""
    import torch

    sb = h2e.ScriptBuilder()
    a = sb.assign_w(0)
    b = sb.assign_w(1)
    x = a
    for i in range(4200):  # > 4096 macro-ops: a "long" program (lazy reductions are inserted by the tracer as in the reference)
        x = sb.int_add(x, b) if i % 3 else sb.int_sub(x, a)
    shape = h2e.Shape.from_script(0, sb.words)
    assert shape.n_instr >= 4096
    n_inst = 32 * 80 - 7  # 80 tiles > 37: three cooperative launches
    inputs = _inputs(oracle, 16, seed=21)
    rows = [inputs[i % 16] for i in range(n_inst)]
    launches0 = h2e.lib().h2e_launch_count()
    rec_c, status = shape.run_records(torch.from_numpy(h2e.pack_inputs(rows)).cuda(), h2e.REC_COMPACT)
    torch.cuda.synchronize()
    assert h2e.lib().h2e_launch_count() - launches0 == 3  # three cooperative launches of 27, 27 and 26 tiles
    assert int(status[:n_inst].abs().max()) == 0
    cells = None
    for inst in (0, 27 * 32 - 1, 27 * 32, 54 * 32 + 5, n_inst - 1):
        rec = oracle.run_script(0, sb.words, rows[inst])
        if cells is None:
            cells = helpers.compare_static(shape, rec)
        t = inst // 32
        tb = shape.records_bytes(h2e.REC_COMPACT, 32)
        tile = shape.records_expand(rec_c[t * tb:(t + 1) * tb].cpu().numpy(), h2e.REC_COMPACT, 32)[0]
        helpers.compare_instance(shape, cells, {t: tile}, inst, rec)


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
def test_device_records_formats(h2e, oracle, fmt):
    """h2e_batch_run_records: every format on the device expands to the same cells, which are the oracle's."""
    import torch

    sb = _int_script(h2e)
    inputs = _inputs(oracle, 41, seed=33)
    shape = helpers.check_script(h2e, oracle, 0, sb.words, inputs[:2], runner=helpers.run_gpu)
    d_in = torch.from_numpy(h2e.pack_inputs(inputs)).cuda()
    wide, st = shape.run(d_in)
    rec, st2 = shape.run_records(d_in, fmt)
    torch.cuda.synchronize()
    assert int(st[:41].abs().max()) == 0 and int(st2[:41].abs().max()) == 0
    got = shape.records_expand(rec.cpu().numpy(), fmt, len(inputs))
    want = wide.cpu().numpy()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1][:, :9], want[1][:, :9])
