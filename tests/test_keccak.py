"""Keccak chip (KeccakChipOps, src/circuit/keccak_chip.rs:53-307) through the op-script: the oracle's restatement is
pinned against an independent byte-level Keccak-256 (with published known answers), and the product (tracer +
witness VM: host emulator here, the GPU in the `gpu` tests) against the oracle, cell by cell."""
import random

import numpy as np
import pytest

import helpers

RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
      0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
      0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
      0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
M64 = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & M64 if n else v


def keccak_f(a):
    """Keccak-f[1600] on a[x][y] 64-bit lanes (FIPS 202 section 3.2, written from the specification)"""
    for rnd in range(24):
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        x, y = 1, 0
        b[0][0] = a[0][0]
        for t in range(24):
            b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], (t + 1) * (t + 2) // 2)
            x, y = y, (2 * x + 3 * y) % 5
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= RC[rnd]
    return a


def keccak256(data: bytes) -> bytes:
    """original Keccak-256 (padding 0x01 .. 0x80, rate 136 bytes), as Ethereum uses it"""
    rate = 136
    p = bytearray(data)
    pad = rate - len(p) % rate
    p += b"\x81" if pad == 1 else b"\x01" + b"\x00" * (pad - 2) + b"\x80"
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(p), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(p[off + 8 * i:off + 8 * i + 8], "little")
        a = keccak_f(a)
    return b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


def test_python_keccak256_known_answers():
    assert keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    assert keccak256(b"a" * 200).hex() != keccak256(b"a" * 199).hex()


def _hash_script(h2e, n):
    sb = h2e.ScriptBuilder()
    vals = [sb.assign(i) for i in range(n)]
    sb.keccak_hash(vals)
    return sb


def _last_base_cell(rec):
    return int.from_bytes(rec.adv[0][rec.base_offset - 1, 4].tobytes(), "little")


@pytest.mark.parametrize("n", [1, 5])
def test_oracle_keccak_hash_is_keccak256(h2e, oracle, n):
    """hash(inputs) == Keccak-256(32-byte big-endian encodings) as a big-endian integer mod r; n = 5 absorbs two blocks.
    Every row the oracle writes passes the gate checker."""
    rng = random.Random(20 + n)
    r = h2e.FR_MODULUS
    sb = _hash_script(h2e, n)
    for inp in ([rng.randrange(r) for _ in range(n)], [0] * n, [r - 1] * n):
        rec = oracle.run_script(0, sb.words, inp)
        assert rec.status == 0, rec.error
        assert rec.gate_ok, rec.gate_msg
        want = int.from_bytes(keccak256(b"".join(v.to_bytes(32, "big") for v in inp)), "big") % r
        assert _last_base_cell(rec) == want
    # one permutation = 24 x (theta 3200 + chi 3200) xor / not_and rows + the iota `not` rows
    assert rec.base_offset > n * 385 + (2 if n == 5 else 1) * 24 * 6400


def test_keccak_hash_product_matches_oracle_emulated(h2e, oracle):
    rng = random.Random(31)
    r = h2e.FR_MODULUS
    sb = _hash_script(h2e, 2)
    inputs = [[rng.randrange(r), rng.randrange(r)], [0, r - 1], [1, 1 << 200]]
    shape = helpers.check_script(h2e, oracle, 0, sb.words, inputs)
    assert 1500 < shape.n_instr < 4096  # one macro-op per 64-bit lane of xor / chi rows, one per `not` (155 k rows)


def _steps_script(h2e):
    """the trait's pieces one by one: decompose, init, absorb (incl. one full permutation), single steps, lanes, compose"""
    sb = h2e.ScriptBuilder()
    vals = [sb.assign(i) for i in range(4)]
    bits = []
    for v in vals:
        bits += sb.keccak_decompose_u256_be(v)
    one = sb.assign_constant(1, 0)
    zero = sb.assign_constant(1, 1)
    bits += [zero] * 7 + [one] + [zero] * (1088 - 1024 - 16) + [one] + [zero] * 7
    st = sb.keccak_init()
    sb.keccak_absorb(st, bits)
    sb.keccak_theta(st)
    sb.keccak_rho_and_pi(st)
    sb.keccak_xi(st)
    sb.keccak_iota(st, 3)
    lane = sb.keccak_lane(st, 2, 1) + sb.keccak_lane(st, 0, 0)
    sb.keccak_compose_to_scalar_be(lane)
    return sb


def test_keccak_trait_pieces_emulated(h2e, oracle):
    rng = random.Random(32)
    r = h2e.FR_MODULUS
    sb = _steps_script(h2e)
    inputs = [[rng.randrange(r) for _ in range(4)] for _ in range(2)]
    helpers.check_script(h2e, oracle, 0, sb.words, inputs, statics=[1, 0])


def test_keccak_vector_macro_ops_and_width_table(h2e):
    """The 64 rows of a lane run as ONE macro-op (OP_BOOLV for theta / absorb, OP_CHIV for chi): a permutation is ~1.9 k
    instructions instead of ~155 k; the static width table covers them (checked against the macro-op code's own stores)."""
    sb = _steps_script(h2e)
    shape = h2e.Shape.from_script(0, sb.words, [1, 0])
    ops = shape.program()[:, 0:2].copy().view(np.uint16).reshape(-1)
    OP_BOOL, OP_BOOLV, OP_CHIV = 20, 39, 40
    # absorb: 17 lanes; 24 rounds + the single steps: theta = 5 x 4 + 5 + 25 vector xors, chi = 25 vector ops
    assert int((ops == OP_BOOLV).sum()) == 17 + 25 * 50 and int((ops == OP_CHIV).sum()) == 25 * 25 and not (ops == OP_BOOL).any()
    _, width, _ = shape.layout(h2e.REC_COMPACT)
    assert np.array_equal(width, helpers.emu_probe_widths(shape))


def test_keccak_bad_records_are_rejected(h2e):
    sb = h2e.ScriptBuilder()
    v = sb.assign(0)
    st = sb.keccak_init()
    sb._emit("KECCAK_STEP", st, 7, 0)
    with pytest.raises(h2e.H2EError):
        h2e.Shape.from_script(0, sb.words)
    sb = h2e.ScriptBuilder()
    v = sb.assign(0)
    sb._emit("KECCAK_ABSORB", 0, v)  # wrong arity
    with pytest.raises(h2e.H2EError):
        h2e.Shape.from_script(0, sb.words)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_keccak_hash_gpu(h2e, oracle, mode):
    """One-block and two-block hashes on the GPU (mode 0: the library's choice = team mode for this 155k-op program;
    1: one thread per instance), several tiles, instances compared with the oracle bit for bit."""
    rng = random.Random(33 + mode)
    r = h2e.FR_MODULUS
    n = 2 if mode == 0 else 1
    sb = _hash_script(h2e, n)
    n_inst = 70 if mode == 0 else 33
    inputs = [[rng.randrange(r) for _ in range(n)] for _ in range(n_inst)]
    inputs[1] = [0] * n
    shape = h2e.Shape.from_script(0, sb.words)
    if mode:
        shape.set_mode(1, 0)
    vals, status = helpers.run_gpu(shape, h2e.pack_inputs(inputs))
    assert (status == 0).all()
    cells = None
    for i in (0, 1, 31, 32, n_inst - 1):
        rec = oracle.run_script(0, sb.words, inputs[i])
        if cells is None:
            cells = helpers.compare_static(shape, rec)
        helpers.compare_instance(shape, cells, vals, i, rec)
        want = int.from_bytes(keccak256(b"".join(v.to_bytes(32, "big") for v in inputs[i])), "big") % r
        assert _last_base_cell(rec) == want


@pytest.mark.gpu
def test_keccak_hash_gpu_team_groups(h2e, oracle):
    """A batch between one cooperative launch (37 tiles) and 4 tiles per SM: this shape's macro-ops are large (470 k cells per
    instance), so the batch runs as team groups, not as one thread per instance; instances of both groups are compared."""
    rng = random.Random(36)
    r = h2e.FR_MODULUS
    sb = _hash_script(h2e, 1)
    n_inst = 41 * 32 - 5
    inputs = [[rng.randrange(r)] for _ in range(n_inst)]
    shape = h2e.Shape.from_script(0, sb.words)
    before = h2e.lib().h2e_launch_count()
    vals, status = helpers.run_gpu(shape, h2e.pack_inputs(inputs))
    assert h2e.lib().h2e_launch_count() - before == 3  # two cooperative launches (21 + 20 tiles) + the expansion to 32-byte cells
    assert (status == 0).all()
    cells = None
    for i in (0, 21 * 32 - 1, 21 * 32, n_inst - 1):
        rec = oracle.run_script(0, sb.words, inputs[i])
        if cells is None:
            cells = helpers.compare_static(shape, rec)
        helpers.compare_instance(shape, cells, vals, i, rec)


@pytest.mark.gpu
def test_keccak_trait_pieces_gpu(h2e, oracle):
    rng = random.Random(35)
    r = h2e.FR_MODULUS
    sb = _steps_script(h2e)
    inputs = [[rng.randrange(r) for _ in range(4)] for _ in range(34)]
    helpers.check_script(h2e, oracle, 0, sb.words, [inputs[0], inputs[33]], statics=[1, 0], runner=helpers.run_gpu)
