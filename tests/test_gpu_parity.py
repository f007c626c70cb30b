"""GPU parity (run on the B200 box): the CUDA witness VM, called through the C ABI, must be
bit-exact with the oracle for every advice cell and fixed cell, on the same seeded inputs."""
import random

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _int_mul_script(h2e, L, ta, tb, with_reduce):
    sb = h2e.ScriptBuilder()
    a = sb.load_int(ta, 0)
    b = sb.load_int(tb, L)
    if with_reduce:
        a, b = sb.reduce(a), sb.reduce(b)
    sb.int_mul(a, b)
    return sb


def _limbs(rng, L, t, lead_bits):
    return [rng.randrange(t << 108) for _ in range(L - 1)] + [rng.randrange(t << lead_bits)]


@pytest.mark.parametrize("field", [0, 1, 2])
@pytest.mark.parametrize("host_api", [False, True])
def test_int_mul_reduce_microbench_shape(h2e, oracle, field, host_api):
    """BASELINE config 2: half reduced operands, half overflowed (times in [2,16]) -> reduce -> int_mul."""
    p = oracle.FIELD_MODULUS[field]
    L = 4 if field == 1 else 3
    lead = p.bit_length() % 108
    rng = random.Random(11 + field)
    runner = (lambda s, i: helpers.run_gpu(s, i, host_api=True)) if host_api else helpers.run_gpu
    # reduced half
    sb = _int_mul_script(h2e, L, 1, 1, False)
    inputs = []
    for i in range(70):
        a, b = rng.randrange(p), rng.randrange(p)
        if i == 0:
            a, b = p - 1, p - 1
        if i == 1:
            a = 0
        inputs.append([(a >> (108 * k)) & ((1 << 108) - 1) for k in range(L)] + [(b >> (108 * k)) & ((1 << 108) - 1) for k in range(L)])
    helpers.check_script(h2e, oracle, field, sb.words, inputs, runner=runner)
    # overflowed half
    for ta, tb in [(2, 16), (16, 9)]:
        sb = _int_mul_script(h2e, L, ta, tb, True)
        inputs = [_limbs(rng, L, ta, lead) + _limbs(rng, L, tb, lead) for _ in range(45)]
        inputs[0] = [(ta << 108) - 1] * (L - 1) + [(ta << lead) - 1] + [(tb << 108) - 1] * (L - 1) + [(tb << lead) - 1]
        helpers.check_script(h2e, oracle, field, sb.words, inputs, runner=runner)


@pytest.mark.parametrize("field", [0, 1, 2])
def test_reference_integer_chip_test_shape_gpu(h2e, oracle, field):
    """src/tests/integer_chip.rs:11-55 on the GPU."""
    p = oracle.FIELD_MODULUS[field]
    rng = random.Random(100 + field)
    sb = h2e.ScriptBuilder()
    a, b = sb.assign_w(0), sb.assign_w(1)
    sb.assert_int_equal(sb.assign_w(2), sb.int_add(a, b))
    sb.assert_int_equal(sb.assign_w(3), sb.int_sub(a, b))
    sb.assert_int_equal(sb.assign_w(4), sb.int_mul(a, b))
    sb.assert_int_equal(sb.assign_w(5), sb.int_div(a, b)[1])
    zero = sb.int_sub(a, a)
    g1, _ = sb.int_div(a, zero)
    sb.assert_true(g1)
    inputs = []
    for i in range(50):
        av, bv = rng.randrange(p), rng.randrange(1, p)
        if i == 0:
            av = 0
        if i == 1:
            av, bv = p - 1, p - 1
        inputs.append([av, bv, (av + bv) % p, (av - bv) % p, av * bv % p, av * pow(bv, -1, p) % p])
    helpers.check_script(h2e, oracle, field, sb.words, inputs, runner=helpers.run_gpu)


@pytest.mark.parametrize("field", [0, 1])
def test_misc_ops_gpu(h2e, oracle, field):
    p = oracle.FIELD_MODULUS[field]
    r = oracle.MODULI["bn256_fr"]
    rng = random.Random(300 + field)
    sb = h2e.ScriptBuilder()
    a, b = sb.assign_w(0), sb.assign_w(1)
    k = sb.assign_int_constant(1, 0)
    kin = sb.assign_int_constant(0, 2)
    m3 = sb.mul_small_const(a, 3)
    m2 = sb.mul_small_const(b, 2)
    ce = sb.is_int_equal(a, b)
    ce2 = sb.is_int_equal(a, a)
    bi = sb.bisec_int(ce2, m3, m2)
    sb.bisec_int(ce, k, kin)
    inv = sb.int_unsafe_invert(b)
    sb.int_mul(sb.int_square(bi), inv)
    n1 = sb.int_neg(sb.int_add(sb.int_sub(a, b), m3))
    sb.int_mul(n1, n1)
    x, y = sb.assign(3), sb.assign(4)
    bit0, bit1 = sb.assign_bit(5), sb.assign_bit(6)
    cst = sb.assign_constant(1, 1)
    cin = sb.assign_constant(0, 3)
    m = sb.mul(sb.add(x, y), sb.sub(x, y))
    sb.is_zero(m)
    sb.is_zero(sb.sub(x, x))
    for f in (sb.and_, sb.or_, sb.xor, sb.xnor, sb.not_and):
        f(bit0, bit1)
    sb.bisec(sb.not_(bit0), x, cst)
    sb.bisec(bit1, cin, y)
    statics = [rng.randrange(p), rng.randrange(r)]
    inputs = [[rng.randrange(p) if i else 0, rng.randrange(1, p), rng.randrange(p), rng.randrange(r), rng.randrange(r),
               rng.randrange(2), rng.randrange(2)] for i in range(40)]
    helpers.check_script(h2e, oracle, field, sb.words, inputs, statics, runner=helpers.run_gpu)


def test_status_bits_gpu(h2e, oracle):
    p = oracle.FIELD_MODULUS[0]
    sb = h2e.ScriptBuilder()
    a, b = sb.assign_w(0), sb.assign_w(1)
    sb.assert_int_equal(a, b)
    inputs = [[5, 5], [5, 6], [p - 1, p - 1], [0, p - 1]]
    shape = h2e.Shape.from_script(0, sb.words)
    _, status = helpers.run_gpu(shape, h2e.pack_inputs(inputs))
    assert list(status) == [0, h2e.ST_ASSERT_VALUE, 0, h2e.ST_ASSERT_VALUE]


def test_full_size_batch_properties(h2e, oracle):
    """BASELINE config 2 at full size (2^20 ops would be 4 GiB of records; 2^17 here keeps the test
    in seconds) checked through size-independent properties: the remainder limbs re-assemble to
    a*b mod w for every instance, every range chunk is < 2^18 and re-assembles to its limb, and a
    sampled subset is bit-exact with the oracle."""
    import torch

    field, L = 0, 3
    p = oracle.FIELD_MODULUS[field]
    n = 1 << 17
    rng = np.random.default_rng(5)
    sb = _int_mul_script(h2e, L, 1, 1, False)
    shape = h2e.Shape.from_script(field, sb.words)
    raw = rng.integers(0, 1 << 62, size=(n, 2, 5), dtype=np.uint64)
    vals_a = [int(sum(int(raw[i, 0, k]) << (62 * k) for k in range(5))) % p for i in range(n)]
    vals_b = [int(sum(int(raw[i, 1, k]) << (62 * k) for k in range(5))) % p for i in range(n)]
    mask = (1 << 108) - 1
    inputs = np.zeros((n, shape.n_input_cells, 32), dtype=np.uint8)
    for i in range(n):
        for k in range(L):
            inputs[i, 2 * k, :16] = np.frombuffer(((vals_a[i] >> (108 * k)) & mask).to_bytes(16, "little"), dtype=np.uint8)
            inputs[i, 2 * (L + k), :16] = np.frombuffer(((vals_b[i] >> (108 * k)) & mask).to_bytes(16, "little"), dtype=np.uint8)
    vals, status = shape.run(torch.from_numpy(inputs).cuda())
    torch.cuda.synchronize()
    assert int(status.abs().max()) == 0
    v = vals.cpu().numpy()  # [tiles, slots, 32, 32]
    base = 2 * (L + 1)  # int_mul cells start after the two load_int preludes
    # rem limb acc cells: slots base+6, base+13 (3-line limbs), base+18 (2-line leading)
    def cell_int(slot):
        x = v[:, slot].reshape(-1, 32)[:n]
        return [int.from_bytes(x[i].tobytes(), "little") for i in range(0, n, 97)]
    l0, l1, l2 = cell_int(base + 6), cell_int(base + 13), cell_int(base + 18)
    for j, i in enumerate(range(0, n, 97)):
        assert l0[j] + (l1[j] << 108) + (l2[j] << 216) == vals_a[i] * vals_b[i] % p
    # every chunk cell of the first limb < 2^18 and re-assembles
    chunks = v[:, base : base + 6].astype(np.uint64)
    lo = chunks[..., 0] | (chunks[..., 1] << 8) | (chunks[..., 2] << 16)
    assert (chunks[..., 3:] == 0).all() and (lo < (1 << 18)).all()
    # sampled bit-exact check vs oracle
    cells = None
    for i in [0, 1, 31, 32, n // 2, n - 1]:
        li = [(vals_a[i] >> (108 * k)) & mask for k in range(L)] + [(vals_b[i] >> (108 * k)) & mask for k in range(L)]
        rec = oracle.run_script(field, sb.words, li)
        if cells is None:
            cells = helpers.compare_static(shape, rec)
        helpers.compare_instance(shape, cells, v, i, rec)


def test_no_cpu_fallback(h2e):
    """The value path must be the CUDA library: it is loaded and counts its own launches."""
    before = h2e.lib().h2e_launch_count()
    sb = h2e.ScriptBuilder()
    sb.assign_w(0)
    shape = h2e.Shape.from_script(0, sb.words)
    helpers.run_gpu(shape, h2e.pack_inputs([[1], [2]]))
    assert h2e.lib().h2e_launch_count() == before + 2  # the VM (COMPACT records) + the expansion to 32-byte cells


# ---------------------------------------------------------------------------------------------
# whole circuits (BASELINE configs 1, 3, 4, 5 at oracle-sized parameters)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,n,n_inst", [(0, 7, 3), (0, 1, 33), (1, 3, 2), (4, 2, 2)])
def test_msm_circuits_gpu(h2e, oracle, kind, n, n_inst):
    import circuits_util as cu
    import ecmath as em
    from test_circuits_cpu import check_circuit

    C = em.BLS12_381 if kind == 4 else em.BN256
    inputs = [cu.msm_inputs(C, n, 777 + i) for i in range(n_inst)]
    if n_inst <= 3:
        check_circuit(h2e, oracle, kind, [n], inputs, runner=helpers.run_gpu)
    else:
        _check_subset(h2e, oracle, kind, [n], inputs, [0, 31, 32])


def _check_subset(h2e, oracle, kind, params, inputs, which):
    """Run all instances on the GPU, compare the listed ones with the oracle, and require status 0
    (i.e. the in-circuit self-check passed) for every instance."""
    shape = h2e.Shape.build(kind, params)
    vals, status = helpers.run_gpu(shape, h2e.pack_inputs(inputs))
    assert (status == 0).all(), status
    cells = None
    for i in which:
        rec = oracle.run_circuit(kind, params, inputs[i])
        assert rec.status == 0, rec.error
        if cells is None:
            cells = helpers.compare_static(shape, rec)
        helpers.compare_instance(shape, cells, vals, i, rec)


def test_bn256_check_pairing_gpu(h2e, oracle):
    """BASELINE config 4 shape (test_bn256_pairing_chip_over_bn256_fr): 3 instances on the GPU, two
    compared cell-by-cell with the oracle; status 0 = the pairing product is one in-circuit."""
    import circuits_util as cu

    inputs = [cu.bn_check_pairing_inputs(1000003 + i, 2000003 + 5 * i) for i in range(3)]
    _check_subset(h2e, oracle, 2, [], inputs, [0, 2])


def test_bls12_381_check_pairing_gpu(h2e, oracle):
    """BASELINE config 5 shape (test_bls12_381_pairing_chip_over_bn256_fr)."""
    import circuits_util as cu

    inputs = [cu.bls_check_pairing_inputs(424242 + i, 171717 + i, 99999999999 + i) for i in range(2)]
    _check_subset(h2e, oracle, 3, [], inputs, [1])


def test_pairing_rejects_bad_witness_gpu(h2e):
    """A pair that is not (a, -a) must raise the instance's status (the reference would panic in
    assert_constant, base_chip.rs:375-379) and leave the good instance untouched."""
    import circuits_util as cu
    import ecmath as em

    good = cu.bn_check_pairing_inputs(5, 7)
    bad = list(good)
    a2 = em.BN256.mul(em.BN256.g1, 6, 1)
    bad[7], bad[8] = a2
    shape = h2e.Shape.build(2, [])
    _, status = helpers.run_gpu(shape, h2e.pack_inputs([good, bad]))
    assert status[0] == 0 and status[1] & h2e.ST_ASSERT_VALUE


def _pairing_rows(n, seed):
    """n distinct bn256 check_pairing input rows (a_i = a_0 + i*G1, b_i = b_0 + i*G2)."""
    import circuits_util as cu
    import ecmath as em

    C = em.BN256
    a, b = C.mul(C.g1, 31337 + seed, 1), C.mul(C.g2, 271828 + seed, 2)
    rows = []
    for _ in range(n):
        na = C.neg(a, 1)
        rows.append(cu.g2_flat(b) + [na[0], na[1], 0, a[0], a[1], 0])
        a, b = C.add(a, C.g1, 1), C.add(b, C.g2, 2)
    return rows


@pytest.mark.parametrize("mode,cluster", [(2, 1), (2, 2), (2, 4), (2 | (3 << 8), 8), (1, 0)])
def test_pairing_execution_modes_agree_with_oracle(h2e, oracle, mode, cluster):
    """Every execution mode (thread per instance; team mode with 1/2/4/8 CTAs per tile and different
    critical/tail warp splits) must produce the same bit-exact records, on more than one tile and
    with a ragged last tile."""
    rows = _pairing_rows(40, mode * 10 + cluster)
    shape = h2e.Shape.build(2, [])
    shape.set_mode(mode, cluster)
    vals, status = helpers.run_gpu(shape, h2e.pack_inputs(rows))
    assert (status == 0).all(), status
    cells = None
    for i in (0, 33, 39):
        rec = oracle.run_circuit(2, [], rows[i])
        assert rec.status == 0, rec.error
        if cells is None:
            cells = helpers.compare_static(shape, rec)
        helpers.compare_instance(shape, cells, vals, i, rec)


def _oracle_parallel(oracle, kind, params, rows, threads=8):
    """oracle.run_circuit for several instances at once (the C++ oracle releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(lambda r: oracle.run_circuit(kind, params, r), rows))


def _compare_lane(shape, cells, tile_vals, lane, rec, tag):
    v = tile_vals[:, lane, :].cpu().numpy()
    for reg in range(3):
        m = cells[:, 0] == reg
        assert np.array_equal(v[m], rec.adv[reg][cells[m, 2], cells[m, 1]]), (tag, reg)


def test_bn256_pairing_full_batch(h2e, oracle):
    """BASELINE config 4 at its full size: 1024 instances (two resident passes of 512). Every
    instance must report status 0 -- the circuit itself asserts that the pairing product is one -- and
    16 instances at random positions of both passes are compared cell by cell with the oracle."""
    import random

    import torch

    rows = _pairing_rows(1024, 5)
    shape = h2e.Shape.build(2, [])
    packed = h2e.pack_inputs(rows)
    cells = shape.slot_cells()
    rng = random.Random(1024)
    vals = st = None
    for half in range(2):
        d_in = torch.from_numpy(packed[512 * half: 512 * (half + 1)]).cuda()
        vals, st = shape.run(d_in, vals, st)
        torch.cuda.synchronize()
        assert int(st.abs().max()) == 0
        pick = sorted(set([0, 511] + [rng.randrange(512) for _ in range(6)]))
        assert len(pick) >= 7
        recs = _oracle_parallel(oracle, 2, [], [rows[512 * half + i] for i in pick])
        for i, rec in zip(pick, recs):
            assert rec.status == 0, rec.error
            _compare_lane(shape, cells, vals[i // 32], i % 32, rec, (half, i))


def test_bls12_381_pairing_several_instances(h2e, oracle):
    """BASELINE config 5 shape on two tiles (one ragged): every instance status 0, five instances spread over both tiles
    compared cell by cell with the oracle."""
    import circuits_util as cu

    inputs = [cu.bls_check_pairing_inputs(31415 + 17 * i, 27182 + 3 * i, 1234567890123 + i) for i in range(36)]
    shape = h2e.Shape.build(3, [])
    vals, status = helpers.run_gpu(shape, h2e.pack_inputs(inputs))
    assert (status == 0).all(), status
    pick = [0, 13, 31, 32, 35]
    recs = _oracle_parallel(oracle, 3, [], [inputs[i] for i in pick], threads=5)
    cells = helpers.compare_static(shape, recs[0])
    for i, rec in zip(pick, recs):
        assert rec.status == 0, rec.error
        helpers.compare_instance(shape, cells, vals, i, rec)


@pytest.mark.parametrize("n", [11, 15])
def test_msm_odd_group_count_gpu(h2e, oracle, n):
    """msm_batch_on_group with an odd number (3) of point groups: the unpaired last group takes the branch at
    ecc_chip.rs:355-362. n = 11 (groups of 5, 5, 1) and n = 15 (5, 5, 5)."""
    import circuits_util as cu
    import ecmath as em
    from test_circuits_cpu import check_circuit

    inputs = [cu.msm_inputs(em.BN256, n, 4242 + i) for i in range(2)]
    check_circuit(h2e, oracle, 0, [n], inputs, runner=helpers.run_gpu)


@pytest.mark.slow
def test_msm_config1_full_size_bit_exact(h2e, oracle):
    """BASELINE configs[0] at its size: the reference's own MSM test (1000 points, select chip;
    src/tests/native_scalar_ecc_chip.rs:13-61). One full instance -- 37.09 M advice cells, 15.6 M permutation pairs,
    every fixed cell -- is compared bit for bit with the oracle; a second instance beside it must report status 0."""
    import torch

    import circuits_util as cu
    import ecmath as em

    C = em.BN256
    n = 1000
    # points a_i * G by repeated addition (cheap), scalars from the seeded stream, result from the known discrete logs
    g = em.scalar_stream(20240601, C.r)
    a0 = next(g) or 1
    P, pts, rows = C.mul(C.g1, a0, 1), [], []
    for _ in range(n):
        pts += [P[0], P[1], 0]
        P = C.add(P, C.g1, 1)
    r1, r2 = C.mul(C.g1, next(g) or 1, 1), C.mul(C.g1, next(g) or 1, 1)
    for _ in range(2):
        sc = [next(g) for _ in range(n)]
        acc = C.mul(C.g1, sum(b * (a0 + j) for j, b in enumerate(sc)) % C.r, 1)
        rows.append(pts + sc + [r1[0], r1[1], r2[0], r2[1]] + [acc[0], acc[1], 0])
    shape = h2e.Shape.build(h2e.CIRCUIT_MSM_BN256_SELECT, [n])
    assert (shape.base_offset, shape.range_offset, shape.select_offset) == (6292311, 6634908, 457600)
    vals, status = shape.run(torch.from_numpy(h2e.pack_inputs(rows)).cuda())
    torch.cuda.synchronize()
    assert int(status[:2].abs().max()) == 0
    rec = oracle.run_circuit(0, [n], rows[1])
    assert rec.status == 0, rec.error
    cells = helpers.compare_static(shape, rec)
    assert rec.n_adv == shape.n_slots
    helpers.compare_instance(shape, cells, {0: vals[0][:, 1:2, :].cpu().numpy()}, 0, rec)


def test_montgomery_export(h2e, oracle):
    """H2E_EXPORT_MONTGOMERY: every cell becomes x * 2^256 mod r (halo2's in-memory Fr), both through
    the device-side conversion and through the host entry point."""
    import torch

    p = oracle.FIELD_MODULUS[0]
    rng = random.Random(5)
    sb = _int_mul_script(h2e, 3, 1, 1, False)
    inputs = []
    for _ in range(37):
        a, b = rng.randrange(p), rng.randrange(p)
        inputs.append([(a >> (108 * k)) & ((1 << 108) - 1) for k in range(3)] + [(b >> (108 * k)) & ((1 << 108) - 1) for k in range(3)])
    shape = h2e.Shape.from_script(0, sb.words)
    packed = h2e.pack_inputs(inputs)
    canon, status = helpers.run_gpu(shape, packed)
    assert (status == 0).all()
    r = h2e.FR_MODULUS

    def to_mont(cells):  # uint8 [..., 32] -> same shape
        flat = cells.reshape(-1, 32)
        out = np.zeros_like(flat)
        for i in range(flat.shape[0]):
            x = int.from_bytes(flat[i].tobytes(), "little")
            out[i] = np.frombuffer(((x << 256) % r).to_bytes(32, "little"), dtype=np.uint8)
        return out.reshape(cells.shape)

    want = to_mont(canon[0][:, :5])  # lanes 0..4 of tile 0, every slot
    d_vals, _ = shape.run(torch.from_numpy(packed).cuda())
    shape.to_montgomery(d_vals)
    torch.cuda.synchronize()
    got = d_vals.cpu().numpy()
    assert np.array_equal(got[0][:, :5], want)
    assert np.array_equal(got[1][:, :5], to_mont(canon[1][:, :5]))
    shape.set_export(h2e.EXPORT_MONTGOMERY)
    host_vals, status = shape.run_host(packed)
    assert (status == 0).all() and np.array_equal(host_vals[0][:, :5], want)
    shape.set_export(h2e.EXPORT_CANONICAL)
    host_vals, _ = shape.run_host(packed)
    assert np.array_equal(host_vals[:, :, :5], canon[:, :, :5])


def test_msm_config3_size_one_tile(h2e):
    """BASELINE config 3 at its per-instance size: a 4096-point MSM (151 M advice cells = 4.83 GB per
    instance; one 32-instance tile fills the 180 GB of HBM). Size-independent property: the circuit
    itself asserts that the accumulated point equals the expected MSM result (computed on the host from
    the known discrete logs), so status 0 for every instance means 4096 x 254 selections, 234 k
    additions and the final equality all went through."""
    import torch

    sys_path_bench = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    import sys
    if sys_path_bench not in sys.path:
        sys.path.insert(0, sys_path_bench)
    import bench

    import gc

    gc.collect()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 80 * (1 << 30):
        pytest.skip("needs ~70 GB of free HBM")
    rows = bench._circuit_inputs("msm:4096", 32, seed=3)
    shape = h2e.Shape.build(0, [4096])
    assert shape.n_slots == 150993925
    # SURVEY Appendix B for this size (synthetic: exceeds the reference's MAX_ROWS): rows, permutation pairs, int_mul / int_div / reduce calls
    assert (shape.base_offset, shape.range_offset, shape.select_offset) == (25621431, 26995020, 1875920)
    assert shape.n_perms == 63587339
    ops = shape.program()[:, 0:2].copy().view(np.uint16).reshape(-1)
    assert (int((ops == 9).sum()), int((ops == 10).sum()), int((ops == 8).sum())) == (480913, 234180, 575568)
    d_in = torch.from_numpy(h2e.pack_inputs(rows)).cuda()
    # the VM's own record layout (COMPACT): 69 GB for the tile instead of 155 GB of 32-byte cells
    vals, st = shape.run_records(d_in, h2e.REC_COMPACT)
    torch.cuda.synchronize()
    assert int(st.abs().max()) == 0
    # size-independent property of the records themselves: every copy cell equals its root (the 63.6 M permutation pairs hold),
    # checked on the device for a sample of slots of every lane
    off, width, root = shape.layout(h2e.REC_COMPACT)
    rng = np.random.default_rng(1)
    copies = np.nonzero(root != np.arange(shape.n_slots))[0]
    pick = np.sort(rng.choice(copies, size=200000, replace=False))
    words = vals.view(torch.int32)

    def cell_words(slots, k):  # word k of the cells `slots` for all 32 lanes -> [len(slots), 32]
        w = torch.from_numpy(width[slots].astype(np.int64)).cuda()
        base = torch.from_numpy(off[slots].astype(np.int64)).cuda() * 32
        idx = base[:, None] + torch.arange(32, device="cuda")[None, :] * w[:, None] + k
        v = words[idx.clamp(max=words.numel() - 1)]
        return torch.where((w > k)[:, None], v, torch.zeros_like(v))

    for k in range(8):
        assert torch.equal(cell_words(pick, k), cell_words(root[pick], k)), f"word {k}: a copy cell differs from its root"
    # a wrong expected point must be caught
    bad = list(rows[0])
    bad[-3] ^= 1
    rows2 = [bad] + rows[1:]
    d_in = torch.from_numpy(h2e.pack_inputs(rows2)).cuda()
    vals, st = shape.run_records(d_in, h2e.REC_COMPACT, vals, st)
    torch.cuda.synchronize()
    s = st.cpu().numpy()
    assert s[0] != 0 and (s[1:32] == 0).all()


def test_unsafe_error_status_gpu(h2e, oracle):
    """UnsafeError::AddSameOrNegPoint on the GPU: flagged instance gets the code, its neighbour's
    records stay bit-exact (team mode, 2 instances in one tile)."""
    import circuits_util as cu
    import ecmath as em

    good = cu.msm_inputs(em.BN256, 1, 5)
    bad = list(good)
    bad[6], bad[7] = good[0], good[1]
    shape = h2e.Shape.build(0, [1])
    vals, status = helpers.run_gpu(shape, h2e.pack_inputs([bad, good]))
    assert status[0] & h2e.ST_ADD_SAME_OR_NEG and status[1] == 0
    rec = oracle.run_circuit(0, [1], good)
    cells = helpers.compare_static(shape, rec)
    helpers.compare_instance(shape, cells, vals, 1, rec)


def test_compact_export_is_lossless(h2e, oracle):
    """Compact export: static width classes probed on the device, packed records delivered to the host,
    host-side expansion == the plain host path, bit for bit (thread mode and team mode shapes)."""
    import circuits_util as cu

    p = oracle.FIELD_MODULUS[0]
    rng = random.Random(17)
    sb = _int_mul_script(h2e, 3, 16, 9, True)
    inputs = [_limbs(rng, 3, 16, 38) + _limbs(rng, 3, 9, 38) for _ in range(70)]
    shape = h2e.Shape.from_script(0, sb.words)
    packed = h2e.pack_inputs(inputs)
    full, st_full = shape.run_host(packed)
    compact, st_c = shape.run_host_compact(packed)
    w = shape.compact_widths()
    assert set(np.unique(w)) <= {1, 4, 8} and len(w) == shape.n_slots
    # the int_mul block (last 125 slots): 60 range chunks + v_h cells are 1 word wide, limbs 4, natives / sums 8
    blk = w[-125:]
    assert (blk == 1).sum() >= 60 and (blk == 4).sum() >= 30 and (blk == 8).sum() >= 10
    assert compact.nbytes == shape.compact_bytes(len(inputs)) == 3 * 32 * 4 * int(w.astype(np.int64).sum())
    assert compact.nbytes < 0.5 * full.nbytes
    assert np.array_equal(st_full, st_c)
    assert np.array_equal(shape.expand_compact(compact, len(inputs)), full)
    # a team-mode shape (bn256 MSM, one point)
    rows = [cu.msm_inputs(__import__("ecmath").BN256, 1, 900 + i) for i in range(3)]
    shape = h2e.Shape.build(0, [1])
    packed = h2e.pack_inputs(rows)
    full, st_full = shape.run_host(packed)
    compact, st_c = shape.run_host_compact(packed)
    assert (st_full == 0).all() and np.array_equal(st_full, st_c)
    assert np.array_equal(shape.expand_compact(compact, len(rows)), full)
    assert compact.nbytes < 0.5 * full.nbytes
