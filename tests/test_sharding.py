"""Multi-GPU host logic on CPU: instances shard contiguously across ranks with no data-path
collective; only the per-instance status words are gathered. world_size 2, gloo. The VM
itself needs a GPU, so each rank evaluates its shard on the host emulator (test infrastructure) and
the union is compared with the oracle."""
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition(h2e):
    for n in (0, 1, 31, 32, 33, 1000, 1024, 4096):
        for world in (1, 2, 3, 4, 8):
            ranges = [h2e.shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 32  # whole tiles, balanced to one tile
            assert all(a % 32 == 0 for a, _ in ranges)


def _worker(rank, world, port, n_inst, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import __graft_entry__ as ge
    import helpers

    h2e = ge.load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
    rng = random.Random(99)
    inputs = [[rng.randrange(p), rng.randrange(1, p)] for _ in range(n_inst)]
    inputs[40][1] = 0  # division by zero -> assert_false(is_zero) raises that instance's status
    sb = h2e.ScriptBuilder()
    a, b = sb.assign_w(0), sb.assign_w(1)
    sb.int_mul(sb.int_add(a, b), sb.int_unsafe_invert(b))
    shape = h2e.Shape.from_script(0, sb.words)
    lo, hi = h2e.shard_range(n_inst, world, rank)
    vals, status = helpers.run_emulated(shape, h2e.pack_inputs(inputs[lo:hi])) if hi > lo else (None, np.zeros(0, np.uint32))
    all_status = h2e.gather_status(status, n_inst, world, rank)
    # one cell of every local instance travels back for the cross-check in the parent
    sample = [] if vals is None else [bytes(vals[i // 32][shape.n_slots - 1, i % 32]) for i in range(hi - lo)]
    q.put((rank, lo, hi, all_status.tolist(), sample))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_matches_single_process(h2e, oracle):
    import torch.multiprocessing as mp

    import helpers

    n_inst, world, port = 70, 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_inst, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = sorted(q.get(timeout=300) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert [(g[1], g[2]) for g in got] == [(0, 32), (32, 70)]  # whole tiles per rank, ragged tile last
    assert got[0][3] == got[1][3] and len(got[0][3]) == n_inst  # every rank sees every status
    status = got[0][3]
    assert status[40] & h2e.ST_ASSERT_VALUE and all(s == 0 for i, s in enumerate(status) if i != 40)
    # single-process evaluation of the same batch gives the same cells
    p = oracle.FIELD_MODULUS[0]
    rng = random.Random(99)
    inputs = [[rng.randrange(p), rng.randrange(1, p)] for _ in range(n_inst)]
    inputs[40][1] = 0
    sb = h2e.ScriptBuilder()
    a, b = sb.assign_w(0), sb.assign_w(1)
    sb.int_mul(sb.int_add(a, b), sb.int_unsafe_invert(b))
    shape = h2e.Shape.from_script(0, sb.words)
    vals, _ = helpers.run_emulated(shape, h2e.pack_inputs(inputs))
    sample = got[0][4] + got[1][4]
    assert sample == [bytes(vals[i // 32][shape.n_slots - 1, i % 32]) for i in range(n_inst)]
