"""Shared test helpers: run a shape on the GPU (through the C ABI) or on the host emulator
(tests/emu, test infrastructure only), and compare everything with the oracle's records."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_SRC = os.path.join(HERE, "emu", "emu.cpp")
EMU_LIB = os.path.join(HERE, "emu", "libh2e_emu.so")
CSRC = os.path.join(HERE, "..", "halo2ecc-s_b200", "csrc")

_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        deps = [EMU_SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
        if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", EMU_LIB, EMU_SRC])
        L = ctypes.CDLL(EMU_LIB)
        L.emu_run.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                              ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _emu = L
    return _emu


_emu_probe = None


def emu_probe_widths(shape):
    """Width class (1, 4 or 8 words) of every slot as the macro-op CODE stores it: the emulator built with
    -DH2E_WIDTH_PROBE (every store records its width class instead of its value) runs the program once."""
    global _emu_probe
    lib_path = os.path.join(HERE, "emu", "libh2e_emu_probe.so")
    if _emu_probe is None:
        deps = [EMU_SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
        if not os.path.exists(lib_path) or any(os.path.getmtime(d) > os.path.getmtime(lib_path) for d in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DH2E_WIDTH_PROBE", "-o", lib_path, EMU_SRC])
        L = ctypes.CDLL(lib_path)
        L.emu_run.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                              ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _emu_probe = L
    cells = np.zeros((shape.n_slots, 8), dtype=np.uint32)
    inputs = np.zeros((1, max(shape.n_input_cells, 1), 32), dtype=np.uint8)
    status = np.zeros(1, dtype=np.uint32)
    prog, consts, tables = shape.program(), shape.consts(), shape.tables()
    _emu_probe.emu_run(prog.ctypes.data, prog.shape[0], consts.ctypes.data, tables.ctypes.data, shape.n_tables, shape.n_slots, shape.n_input_cells, 1,
                       inputs.ctypes.data, None, None, cells.ctypes.data, status.ctypes.data)
    return cells[:, 0].astype(np.uint8)


def run_emulated(shape, inputs_np, program=None):
    """inputs_np uint8 [n_inst, n_cells, 32] -> (vals uint8 [tiles, n_slots, 32, 32], status)"""
    n_inst = inputs_np.shape[0]
    inputs_np = np.ascontiguousarray(inputs_np[:, : shape.n_input_cells])
    tiles = (n_inst + 31) // 32
    vals = np.zeros((tiles, shape.n_slots, 32, 32), dtype=np.uint8)
    status = np.zeros(n_inst, dtype=np.uint32)
    prog = shape.program() if program is None else program
    consts = shape.consts()
    tables = shape.tables()
    off, width, _ = shape.layout(1)  # the VM's own record layout (COMPACT); the emulator expands it to plain cells
    emu_lib().emu_run(prog.ctypes.data, prog.shape[0], consts.ctypes.data, tables.ctypes.data, shape.n_tables, shape.n_slots, shape.n_input_cells, n_inst,
                      inputs_np.ctypes.data, off.ctypes.data, width.ctypes.data, vals.ctypes.data, status.ctypes.data)
    return vals, status


def run_gpu(shape, inputs_np, host_api=False):
    import torch

    if host_api:
        return shape.run_host(inputs_np)
    t = torch.from_numpy(np.ascontiguousarray(inputs_np)).cuda()
    vals, status = shape.run(t)
    torch.cuda.synchronize()
    return vals.cpu().numpy(), status.cpu().numpy()[: inputs_np.shape[0]].astype(np.uint32)


def compare_static(shape, rec):
    """Static half of the records (heights, offsets, which cells are set, permute flags, fixed cells
    that do not depend on the instance, permutation list) against one oracle run."""
    assert (shape.base_height, shape.range_height, shape.select_height) == (rec.base_height, rec.range_height, rec.select_height)
    assert (shape.base_offset, shape.range_offset, shape.select_offset) == (rec.base_offset, rec.range_offset, rec.select_offset)
    cells = shape.slot_cells()
    perms = shape.perms()
    assert perms.shape == rec.perms.shape, (perms.shape, rec.perms.shape)
    assert np.array_equal(perms, rec.perms), "permutation list differs"
    for reg in range(3):
        rows = rec.rows[reg]
        some = np.zeros((rows, rec.adv[reg].shape[1]), dtype=np.int32)
        m = cells[:, 0] == reg
        assert (cells[m, 2] < rows).all()
        np.add.at(some, (cells[m, 2], cells[m, 1]), 1)
        assert some.max(initial=0) <= 1, "a cell is assigned twice"
        assert np.array_equal(some.astype(np.uint8), rec.advf[reg] & 1), f"set of assigned advice cells differs in region {reg}"
        flag = np.zeros_like(some)
        for k in (0, 3):
            pm = perms[:, k] == reg
            flag[perms[pm, k + 2], perms[pm, k + 1]] = 1
        assert np.array_equal(flag.astype(np.uint8), (rec.advf[reg] >> 1) & 1), f"permute flags differ in region {reg}"
    return cells


def compare_instance(shape, cells, vals, inst, rec):
    """Advice values and fixed cells of one instance vs the oracle's records (bit-exact)."""
    tile, lane = divmod(inst, 32)
    v = vals[tile][:, lane, :]  # [n_slots, 32]
    for reg in range(3):
        m = cells[:, 0] == reg
        want = rec.adv[reg][cells[m, 2], cells[m, 1]]
        got = v[m]
        if not np.array_equal(got, want):
            bad = np.nonzero((got != want).any(axis=1))[0]
            i = int(bad[0])
            slot = int(np.nonzero(m)[0][i])
            raise AssertionError(
                f"instance {inst}: {len(bad)} advice cells differ in region {reg}; first slot {slot} cell "
                f"(col {cells[slot,1]}, row {cells[slot,2]}): got {int.from_bytes(got[i].tobytes(),'little'):#x} "
                f"want {int.from_bytes(want[i].tobytes(),'little'):#x}")
    fixed = shape.fixed()
    consts = shape.consts()
    for reg in range(3):
        rows = rec.rows[reg]
        fm = fixed[:, 0] == reg
        f = fixed[fm]
        some = np.zeros((rows, rec.fix[reg].shape[1]), dtype=np.uint8)
        fv = np.zeros((rows, rec.fix[reg].shape[1], 32), dtype=np.uint8)
        is_slot = (f[:, 3] & 0x80000000) != 0
        val = np.zeros((len(f), 32), dtype=np.uint8)
        val[~is_slot] = consts[f[~is_slot, 3]]
        val[is_slot] = v[f[is_slot, 3] & 0x7FFFFFFF]
        some[f[:, 2], f[:, 1]] = 1
        fv[f[:, 2], f[:, 1]] = val
        assert np.array_equal(some, rec.fixf[reg]), f"set of fixed cells differs in region {reg}"
        assert np.array_equal(fv, rec.fix[reg] * rec.fixf[reg][:, :, None]), f"fixed values differ in region {reg}"


def check_script(h2e, oracle, field, words, inputs_per_instance, statics=(), runner=run_emulated, expect_status=None):
    """Run one script on `runner` for every instance and compare each with the oracle."""
    shape = h2e.Shape.from_script(field, words, statics)
    packed = h2e.pack_inputs(inputs_per_instance)
    vals, status = runner(shape, packed)
    cells = None
    for i, inp in enumerate(inputs_per_instance):
        rec = oracle.run_script(field, words, inp, statics)
        if expect_status is None:
            assert rec.status == 0, rec.error
            assert rec.gate_ok, rec.gate_msg
            assert status[i] == 0, f"instance {i} status {status[i]}"
        if cells is None:
            cells = compare_static(shape, rec)
            assert rec.n_adv == shape.n_slots
        if expect_status is None:
            compare_instance(shape, cells, vals, i, rec)
        else:
            assert status[i] & expect_status[i] == expect_status[i], (i, status[i], expect_status[i])
    return shape
