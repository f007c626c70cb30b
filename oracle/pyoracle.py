"""ORACLE (test infrastructure, NOT product code).

ctypes binding of oracle/_ref/liboracle.so -- the CPU restatement of halo2ecc-s witness
generation. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "liboracle.so")

FIELD_BN256_FQ, FIELD_BLS12_381_FQ, FIELD_BLS12_381_FR = 0, 1, 2

MODULI = {
    "bn256_fr": 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    "bn256_fq": 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    "bls12_381_fq": 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    "bls12_381_fr": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
}
FIELD_MODULUS = {0: MODULI["bn256_fq"], 1: MODULI["bls12_381_fq"], 2: MODULI["bls12_381_fr"]}

# script opcodes (oracle/script.h)
OPS = dict(
    LOAD_INT=0, ASSIGN_W=1, ASSIGN_INT_CONSTANT=2, INT_ADD=3, INT_SUB=4, INT_NEG=5, INT_MUL=6, INT_SQUARE=7,
    INT_DIV=8, REDUCE=9, MUL_SMALL_CONST=10, BISEC_INT=11, IS_INT_ZERO=12, IS_INT_EQUAL=13,
    ASSERT_INT_EQUAL=14, INT_UNSAFE_INVERT=15, LOAD_INT_PACKED=16, ASSIGN=20, ASSIGN_CONSTANT=21, ASSIGN_BIT=22, AND=23, OR=24,
    NOT=25, XOR=26, XNOR=27, NOT_AND=28, BISEC=29, ADD=30, SUB=31, MUL=32, ASSERT_TRUE=34, ASSERT_FALSE=35,
    IS_ZERO=36, ASSERT_EQUAL=37,
)

ADV_COLS = {0: 5, 1: 3, 2: 2}
FIX_COLS = {0: 9, 1: 2, 2: 2}


def build(force=False):
    if force or not os.path.exists(_LIB) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
        for f in os.listdir(_HERE)
        if f.endswith((".h", ".cpp"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        L.orc_run_script.restype = ctypes.c_void_p
        L.orc_run_script.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_void_p, ctypes.c_size_t]
        L.orc_run_circuit.restype = ctypes.c_void_p
        L.orc_run_circuit.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.orc_run_circuit_result.restype = ctypes.c_void_p
        L.orc_run_circuit_result.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                             ctypes.c_void_p, ctypes.c_size_t]
        L.orc_pairing_constants.argtypes = [ctypes.c_void_p]
        L.orc_free.argtypes = [ctypes.c_void_p]
        L.orc_status.argtypes = [ctypes.c_void_p]
        L.orc_error.argtypes = [ctypes.c_void_p]
        L.orc_error.restype = ctypes.c_char_p
        L.orc_heights.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_export.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t] + [ctypes.c_void_p] * 4
        L.orc_perms.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_gate_check.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        L.orc_count_adv.argtypes = [ctypes.c_void_p]
        L.orc_count_adv.restype = ctypes.c_uint64
        L.orc_bench_int_mul.restype = ctypes.c_double
        L.orc_bench_int_mul.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_void_p]
        L.orc_bench_circuit.restype = ctypes.c_double
        L.orc_bench_circuit.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p,
                                        ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        _lib = L
    return _lib


def pack64(values):
    """list of python ints -> uint8 array [n,64] little-endian"""
    out = np.zeros((len(values), 64), dtype=np.uint8)
    for i, v in enumerate(values):
        out[i] = np.frombuffer(int(v).to_bytes(64, "little"), dtype=np.uint8)
    return out


class Records:
    """Python view of one instance's records, as the reference's `Records` holds them
    (src/context.rs:241-301): per region advice values + (some, permute) flags, fixed values +
    some flags, heights, permutation list."""

    def __init__(self, handle):
        L = lib()
        self.status = L.orc_status(handle)
        self.error = L.orc_error(handle).decode()
        h = np.zeros(7, dtype=np.uint64)
        L.orc_heights(handle, h.ctypes.data)
        self.base_height, self.range_height, self.select_height, nperm = (int(x) for x in h[:4])
        self.base_offset, self.range_offset, self.select_offset = (int(x) for x in h[4:7])
        self.rows = {0: self.base_height, 1: self.range_height + 1, 2: self.select_height + 1}
        self.adv, self.advf, self.fix, self.fixf = {}, {}, {}, {}
        for reg in range(3):
            rows = self.rows[reg]
            a = np.zeros((rows, ADV_COLS[reg], 32), dtype=np.uint8)
            af = np.zeros((rows, ADV_COLS[reg]), dtype=np.uint8)
            f = np.zeros((rows, FIX_COLS[reg], 32), dtype=np.uint8)
            ff = np.zeros((rows, FIX_COLS[reg]), dtype=np.uint8)
            L.orc_export(handle, reg, rows, a.ctypes.data, af.ctypes.data, f.ctypes.data, ff.ctypes.data)
            self.adv[reg], self.advf[reg], self.fix[reg], self.fixf[reg] = a, af, f, ff
        self.perms = np.zeros((nperm, 6), dtype=np.uint32)
        if nperm:
            L.orc_perms(handle, self.perms.ctypes.data)
        buf = ctypes.create_string_buffer(512)
        self.gate_ok = L.orc_gate_check(handle, buf, 512) == 0
        self.gate_msg = buf.value.decode()
        self.n_adv = int(L.orc_count_adv(handle))

    def adv_int(self, reg, row, col):
        return int.from_bytes(self.adv[reg][row, col].tobytes(), "little")


def run_script(field, script, inputs, statics=()):
    L = lib()
    s = np.asarray(script, dtype=np.uint32)
    i = pack64(inputs)
    st = pack64(statics)
    h = L.orc_run_script(field, s.ctypes.data, len(s), i.ctypes.data, len(inputs), st.ctypes.data, len(statics))
    try:
        return Records(h)
    finally:
        L.orc_free(h)


def run_circuit(kind, params, inputs):
    L = lib()
    p = np.asarray(params, dtype=np.uint64)
    i = pack64(inputs)
    h = L.orc_run_circuit(kind, p.ctypes.data, len(p), i.ctypes.data, len(inputs))
    try:
        return Records(h)
    finally:
        L.orc_free(h)


def run_circuit_result(kind, params, inputs, n_result=12):
    """kinds 5/6: returns (Records, [12 Fq coefficients of the pairing value])"""
    L = lib()
    p = np.asarray(params, dtype=np.uint64)
    i = pack64(inputs)
    out = np.zeros((n_result, 64), dtype=np.uint8)
    h = L.orc_run_circuit_result(kind, p.ctypes.data, len(p), i.ctypes.data, len(inputs), out.ctypes.data, n_result)
    try:
        return Records(h), [int.from_bytes(out[k].tobytes(), "little") for k in range(n_result)]
    finally:
        L.orc_free(h)


def pairing_constants():
    """dict of the Frobenius/twist constants the oracle derives, as (c0, c1) int pairs"""
    out = np.zeros((30, 2, 64), dtype=np.uint8)
    lib().orc_pairing_constants(out.ctypes.data)
    v = [(int.from_bytes(out[i, 0].tobytes(), "little"), int.from_bytes(out[i, 1].tobytes(), "little")) for i in range(30)]
    return dict(fq2_c1=v[0:2], fq6_c1=v[2:8], fq6_c2=v[8:14], fq12_c1=v[14:26], xi_to_q_minus_1_over_2=v[26],
                bls_fq6_c1=v[27], bls_fq6_c2=v[28], bls_fq12_c1=v[29])


def bench_int_mul(field, limbs_a_b, times, threads):
    """limbs_a_b: python ints, n*2*L; times: n*2. Returns (seconds, cells)."""
    L = lib()
    n = len(times) // 2
    i = pack64(limbs_a_b)
    t = np.asarray(times, dtype=np.uint32)
    cells = ctypes.c_uint64(0)
    sec = L.orc_bench_int_mul(field, n, i.ctypes.data, t.ctypes.data, threads, ctypes.byref(cells))
    return sec, cells.value


def bench_int_mul_packed(field, packed, times, threads):
    """Same with the operands already packed: uint8 [n, 2*L, 64] (one 64-byte little-endian value per limb -- the
    product's own input layout) and times uint32 [n, 2]. Returns (seconds, cells)."""
    L = lib()
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    t = np.ascontiguousarray(times, dtype=np.uint32).reshape(-1)
    n = t.shape[0] // 2
    assert packed.size == n * (packed.size // max(n, 1)) and packed.size % 64 == 0
    cells = ctypes.c_uint64(0)
    sec = L.orc_bench_int_mul(field, n, packed.ctypes.data, t.ctypes.data, threads, ctypes.byref(cells))
    return sec, cells.value


def bench_circuit(kind, params, n, inputs_packed, n_inputs_per, threads):
    L = lib()
    p = np.asarray(params, dtype=np.uint64)
    cells = ctypes.c_uint64(0)
    sec = L.orc_bench_circuit(kind, p.ctypes.data, len(p), n, inputs_packed.ctypes.data, n_inputs_per, threads,
                              ctypes.byref(cells))
    return sec, cells.value


class ScriptBuilder:
    """Assemble op scripts; returns indices into the int / val result lists."""

    def __init__(self):
        self.words = []
        self.n_int = 0
        self.n_val = 0

    def _emit(self, op, *args):
        self.words += [OPS[op], len(args)] + [int(a) for a in args]

    def _int(self):
        self.n_int += 1
        return self.n_int - 1

    def _val(self):
        self.n_val += 1
        return self.n_val - 1

    def load_int(self, times, in_idx): self._emit("LOAD_INT", times, in_idx); return self._int()
    def load_int_packed(self, times, in_idx): self._emit("LOAD_INT_PACKED", times, in_idx); return self._int()  # L limbs of 16 bytes in one logical input
    def assign_w(self, in_idx): self._emit("ASSIGN_W", in_idx); return self._int()
    def assign_int_constant(self, src, idx): self._emit("ASSIGN_INT_CONSTANT", src, idx); return self._int()
    def int_add(self, a, b): self._emit("INT_ADD", a, b); return self._int()
    def int_sub(self, a, b): self._emit("INT_SUB", a, b); return self._int()
    def int_neg(self, a): self._emit("INT_NEG", a); return self._int()
    def int_mul(self, a, b): self._emit("INT_MUL", a, b); return self._int()
    def int_square(self, a): self._emit("INT_SQUARE", a); return self._int()
    def int_div(self, a, b): self._emit("INT_DIV", a, b); return self._val(), self._int()
    def reduce(self, a): self._emit("REDUCE", a); return self._int()
    def mul_small_const(self, a, k): self._emit("MUL_SMALL_CONST", a, k); return self._int()
    def bisec_int(self, c, a, b): self._emit("BISEC_INT", c, a, b); return self._int()
    def is_int_zero(self, a): self._emit("IS_INT_ZERO", a); return self._val()
    def is_int_equal(self, a, b): self._emit("IS_INT_EQUAL", a, b); return self._val()
    def assert_int_equal(self, a, b): self._emit("ASSERT_INT_EQUAL", a, b)
    def int_unsafe_invert(self, a): self._emit("INT_UNSAFE_INVERT", a); return self._int()
    def assign(self, in_idx): self._emit("ASSIGN", in_idx); return self._val()
    def assign_constant(self, src, idx): self._emit("ASSIGN_CONSTANT", src, idx); return self._val()
    def assign_bit(self, in_idx): self._emit("ASSIGN_BIT", in_idx); return self._val()
    def and_(self, a, b): self._emit("AND", a, b); return self._val()
    def or_(self, a, b): self._emit("OR", a, b); return self._val()
    def not_(self, a): self._emit("NOT", a); return self._val()
    def xor(self, a, b): self._emit("XOR", a, b); return self._val()
    def xnor(self, a, b): self._emit("XNOR", a, b); return self._val()
    def not_and(self, a, b): self._emit("NOT_AND", a, b); return self._val()
    def bisec(self, c, a, b): self._emit("BISEC", c, a, b); return self._val()
    def add(self, a, b): self._emit("ADD", a, b); return self._val()
    def sub(self, a, b): self._emit("SUB", a, b); return self._val()
    def mul(self, a, b): self._emit("MUL", a, b); return self._val()
    def assert_true(self, a): self._emit("ASSERT_TRUE", a)
    def assert_false(self, a): self._emit("ASSERT_FALSE", a)
    def is_zero(self, a): self._emit("IS_ZERO", a); return self._val()
    def assert_equal(self, a, b): self._emit("ASSERT_EQUAL", a, b)
