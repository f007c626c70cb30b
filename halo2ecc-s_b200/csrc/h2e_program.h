// Program format shared by the host tracer and the CUDA witness VM.
//
// A circuit *shape* (everything in halo2ecc-s's Records that does not depend on witness values:
// row offsets, fixed cells, tags, encodes, permutations) is traced once on the host. The value
// side is a straight-line program of macro-ops, one per chip call of the reference
// (IntegerChipOps / BaseChipOps / SelectChipOps), which the GPU evaluates for every instance.
//
// Advice cells are numbered densely in assignment order ("slots"); a macro-op writes a contiguous
// slot range starting at Instr::out. Values live in HBM tile-interleaved over 32 instances:
//   vals[tile][slot][lane 0..31][8 x u32]      (one 32-byte canonical little-endian Fr per cell)
// so that a warp (lane = instance) stores 1 KiB contiguous per cell.
#pragma once
#include <stdint.h>

namespace h2e {

enum Field : uint32_t { F_BN256_FQ = 0, F_BLS12_381_FQ = 1, F_BLS12_381_FR = 2, F_COUNT = 3 };

static const int TILE = 32;        // instances per tile
static const int LIMB_BITS = 108;  // RangeInfo::limb_bits (range_info.rs:98)
static const int MAX_L = 4;
static const uint32_t NONE = 0xffffffffu;

// status bits (per instance). Low codes mirror UnsafeError (ecc_chip.rs:23-28); the rest are the
// reference's panics (assert_eq!/unwrap) surfaced as flags.
enum Status : uint32_t {
    ST_ADD_SAME_OR_NEG = 1u << 0,
    ST_ADD_IDENTITY = 1u << 1,
    ST_ASSIGN_IDENTITY = 1u << 2,
    ST_ASSERT_VALUE = 1u << 4,        // assert_constant / assert_true / assert_false value mismatch
    ST_NONZERO_REMAINDER = 1u << 5,   // integer_chip.rs:120,148,178,346
    ST_NEGATIVE = 1u << 6,            // BigUint underflow in the reference
    ST_RANGE = 1u << 7,               // value does not fit the range rows it is assigned to
};

enum Op : uint16_t {
    OP_NOP = 0,
    // ---- integer chip (integer_chip.rs) ----
    OP_LOAD_INT,        // a0=input cell (L limb values), a1=packed   -> L+1 `assign` rows (harness prelude)
    OP_ASSIGN_W,        // a0=input cell | const-pool idx, a1=src (0 input, 1 const pool) -> assign_w cells
    OP_ASSIGN_INT_CONST,  // a0=src(0 input,1 const pool), a1=idx      -> L+1 assign_constant rows
    OP_INT_ADD,         // a[0..L)=a limbs, a[L..2L)=b limbs, a[2L]=a.native, a[2L+1]=b.native
    OP_INT_SUB,         // a[0..L)=a limbs, a[L..2L)=b limbs, a[2L]=b.times, a[2L+1]=a.native, a[2L+2]=b.native
    OP_INT_NEG,         // a[0..L)=limbs, a[L]=a.times, a[L+1]=a.native
    OP_MUL_SMALL,       // a[0..L)=limbs, a[L]=k, a[L+1]=a.native
    OP_REDUCE,          // a[0..L)=limbs, a[L]=native
    OP_INT_MUL,         // a[0..L]=a limbs+native, a[L+1..2L+1]=b limbs+native
    OP_DIV_CORE,        // a[0..L]=masked numerator limbs+native, a[L+1..2L+1]=denominator limbs+native
    OP_IS_INT_ZERO,     // a[0..L]=limbs+native (times==1)            -> cond is the last cell
    OP_MASK_INT,        // a[0..L]=limbs+native, a[L+1]=cond          -> (L+1) mul rows
    OP_BISEC_INT,       // a0=cond, a[1..L+1]=a limbs+native, a[L+2..2L+2]=b limbs+native
    OP_SUM_ASSERT_ZERO, // a[0..L)=limbs                              -> sum row + assert_constant(sum,0) row
    // ---- base chip (base_chip.rs) ----
    OP_ASSIGN,          // a0=input idx
    OP_ASSIGN_CONST,    // a0=src, a1=idx
    OP_ASSIGN_BIT,      // a0=input idx
    OP_LINSUM,          // a0=n, a1=const idx|NONE, then n x (slot, coeff const idx)  sum_with_constant_in_one_line
    OP_MUL,             // a0,a1                                       [a,b,ab]
    OP_BOOL,            // a0,a1,a2=kind (0 and handled by OP_MUL; 1 or, 2 xor, 3 xnor, 4 not_and)
    OP_BISEC,           // a0=cond,a1,a2
    OP_IS_ZERO,         // a0                                          invert(): 2 rows, cond = last cell
    OP_ASSERT_CONST,    // a0=slot, a1=const idx                       value check + 1 cell
    OP_ASSERT_EQUAL,    // a0,a1                                       2 cells
    // ---- scalar / select (native_scalar_ecc_chip.rs, select_chip.rs, ecc_chip.rs) ----
    OP_DECOMPOSE_NATIVE,  // a0=scalar slot                             native_scalar_ecc_chip.rs:97-171
    OP_CACHE_INT,         // a[0..L]=limbs+native                       L+1 select rows (value col)
    OP_SELECT_INT,        // a0=index slot, a1=offset into Shape::tables, a2=#candidates  -> (L+1) x [value, selector]
    OP_DECOMPOSE_LIMB,    // a0=limb slot, a1=bits                      general_scalar_ecc_chip.rs:108-128
    // ---- scheduler-only split ops (team mode; never emitted by the tracer) ----
    OP_INT_MUL_HEAD,      // same operands as OP_INT_MUL: writes only rem / d limb acc cells + natives (what later ops read)
    OP_INT_MUL_TAIL,      // same operands: reads those cells back and writes the rest of the block
    OP_REDUCE_HEAD,       // same operands as OP_REDUCE: writes rem limb acc cells + native + the quotient cell
    OP_REDUCE_TAIL,       // same operands: reads those cells back and writes the whole block
    OP_DIV_INV,           // K = flags & 3 denominators (1..3; the scheduler merges the inversions of one dependency level): limbs of
                          // denominator j at a[j(L+1) ..], its scratch entry in a[j(L+1)+L]: b_j^-1 mod w -> scratch, ONE inversion
                          // for all K (Montgomery's trick); no record cells
    OP_DIV_CORE_S,        // OP_DIV_CORE with b^-1 read from scratch entry a[2L+2] (so the inversion runs beside is_int_zero)
    OP_IS_INT_ZERO_HEAD,  // same operands as OP_IS_INT_ZERO, a[13]=slot of the condition cell: writes only that cell (no inversion)
    OP_IS_INT_ZERO_TAIL,  // writes the whole block of flags & 3 is_int_zero calls with ONE Fr inversion (nothing waits for it):
                          // operands of block j at a[j(L+1) ..], first slot of block j >= 1 in a[11 + j]
    OP_DIV_HEAD_S,        // operands of OP_DIV_CORE_S: c = a * b^-1 mod w, writes only c's limb acc cells + native
    OP_DIV_TAIL,          // operands of OP_DIV_CORE: reads c back, computes d and writes the whole block
    // ---- vector forms of base-chip rows (keccak chip: 64 independent bit rows of a lane in ONE macro-op) ----
    OP_BOOLV,             // a0=kind (1 or, 2 xor, 3 xnor, 4 not_and), a1=n, a2=offset into Shape::tables of n x (a, b) slots:
                          // n rows [a_i, b_i] last(c_i) in order (3n cells) = n consecutive BaseChipOps::{or,xor,xnor,not_and}
    OP_CHIV,              // a1=n, a2=offset into Shape::tables of n x (u, v, s) slots: per element the two rows of keccak's chi
                          // (keccak_chip.rs:104-121): t = not_and(u, v) then xor(s, t) (6n cells)
    OP_COUNT
};

struct Instr {
    uint16_t op;
    uint8_t field;
    uint8_t flags;
    uint32_t out;    // first output slot
    uint32_t a[14];  // operands (slots / immediates)
};
static_assert(sizeof(Instr) == 64, "Instr must be 64 bytes");

// Per-field constants (RangeInfo, range_info.rs:14-54) in 32-bit words, plus Barrett / Montgomery
// precomputation for the modulus.
struct FieldConst {
    uint32_t L, M, R, P;        // limbs, mul_check_limbs, reduce_check_limbs, pure_w_check_limbs
    uint32_t nbits;             // w_ceil_bits
    uint32_t nw;                // words of w (8 or 12)
    uint32_t w_lead_bits;       // w_ceil_bits % 108
    uint32_t d_lead_bits;       // d_bits % 108
    uint32_t w[13];             // modulus
    uint32_t mu[14];            // floor(2^k / w), k = kbits
    uint32_t kbits;
    uint32_t w_limbs[4][4];     // w_modulus_limbs_le
    uint32_t w_native[8];       // w mod r
    uint32_t neg_w_native[8];   // -w_native mod r
    uint32_t neg_w_limbs[4][8]; // -w_limb mod r (for is_pure_w_modulus add_constant)
    uint32_t minv;              // -w^-1 mod 2^32 (Montgomery)
    uint32_t r2[12];            // 2^(2*32*nw) mod w
    uint32_t one_m[12];         // 2^(32*nw) mod w
    uint32_t wm2[12];           // w - 2 (Fermat exponent)
    uint32_t upper[64][4][4];   // w_modulus_of_ceil_times[t][limb] (range_info.rs:334-359), < 2^115
    uint32_t upper_native[64][8];  // sum_i upper[t][i] * 2^(108 i) mod r
};

struct FrConst {
    uint32_t r[8];
    uint32_t mu[9];   // floor(2^512 / r)
    uint32_t minv;    // -r^-1 mod 2^32
    uint32_t r2[8];   // 2^512 mod r
    uint32_t one_m[8];
    uint32_t rm2[8];  // r - 2
    uint32_t r2w1[8]; // 2^(256 + 32) mod r: Montgomery form of a 1-word cell in one CIOS round (mont_mul_short)
    uint32_t r2w4[8]; // 2^(256 + 128) mod r: of a 4-word cell in four rounds
};

struct DeviceConsts {
    FrConst fr;
    FieldConst f[F_COUNT];
};

}  // namespace h2e
