/* h2ecc_b200 -- C ABI of the B200 batched witness generator for halo2ecc-s circuits.
 *
 * The reference (DelphinusLab/halo2ecc-s v0.3.2) has no FFI: its hot path is a set of Rust traits
 * implemented on context structs, whose product is `Records` (src/context.rs:294-301) consumed by
 * `Records::assign_all` (src/context.rs:575-588). This header is the boundary a Rust shim binds
 * (see INTEGRATION.md): the shape-side calls replace the *structural* half of those traits (row
 * layout, fixed cells, permutations), the batch calls replace the *value* half for N instances.
 *
 * Plain pointers and sizes only. All values are canonical little-endian integers.
 * Return codes: 0 ok, <0 error (message via h2e_last_error()).
 */
#ifndef H2ECC_B200_H
#define H2ECC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct h2e_shape h2e_shape;

/* wrong fields W over N = bn256 Fr (IntegerContext<W,N>, src/context.rs:161-188) */
enum { H2E_FIELD_BN256_FQ = 0, H2E_FIELD_BLS12_381_FQ = 1, H2E_FIELD_BLS12_381_FR = 2 };

/* circuit kinds for h2e_shape_build (the shapes of the reference's own tests) */
enum {
    H2E_CIRCUIT_MSM_BN256_SELECT = 0,   /* src/tests/native_scalar_ecc_chip.rs:13-61, params: n_points */
    H2E_CIRCUIT_MSM_BN256_NOSELECT = 1, /* src/tests/native_scalar_ecc_chip.rs:63-110 */
    H2E_CIRCUIT_PAIRING_BN256 = 2,      /* src/tests/native_scalar_pairing_chip.rs:67-97 */
    H2E_CIRCUIT_PAIRING_BLS12_381 = 3,  /* src/tests/general_scalar_pairing_chip.rs:74-105 */
    H2E_CIRCUIT_MSM_BLS12_381 = 4       /* src/tests/general_scalar_ecc_chip.rs:14-49 */
};

/* per-instance status bits. 1/2/4 = UnsafeError::{AddSameOrNegPoint,AddIdentity,AssignIdentity}
 * (src/circuit/ecc_chip.rs:23-28); >= 16 = conditions on which the reference panics. Records of an
 * instance with non-zero status are unspecified. */
enum {
    H2E_ST_ADD_SAME_OR_NEG = 1,
    H2E_ST_ADD_IDENTITY = 2,
    H2E_ST_ASSIGN_IDENTITY = 4,
    H2E_ST_ASSERT_VALUE = 16,
    H2E_ST_NONZERO_REMAINDER = 32,
    H2E_ST_NEGATIVE = 64,
    H2E_ST_RANGE = 128
};

const char* h2e_last_error(void);
int h2e_version(void);

/* ---- shape side (host only; once per circuit shape) --------------------------------------- */

/* Build a shape by replaying chip calls given as an op-script: words = (opcode, nargs, args...)*;
 * opcodes in csrc/script_builder.h. Covered: BaseChipOps (src/circuit/base_chip.rs:81-501),
 * IntegerChipOps of the chosen field (src/circuit/integer_chip.rs:15-70) and, on the curve whose base
 * field that is (bn256 G1 / bls12_381 G1), EccChipBaseOps' safe-point API and msm_unsafe with explicit
 * blinding points (src/circuit/ecc_chip.rs:373-408, 438-812) and PairingChipOps::check_pairing with G2
 * points as per-instance constants (src/circuit/pairing_chip.rs:170-176). `statics64` are shape-level constants,
 * 64 bytes each. Replaces: constructing a Context and calling the trait methods. */
h2e_shape* h2e_shape_from_script(int field, const uint32_t* script, size_t n_words, const uint8_t* statics64, size_t n_statics);

/* Build the shape of one of the reference's test circuits. */
h2e_shape* h2e_shape_build(int circuit_kind, const uint64_t* params, size_t n_params);

void h2e_shape_free(h2e_shape* s);

/* out[0..2] base/range/select height (Records::{base,range,select}_height), out[3..5] final
 * base/range/select offsets (Context), out[6] advice cells ("slots"), out[7] fixed cells,
 * out[8] permutation pairs, out[9] program length, out[10] constants, out[11] per-instance input
 * cells (32 bytes each), out[12] slot-table words (select chip candidate tables). */
int h2e_shape_query(const h2e_shape* s, uint64_t out[16]);

/* slot -> advice cell, 3 x u32 (region, col, row) per slot. Regions: 0 base, 1 range, 2 select. */
int h2e_shape_slot_cells(const h2e_shape* s, uint32_t* out);
/* fixed cells, 4 x u32 (region, col, row, cidx) each. cidx < 2^31: index into the constant
 * table; cidx >= 2^31: the cell equals advice slot (cidx & 0x7fffffff) of the same instance
 * (assign_constant of a per-instance value, src/circuit/base_chip.rs:344-349). */
int h2e_shape_fixed(const h2e_shape* s, uint32_t* out);
/* constant table, 32 bytes each */
int h2e_shape_consts(const h2e_shape* s, uint8_t* out);
/* the value program: 64 bytes per macro-op (csrc/h2e_program.h), for inspection / tooling */
int h2e_shape_program(const h2e_shape* s, uint8_t* out);
/* The levelised program used by team mode: instructions sorted by dependency level (same 64-byte
 * format; every int_mul / reduce / is_int_zero is split into a HEAD and a deferred TAIL instruction, every int_div
 * core into the W inversion, a HEAD and a TAIL; up to three is_int_zero TAILs are merged into one), level l =
 * [level_start[l], level_start[l+1]). Any of the output pointers may be NULL; program_out needs
 * *n_instr * 64 bytes, level_start_out *n_levels + 1 entries. */
int h2e_shape_schedule(h2e_shape* s, uint64_t* n_levels, uint64_t* n_instr, uint8_t* program_out, uint32_t* level_start_out);
/* Team mode executes the levelised program as per-warp instruction streams with explicit
 * dependencies (dataflow; csrc/schedule.h). For tooling and tests: builds the streams for
 * `ctas_per_tile` CTAs per tile and returns the instructions in an order in which a host model of
 * that execution starts them (*n_instr entries of 64 bytes; fails if the streams would deadlock),
 * and/or the modelled makespan in cycles. Output pointers may be NULL. */
int h2e_shape_team_order(h2e_shape* s, int ctas_per_tile, uint64_t* n_instr, uint8_t* program_out, double* est_cycles);
/* slot tables referenced by the select-chip macro-ops (u32 each) */
int h2e_shape_tables(const h2e_shape* s, uint32_t* out);
/* Records::permutations, 6 x u32 (region, col, row) x 2 per pair, in the reference's order */
int h2e_shape_perms(const h2e_shape* s, uint32_t* out);

/* ---- value side (GPU) --------------------------------------------------------------------- */

/* bytes of the advice-value buffer for n_inst instances. Layout: tiles of 32 instances,
 * vals[tile][slot][lane][32 bytes]; instance i is (tile i/32, lane i%32). */
size_t h2e_vals_bytes(const h2e_shape* s, uint64_t n_inst);
/* bytes of the input buffer: inputs[instance][input cell][32 bytes] */
size_t h2e_inputs_bytes(const h2e_shape* s, uint64_t n_inst);

/* Fill the advice cells of n_inst instances. All pointers are DEVICE pointers on `device`;
 * `stream` is a cudaStream_t (NULL = default stream). Asynchronous. Replaces running the chip
 * calls once per instance on the CPU (e.g. src/circuit/integer_chip.rs:466-483 for int_mul). */
int h2e_batch_run(h2e_shape* s, int device, void* stream, uint64_t n_inst, const void* d_inputs, void* d_vals, uint32_t* d_status);

/* Same with HOST buffers: copies inputs to the device, runs, copies values and status back.
 * Values are produced in chunks of tiles and streamed out while the next chunk computes. */
int h2e_batch_run_host(h2e_shape* s, int device, uint64_t n_inst, const void* h_inputs, void* h_vals, uint32_t* h_status);

/* Cell encoding of the value buffers. The VM always computes canonical little-endian integers (the
 * bytes `field_to_bn` would see, src/utils.rs:4-9); H2E_EXPORT_MONTGOMERY rewrites every cell in place
 * as x * 2^256 mod r in four little-endian u64 limbs, i.e. the in-memory representation of halo2's
 * bn256 `Fr`, so that a Rust shim can reinterpret the buffer as `[Fr]` without a per-cell
 * `bn_to_field` (src/utils.rs:11-17). */
enum { H2E_EXPORT_CANONICAL = 0, H2E_EXPORT_MONTGOMERY = 1 };
/* Encoding produced by h2e_batch_run_host for this shape (default canonical). */
int h2e_shape_set_export(h2e_shape* s, int format);
/* In-place conversion of n_cells 32-byte canonical cells at DEVICE pointer d_cells (e.g. the buffer
 * filled by h2e_batch_run) to the Montgomery encoding. Asynchronous on `stream`. */
int h2e_cells_to_montgomery(h2e_shape* s, int device, void* stream, void* d_cells, uint64_t n_cells);

/* ---- compact export ------------------------------------------------------------------------
 * Most cells are narrow: of the 125 cells of an int_mul block 60 are 18-bit range chunks and 40 are
 * 108-bit limbs. Every slot has a static width class -- 1, 4 or 8 significant 32-bit words, fixed by the
 * chip call that assigns it -- and the compact form stores exactly those words:
 *   compact[tile][slot][lane 0..31][w(slot) words], slots back to back (slot s starts at word
 *   32 * sum_{t<s} w(t) of its tile's block).
 * It is lossless (the dropped words are zero) and ~2.7x smaller, which is what the PCIe / host-memory
 * bound host path moves. h2e_compact_prepare derives the widths once per shape (on the device, with a
 * build of the VM whose stores record widths instead of values). */
int h2e_compact_prepare(h2e_shape* s, int device);
/* bytes of the compact buffer for n_inst instances (0 before h2e_compact_prepare) */
size_t h2e_compact_bytes(const h2e_shape* s, uint64_t n_inst);
/* width class (1, 4 or 8) of every slot, n_slots bytes */
int h2e_compact_widths(const h2e_shape* s, uint8_t* out);
/* h2e_batch_run_host, delivering the compact form in h_compact (HOST buffer of h2e_compact_bytes) */
int h2e_batch_run_host_compact(h2e_shape* s, int device, uint64_t n_inst, const void* h_inputs, void* h_compact, uint32_t* h_status);
/* Host-side expansion compact -> vals[tile][slot][lane][32 bytes] with n_threads threads (the Rust shim would
 * expand while scattering cells into Records; this routine serves tests and plain-layout consumers). */
int h2e_expand_compact(const h2e_shape* s, uint64_t n_inst, const void* h_compact, void* h_vals, int n_threads);

/* Execution mode override (tuning / tests): mode 0 = automatic, 1 = one thread per instance,
 * 2 = team mode (`ctas_per_tile` CTAs per 32-instance tile execute the levelised program as a
 * dataflow of per-warp streams); ctas_per_tile 0 = automatic (SM count / tiles). Bits 8..15 of `mode`, if non-zero, set the number
 * of critical warps per CTA in team mode (default: by estimated work). Modes 3 and 4 are timing
 * experiments that skip macro-ops and do NOT produce records. */
int h2e_shape_set_mode(h2e_shape* s, int mode, int ctas_per_tile);

/* Measured peak rate of 32x32->64 multiply-adds (IMAD.WIDE.U32, 8 independent chains per thread, all
 * SMs) on `device`, in operations per second: the denominator of the integer-multiply roofline. */
int h2e_measure_imad_peak(int device, double* imad_per_sec);

/* Number of kernel launches issued by this library since load (for benchmarking evidence). */
uint64_t h2e_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
