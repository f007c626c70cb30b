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
 * points as per-instance constants (src/circuit/pairing_chip.rs:157-176: pairing, check_pairing, multi_miller_loop,
 * final_exponentiation), Fq2 / Fq6 / Fq12ChipOps (src/circuit/fq12.rs:10-459), ecc_mul and the general-scalar MSM
 * (src/circuit/general_scalar_ecc_chip.rs:96-147), and KeccakChipOps (src/circuit/keccak_chip.rs:53-307: hash, init,
 * absorb, permute, theta / rho_and_pi / xi / iota, decompose_scalar_as_u256_be, compose_to_scalar_be).
 * `statics64` are shape-level constants, 64 bytes each. Replaces: constructing a Context and calling the trait methods. */
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
 * calls once per instance on the CPU (e.g. src/circuit/integer_chip.rs:466-483 for int_mul).
 * d_vals holds h2e_vals_bytes(s, n_inst) bytes and d_status ceil(n_inst / 32) * 32 words (whole tiles: the
 * padding lanes of the last tile are written too). Shapes with long programs (pairing, MSM) run as cooperative
 * launches of at most SMs / 4 tiles each (4 CTAs per tile), back to back on `stream`. */
int h2e_batch_run(h2e_shape* s, int device, void* stream, uint64_t n_inst, const void* d_inputs, void* d_vals, uint32_t* d_status);
/* The same, delivering the records in `format` (H2E_REC_*, below) in d_records (h2e_records_bytes). The VM itself
 * writes H2E_REC_COMPACT -- that call is the fast path and needs no other device memory; WIDE (== h2e_batch_run) and
 * UNIQUE are derived from a stream-ordered temporary holding the COMPACT records. */
int h2e_batch_run_records(h2e_shape* s, int device, void* stream, int format, uint64_t n_inst, const void* d_inputs, void* d_records, uint32_t* d_status);

/* Same with HOST buffers: copies inputs to the device, runs, copies values (WIDE layout) and status (n_inst
 * words) back. A convenience wrapper over a stream (below) kept in the shape handle. */
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

/* ---- record formats -------------------------------------------------------------------------
 * The value half of the records exists in four layouts, all tile-interleaved over 32 instances
 * (tile t = instances 32t .. 32t+31, lane = instance % 32):
 *   H2E_REC_WIDE     vals[tile][slot][lane][32 bytes]: every advice cell a canonical 32-byte Fr (what
 *                    h2e_batch_run delivers).
 *   H2E_REC_COMPACT  rec[tile][32 * off(slot) + lane * w(slot) + k] (32-bit words): every cell at its static width
 *                    class w in {1, 4, 8} words. Of the 125 cells of an int_mul block 60 are 18-bit range chunks
 *                    and 40 are 108-bit limbs; the dropped words are zero, so the form is lossless, ~2.4x smaller.
 *                    This is the layout the VM's macro-ops write and read (a warp stores whole 128-byte lines).
 *   H2E_REC_UNIQUE   COMPACT without the cells that are copies. Every permutation pair of the records
 *                    (Records::permutations, src/context.rs:648-658; h2e_shape_perms) ties a new cell to an
 *                    older cell that holds the same value; only the oldest cell of every such class (its
 *                    "root") is stored. ~5x smaller than WIDE: this is what the PCIe / host-memory bound host
 *                    path moved by default in round 2a. The consumer fills cell s from root[s] (h2e_records_expand does).
 *   H2E_REC_PRIMARY  UNIQUE without the range chip's chunk cells. The 18-bit chunks of a limb's range rows
 *                    (assign_nonleading_limb / assign_w_ceil_leading_limb / assign_d_leading_limb / assign_common,
 *                    src/context.rs:835-997) are bit fields of the row's accumulator cell, which is stored:
 *                    chunk = (cell[src] >> shift) & (2^18 - 1) (h2e_shape_layout_derived). ~7x smaller than WIDE;
 *                    the default of the host path. h2e_records_expand rebuilds every cell (shifts and copies only,
 *                    no field arithmetic).
 * Widths, offsets, roots and derivations are static per shape (host side, no device needed). */
enum { H2E_REC_WIDE = 0, H2E_REC_COMPACT = 1, H2E_REC_UNIQUE = 2, H2E_REC_PRIMARY = 3 };
/* Any output pointer may be NULL. off_out[n_slots + 1]: words per lane before slot s in `format` (in UNIQUE a
 * copy takes no room: off[s + 1] == off[s], read it at off[root[s]]); width_out[n_slots]: 1, 4 or 8;
 * root_out[n_slots]: the slot whose value slot s repeats (itself if it is no copy). */
int h2e_shape_layout(h2e_shape* s, int format, uint32_t* off_out, uint8_t* width_out, uint32_t* root_out);
/* Derivations of H2E_REC_PRIMARY: src_out[n_slots] = the (stored) slot that slot s is an 18-bit field of, or
 * 0xffffffff; shift_out[n_slots] = the field's shift in bits, 255 = the cell is the constant 0. Only roots carry a
 * derivation: resolve a copy through root[s] first. Either pointer may be NULL. */
int h2e_shape_layout_derived(h2e_shape* s, uint32_t* src_out, uint8_t* shift_out);
/* bytes of the records of n_inst instances (whole tiles) in `format` */
size_t h2e_records_bytes(h2e_shape* s, int format, uint64_t n_inst);
/* cells of one instance's dense advice array: sum over regions of columns x height (base 5, range 3, select 2) */
uint64_t h2e_shape_dense_cells(const h2e_shape* s);

/* h2e_batch_run_host delivering `format` in h_records (HOST buffer of h2e_records_bytes). */
int h2e_batch_run_host_records(h2e_shape* s, int device, int format, uint64_t n_inst, const void* h_inputs, void* h_records, uint32_t* h_status);

/* Consumer side, on the host (what a Rust shim does while it fills RecordsInner): records in `format` -> plain
 * 32-byte cells with n_threads threads; copies are filled from their roots.
 *   mode 0: vals[tile][slot][lane][32 bytes] (the WIDE layout);
 *   mode 1: out[instance][cell][32 bytes], cells in column-major order (region, column, row): the advice columns
 *           Records::assign_all produces (src/context.rs:303-588);
 *   mode 2: same in row-major order (region, row, column): RecordsInner's own indexing (src/context.rs:241-252).
 * Modes 1 / 2 need n_inst * h2e_shape_dense_cells * 32 bytes; cells no slot maps to are zero. */
int h2e_records_expand(h2e_shape* s, int format, int mode, uint64_t n_inst, const void* h_records, void* h_out, int n_threads);

/* The same hand-off on the DEVICE (records never leave HBM: the GPU prover's advice columns): COMPACT records d_records
 * (as filled by h2e_batch_run_records for n_inst instances) -> d_out[inst0 + instance][cell][32 bytes], order 1 =
 * column-major, 2 = row-major (as modes 1 / 2 above), encoding H2E_EXPORT_CANONICAL or H2E_EXPORT_MONTGOMERY.
 * The caller zeroes d_out beforehand if unassigned cells matter. Asynchronous on `stream`. */
int h2e_records_scatter(h2e_shape* s, int device, void* stream, uint64_t n_inst, const void* d_records, void* d_out, uint64_t inst0, int order, int encoding);

/* ---- streaming (chunked) host path ------------------------------------------------------------
 * The batch sizes of the real workloads do not fit one buffer (1024 pairing checks = 202 GB of cells, 4096
 * MSMs of 4096 points = 19.8 TB), and the reference bounds memory per instance (src/context.rs:254-292). A stream
 * processes the batch chunk by chunk: inputs up, VM, export kernel, records down into the CALLER's pinned host
 * buffer, on two CUDA streams so that chunk k+1 computes while chunk k crosses the bus. The caller keeps a ring of
 * host buffers, submits a chunk per buffer, and reuses a buffer once its ticket has completed. The library
 * retains no caller pointer past the completion of the ticket. One thread at a time per stream handle; different
 * handles (devices) are independent. */
typedef struct h2e_stream h2e_stream;
/* chunk_bytes_hint: target size of one chunk's WIDE cells on the device (0 = default: 128 MiB for short programs,
 * as many tiles as one team-mode launch takes for long ones); the actual geometry is read with h2e_stream_query. */
h2e_stream* h2e_stream_open(h2e_shape* s, int device, int format, size_t chunk_bytes_hint);
/* out[0] instances per chunk (max; a multiple of 32), out[1] bytes of a full chunk's records, out[2] bytes per
 * tile, out[3] chunks in flight on the device, out[4] 1 if a tile is exported in slot-range pieces (it does not
 * fit a staging buffer), out[5] tickets that may be outstanding, out[6] chunks submitted so far. */
int h2e_stream_query(const h2e_stream* st, uint64_t out[8]);
/* Queue one chunk: n_inst <= out[0] instances, inputs[n_inst][input cell][32 bytes], records for
 * ceil(n_inst / 32) tiles, status[n_inst]. Host pointers; h_inputs and h_records should be pinned (a copy from or to
 * pageable memory blocks the call until the chunk has drained, which serialises the pipeline); h_status may be
 * ordinary memory: it is filled when h2e_stream_poll / h2e_stream_wait reports the ticket complete. Returns at once. */
int h2e_stream_submit(h2e_stream* st, uint64_t n_inst, const void* h_inputs, void* h_records, uint32_t* h_status, uint64_t* ticket);
/* 0 = the chunk's records and status are in the host buffers, 1 = not yet, < 0 = error */
int h2e_stream_poll(h2e_stream* st, uint64_t ticket);
int h2e_stream_wait(h2e_stream* st, uint64_t ticket);
int h2e_stream_close(h2e_stream* st);

/* ---- compact export, round-1 names (== H2E_REC_COMPACT) ---------------------------------------
 * h2e_compact_prepare additionally cross-checks the static width table against the device code: a build of
 * the VM whose stores record their width class runs the program once. */
int h2e_compact_prepare(h2e_shape* s, int device);
size_t h2e_compact_bytes(const h2e_shape* s, uint64_t n_inst);
/* width class (1, 4 or 8) of every slot, n_slots bytes */
int h2e_compact_widths(const h2e_shape* s, uint8_t* out);
int h2e_batch_run_host_compact(h2e_shape* s, int device, uint64_t n_inst, const void* h_inputs, void* h_compact, uint32_t* h_status);
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
