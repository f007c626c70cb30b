// The CUDA witness-VM kernels for sm_100a. Compiled once per variant (-DH2E_TEAM_WARPS=8 / 16):
// the macro-op code is register hungry, and the two variants trade registers per thread (255 / 128)
// against resident warps per SM (8 / 16). Each variant exports plain host launchers (vm_kernel.h).
#include <cuda_runtime.h>

#include "schedule.h"
#include "vm_kernel.h"
#include "vm_ops.cuh"

using namespace h2e;

#ifndef H2E_TEAM_WARPS
#error "compile with -DH2E_TEAM_WARPS=8 or 16"
#endif
#define H2E_CAT2(a, b) a##b
#define H2E_CAT(a, b) H2E_CAT2(a, b)
#if defined(H2E_WIDTH_PROBE)
#define H2E_SFX probe  // width-probe build (compact export): stores write width classes, see vm_ops.cuh
#else
#define H2E_SFX H2E_TEAM_WARPS
#endif
#define h2e_vm_kernel H2E_CAT(h2e_vm_kernel_w, H2E_SFX)
#define h2e_montgomery_kernel H2E_CAT(h2e_montgomery_kernel_w, H2E_SFX)

namespace h2e {
__constant__ DeviceConsts g_consts;
}


// One thread = one circuit instance; a warp = 32 consecutive instances = one value tile, so every
// cell store of a warp is one contiguous 1 KiB run (one 256-bit store per lane). The program is
// uniform across the grid.
#ifndef H2E_BLOCK
#define H2E_BLOCK 128
#endif

// Thread mode (mode 0 of h2e_vm_kernel): many instances, short program (e.g. 2^20 int_mul blocks).
// One warp owns one tile and walks the whole program (P.crit, P.n_levels instructions) in order.
//
// Team mode (mode != 0): few instances, long program (a pairing check is ~175k macro-ops, an MSM
// millions, and only a few hundred instances fit in HBM). Dataflow execution, see schedule.h: `G` CTAs
// own one tile; each warp walks its own instruction stream and starts an instruction when the
// (warp, count) pairs of its dependency record are covered by the tile's progress counters.
//  * critical warps run the instructions other instructions depend on and publish their progress
//    (fence + store) after the instructions some other warp waits for;
//  * tail warps run the deferred instructions (the bulk of the record cells: int_mul / reduce TAILs,
//    asserts, ...), which nobody waits for, so record write-out streams while the critical path advances.
// All CTAs of the grid must be co-resident (the host sizes the grid to the SM count).

__device__ __forceinline__ u32 ld_progress(const u32* p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u32 ld_progress_acquire(const u32* p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fetch_dep(DepRec& d, const DepRec* p) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    d.n = v.x;
    d.d[0] = v.y;
    d.d[1] = v.z;
    d.d[2] = v.w;
}
// Wait until every dependency of `r` is covered by the tile's progress counters. Lane j polls
// dependency j; the loop leaves when all lanes are satisfied.
// acq_poll (tuning, H2E_ACQ_POLL=1): the polls themselves are acquire loads and no fence follows the loop.
__device__ __forceinline__ void wait_deps(const DepRec& r, const uint32_t* __restrict__ extra, const u32* progress, const volatile u32* s_progress,
                                          unsigned G, unsigned rank, unsigned lane, bool acq_poll) {
    const u32 n = r.n & 0xffffu;
    bool nonlocal = false;
    for (u32 base = 0; base < n; base += 32) {
        const u32 j = base + lane;
        u32 d = NONE;
        if (j < n) {
            if (n <= 3) d = j == 0 ? r.d[0] : (j == 1 ? r.d[1] : r.d[2]);
            else d = j < 2 ? (j == 0 ? r.d[0] : r.d[1]) : __ldg(extra + r.d[2] + j - 2);
        }
        const u32 dw = d == NONE ? 0 : (d >> DEP_SEQ_BITS);
        const bool local = dw % G == rank;  // producer stream runs in this CTA: its count is in shared memory
        nonlocal |= d != NONE && !local;
        const u32* addr = progress + dw;
        const volatile u32* saddr = s_progress + dw / G;
        const u32 want = d & ((1u << DEP_SEQ_BITS) - 1u);
        for (;;) {
            bool ok = d == NONE || (local ? *saddr : (acq_poll ? ld_progress_acquire(addr) : ld_progress(addr))) > want;
            if (__all_sync(0xffffffffu, ok)) break;
            __nanosleep(20);
        }
    }
    // Acquire side of the hand-over. The polls are relaxed (cheap to spin on) and only lane j saw counter j, so one
    // fence after the loop orders every lane's later operand loads behind the observed counts: CTA scope when all
    // producers publish through shared memory, GPU scope (pairs with the producer's st.release.gpu) otherwise.
    if (n) {
        if (!acq_poll && __any_sync(0xffffffffu, nonlocal)) __threadfence();
        else __threadfence_block();
    }
}

// (one kernel for both modes: two kernels calling the macro-op dispatcher crash cicc 12.9)
__global__ void __launch_bounds__(H2E_TEAM_WARPS * 32, 1)
    h2e_vm_kernel(TeamProg P, u32* __restrict__ vals, const u32* __restrict__ inputs, const u32* __restrict__ cpool,
                  const u32* __restrict__ tables, u32* __restrict__ status, u32* __restrict__ progress, u32* __restrict__ scratch, uint32_t n_scratch,
                  uint64_t tile_words, uint32_t n_in_cells, uint64_t n_inst, uint64_t n_tiles, int mode) {
    const unsigned lane = threadIdx.x % TILE, warp = threadIdx.x / TILE;
    // The macro-ops are out-of-line functions taking (LaneCtx&, const Instr&): both objects must live in memory. On the stack
    // (local memory) every access to them is an L1 transaction queued behind the warp's own record stores -- ncu's source view
    // showed the stall samples of the thread-mode kernel spread over the consumers of such LDL loads (75 per tile-program, each
    // waiting out the store queue). In shared memory they are 30-cycle accesses on a path the store stream does not block.
    __shared__ Instr s_instr[H2E_TEAM_WARPS];
    __shared__ LaneCtx s_lane[H2E_TEAM_WARPS * TILE];
    LaneCtx& ln = s_lane[threadIdx.x];
    Instr& in = s_instr[warp];
    if (mode == 0) {
        uint64_t tile = (uint64_t)blockIdx.x * (blockDim.x / TILE) + warp;
        if (tile >= n_tiles) return;
        uint64_t inst = tile * TILE + lane;
        uint64_t in_inst = inst < n_inst ? inst : (n_inst - 1);  // padding lanes recompute the last instance
#if defined(H2E_WIDTH_PROBE)
        ln.vals = vals;  // one cell per slot, shared by the 32 lanes (they all store the same width class)
        ln.lane = 0;
#else
        ln.vals = vals + tile * tile_words;  // this tile's block of the COMPACT records (layout.h)
        ln.lane = lane;
#endif
        ln.inputs = inputs + in_inst * (uint64_t)n_in_cells * 8;
        ln.cpool = cpool;
        ln.tables = tables;
        ln.scratch = nullptr;
        ln.status = 0;
        // the instance's inputs are the only DRAM reads of a short program: pull them towards the SM while the first
        // instructions are fetched
        for (uint32_t c = 0; c < n_in_cells && c < 64; c += 4)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ln.inputs + (size_t)c * 8));
        const uint32_t n_instr = P.n_levels;
        for (uint32_t pc = 0; pc < n_instr; pc++) {
            // the warp's instruction: lanes 0..3 copy 16 bytes each into the warp's shared-memory slot
            __syncwarp();
            if (lane < 4) reinterpret_cast<uint4*>(&in)[lane] = __ldg(reinterpret_cast<const uint4*>(P.crit + pc) + lane);
            __syncwarp();
            exec_instr(ln, in);
        }
        status[inst] = ln.status;
        return;
    }
    const bool acq_poll = (mode & 0x100) != 0;
    mode &= 0xff;
    const int dry_run = mode == 3, dry_tail = mode == 3 || mode == 4;  // modes 3, 4: timing experiments only
    __shared__ volatile u32 s_progress[H2E_TEAM_WARPS];  // counts of this CTA's critical streams
    if (threadIdx.x < H2E_TEAM_WARPS) s_progress[threadIdx.x] = 0;
    __syncthreads();
    const unsigned G = P.G, rank = blockIdx.x % G;
    const uint64_t tile = blockIdx.x / G;
    const uint64_t inst = tile * TILE + lane;
    const uint64_t in_inst = inst < n_inst ? inst : (n_inst - 1);
    ln.vals = vals + tile * tile_words;
    ln.lane = lane;
    ln.inputs = inputs + in_inst * (uint64_t)n_in_cells * 8;
    ln.cpool = cpool;
    ln.tables = tables;
    ln.scratch = scratch + (tile * n_scratch * TILE + lane) * 16;
    ln.status = 0;
    u32* prog_tile = progress + tile * P.twc;
    // role and stream of this warp (TeamLayout, schedule.h); neighbouring streams sit on different SMs
    const bool split = P.g_crit != 0;
    const bool critical = split ? rank < P.g_crit : warp < P.n_crit;
    const unsigned cstride = split ? P.g_crit : G;  // critical stream w runs in CTA w % cstride
    const unsigned tw = split ? (critical ? warp * P.g_crit + rank : warp * (G - P.g_crit) + (rank - P.g_crit))
                              : (critical ? warp : warp - P.n_crit) * G + rank;
    const Instr* code = critical ? P.crit : P.tail;
    const DepRec* deps = critical ? P.crit_dep : P.tail_dep;
    const uint32_t* off = critical ? P.crit_off : P.tail_off;
    const uint32_t b = __ldg(off + tw), e = __ldg(off + tw + 1);
    // next instruction, prefetched: lanes 0..3 hold 16 bytes of it each
    uint4 nxt = make_uint4(0, 0, 0, 0);
    DepRec nxt_dep;
    if (b < e) {
        if (lane < 4) nxt = __ldg(reinterpret_cast<const uint4*>(code + b) + lane);
        fetch_dep(nxt_dep, deps + b);
    }
#ifdef H2E_PROFILE
    // development build: cycle counters per warp, and per opcode for the CTAs of tile 0 (exp/profile_team.py)
    long long t_wait = 0, t_exec = 0, t_pub = 0, t0;
    __shared__ unsigned long long s_op_exec[48], s_op_wait[48], s_op_cnt[48];
    for (unsigned i = threadIdx.x; i < 48; i += blockDim.x) s_op_exec[i] = s_op_wait[i] = s_op_cnt[i] = 0;
    __syncthreads();
    const long long t_begin = clock64();
#define PROF_T0() t0 = clock64()
#define PROF_ADD(x) x += clock64() - t0
#define PROF_OP(arr)                                                                   \
    if (lane == 0 && tile == 0) atomicAdd(&arr[in.op < 48 ? in.op : 47], (unsigned long long)(clock64() - t0))
#else
#define PROF_OP(arr)
#define PROF_T0()
#define PROF_ADD(x)
#endif
    for (uint32_t k = b; k < e; k++) {
        __syncwarp();
        if (lane < 4) reinterpret_cast<uint4*>(&in)[lane] = nxt;
        __syncwarp();
        const DepRec dep = nxt_dep;
        if (k + 1 < e) {
            if (lane < 4) nxt = __ldg(reinterpret_cast<const uint4*>(code + k + 1) + lane);
            fetch_dep(nxt_dep, deps + k + 1);
        }
        PROF_T0();
        wait_deps(dep, P.extra, prog_tile, s_progress, cstride, rank, lane, acq_poll);
        PROF_OP(s_op_wait);
        PROF_ADD(t_wait);
        PROF_T0();
        if (!(critical ? dry_run : dry_tail)) exec_instr(ln, in);
        else ln.status |= (in.op == 0xffff);
        PROF_OP(s_op_exec);
        PROF_ADD(t_exec);
#ifdef H2E_PROFILE
        if (lane == 0 && tile == 0) atomicAdd(&s_op_cnt[in.op < 48 ? in.op : 47], 1ull);
#endif
        PROF_T0();
        if (dep.n & (3u << 16)) {
            // release: this warp's cells, then the count (other warps read the cells after seeing the count).
            // Warps of this CTA watch the shared-memory count (CTA-scope fence), other CTAs the global one.
            __syncwarp();
            if (lane == 0) {
                if (dep.n & (2u << 16)) {
                    __threadfence_block();
                    s_progress[warp] = k - b + 1;
                }
                if (dep.n & (1u << 16)) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(prog_tile + tw), "r"(k - b + 1) : "memory");
            }
        }
        PROF_ADD(t_pub);
    }
#ifdef H2E_PROFILE
    if (lane == 0 && tile == 0)
        printf("W %s cta %u warp %u n %u wait %lld exec %lld publish %lld total %lld\n", critical ? "crit" : "tail", rank, warp, e - b, t_wait, t_exec,
               t_pub, (long long)(clock64() - t_begin));
    __syncthreads();
    if (tile == 0 && threadIdx.x < 48 && s_op_cnt[threadIdx.x])
        printf("O %s cta %u op %u cnt %llu exec %llu wait %llu\n", critical ? "crit" : "tail", rank, threadIdx.x, s_op_cnt[threadIdx.x],
               s_op_exec[threadIdx.x], s_op_wait[threadIdx.x]);
#endif
    if (ln.status) atomicOr(&status[inst], ln.status);
}

// Export pass: canonical little-endian cells -> halo2's in-memory Fr (Montgomery form x * 2^256 mod r,
// four little-endian u64 limbs), in place. One thread per cell, 256-bit load and store; HBM-bound.
__global__ void __launch_bounds__(256) h2e_montgomery_kernel(u32* __restrict__ cells, uint64_t n_cells) {
    const FrConst& F = g_consts.fr;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += (uint64_t)gridDim.x * blockDim.x) {
        u32 x[8], y[8];
        ld8(x, cells + i * 8);
        mont_mul<8>(y, x, F.r2, F.r, F.minv);
        st_raw8(cells + i * 8, y);
    }
}


// Integer-multiply roofline probe: every thread runs 8 independent IMAD.WIDE.U32 chains. The measured
// rate is the denominator of the "fraction of the integer-multiply roofline" bench.py reports.
__global__ void __launch_bounds__(256) H2E_CAT(h2e_imad_probe_w, H2E_SFX)(u64* out, uint32_t iters, u32 seed) {
    u64 acc[8];
    u32 a = seed + threadIdx.x, b = seed * 2654435761u + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = (u64)k * 0x9e3779b97f4a7c15ull + a;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = (u64)(a + k) * (u32)(b ^ (u32)acc[k]) + acc[k];  // IMAD.WIDE.U32 with 64-bit accumulate
    }
    u64 r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= acc[k];
    if (r == 0x1234567u) out[0] = r;  // keep the chains alive
}

#if H2E_TEAM_WARPS == 8 && !defined(H2E_WIDTH_PROBE)
// 256-bit evict-first store (STG.E.EF.ENL2.256) for the export kernels: their output is read next by the copy engine or the host
__device__ __forceinline__ void st256_cs(u32* p, const u32* c) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]),
                 "r"(c[6]), "r"(c[7])
                 : "memory");
}
// Record export kernels (layout.h). The VM's own records are COMPACT: for every slot its static width class
// w(s) in {1, 4, 8} words per lane, rec[tile][32 * off(s) + lane * w + k]. One warp per (tile, slot) everywhere:
// a cell's 32 lanes are 128 / 512 / 1024 contiguous bytes.
__device__ __forceinline__ void ld_cell(u32* c, const u32* src, u32 w, unsigned lane) {
    if (w == 8) {
        ld8(c, src + lane * 8);
    } else {
#pragma unroll
        for (int k = 1; k < 8; k++) c[k] = 0;
        if (w == 4) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src + lane * 4));
            c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
        } else {
            c[0] = __ldcs(src + lane);
        }
    }
}
// UNIQUE form: the selected slots `sel[i0 .. i0 + n_i)` (the roots of the copy classes) back to back, uoff[i] =
// words per lane before selected slot i. Packs n_tiles tiles into `out`, tile stride out_tile_words, the piece
// starting at word 0 of every tile's block.
__global__ void __launch_bounds__(256) h2e_pack_kernel(const u32* __restrict__ rec, u32* __restrict__ out, const u32* __restrict__ sel,
                                                       const u32* __restrict__ uoff, const u32* __restrict__ coff, uint64_t tile_words, uint32_t i0,
                                                       uint32_t n_i, uint64_t n_tiles, uint64_t out_tile_words) {
    const unsigned lane = threadIdx.x % TILE;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x / TILE), total = (uint64_t)n_i * n_tiles;
    const u32 o0 = __ldg(uoff + i0);
    for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x / TILE) + threadIdx.x / TILE; i < total; i += warps) {
        const uint64_t tile = i / n_i;
        const u32 k = i0 + (u32)(i % n_i);
        const u32 s = __ldg(sel + k);
        const u32 o = __ldg(uoff + k), w = __ldg(uoff + k + 1) - o;
        u32 c[8];
        ld_cell(c, rec + tile * tile_words + (uint64_t)__ldg(coff + s) * TILE, w, lane);
        u32* dst = out + tile * out_tile_words + ((uint64_t)(o - o0) * TILE + (uint64_t)lane * w);
        if (w == 8) st256_cs(dst, c);
        else if (w == 4) __stcs(reinterpret_cast<uint4*>(dst), make_uint4(c[0], c[1], c[2], c[3]));
        else __stcs(dst, c[0]);
    }
}
// WIDE form: vals[tile][slot][lane][8 words], slots [s0, s0 + n_s) of n_tiles tiles, written at out + tile * out_tile_words + (s - s0) * 256.
__global__ void __launch_bounds__(256) h2e_expand_kernel(const u32* __restrict__ rec, u32* __restrict__ out, const u32* __restrict__ coff,
                                                         uint64_t tile_words, uint64_t s0, uint64_t n_s, uint64_t n_tiles, uint64_t out_tile_words) {
    const unsigned lane = threadIdx.x % TILE;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x / TILE), total = n_s * n_tiles;
    for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x / TILE) + threadIdx.x / TILE; i < total; i += warps) {
        const uint64_t tile = i / n_s, s = s0 + i % n_s;
        const u32 o = __ldg(coff + s), w = __ldg(coff + s + 1) - o;
        u32 c[8];
        ld_cell(c, rec + tile * tile_words + (uint64_t)o * TILE, w, lane);
        st256_cs(out + tile * out_tile_words + ((s - s0) * TILE + lane) * 8, c);
    }
}

// Prover hand-off (the step after the hot path: Records::assign_all / _assign_to_*, context.rs:303-588, lays the
// records out as advice columns): scatter the records into one dense cell array per instance,
//   out[instance][dst[slot]][8 words],
// dst[slot] = index of the slot's advice cell in the caller's order (column-major = the prover's advice columns,
// or row-major = RecordsInner's [row][col]); cells no slot maps to keep the zeros the caller put there. With
// `mont` the cells are written as x * 2^256 mod r (halo2's in-memory Fr).
// The records are instance-minor (a cell's 32 instances are contiguous), the output instance-major: a CTA
// transposes one (tile, group of 32 slots) block through shared memory. Slots are taken in the order of their
// destination (`ord` = slots sorted by dst), so the 32 cells a warp then writes for one instance are one
// contiguous 1 KiB run wherever the destination cells are consecutive (a column's rows), instead of 32 single
// sectors 32 x cells_per_inst bytes apart.
static const int SC_GROUP = 32;          // slots per block
static const int SC_STRIDE = 32 * 8 + 4;  // words per slot row in shared memory (+4: the transposed 128-bit reads hit 8 different bank quads)
__global__ void __launch_bounds__(256) h2e_scatter_kernel(const u32* __restrict__ rec, u32* __restrict__ out, const u32* __restrict__ dst,
                                                          const u32* __restrict__ ord, const u32* __restrict__ coff, uint64_t tile_words,
                                                          uint64_t n_slots, uint64_t inst0, uint64_t n_inst, uint64_t cells_per_inst, int mont) {
    const FrConst& F = g_consts.fr;
    __shared__ __align__(16) u32 sm[SC_GROUP * SC_STRIDE];
    const unsigned lane = threadIdx.x % TILE, warp = threadIdx.x / TILE;
    const uint64_t tiles = (n_inst + TILE - 1) / TILE, groups = (n_slots + SC_GROUP - 1) / SC_GROUP;
    for (uint64_t item = blockIdx.x; item < groups * tiles; item += gridDim.x) {
        const uint64_t tile = item / groups, k0 = (item % groups) * SC_GROUP;
        // phase 1: 4 slots per warp, lane = instance (one contiguous run of the tile's block per slot)
#pragma unroll
        for (int i = 0; i < SC_GROUP / 8; i++) {
            const unsigned j = warp * (SC_GROUP / 8) + i;
            if (k0 + j >= n_slots) break;
            const u32 s = __ldg(ord + k0 + j);
            const u32 o = __ldg(coff + s), w = __ldg(coff + s + 1) - o;
            u32 c[8];
            ld_cell(c, rec + tile * tile_words + (uint64_t)o * TILE, w, lane);
            if (mont) {
                // x * 2^256 mod r; the slot's width class is uniform over the warp: a 1-word cell costs one CIOS round
                // (16 multiplications), a limb four, a full field element eight
                u32 y[8];
                if (w == 8) mont_mul<8>(y, c, F.r2, F.r, F.minv);
                else if (w == 4) mont_mul_short<8, 4>(y, c, F.r2w4, F.r, F.minv);
                else mont_mul_short<8, 1>(y, c, F.r2w1, F.r, F.minv);
#pragma unroll
                for (int k = 0; k < 8; k++) c[k] = y[k];
            }
            uint4* q = reinterpret_cast<uint4*>(sm + j * SC_STRIDE + lane * 8);
            q[0] = make_uint4(c[0], c[1], c[2], c[3]);
            q[1] = make_uint4(c[4], c[5], c[6], c[7]);
        }
        __syncthreads();
        // phase 2: 4 instances per warp, lane = slot of the group (in destination order)
        const bool live = k0 + lane < n_slots;
        const u32 d = live ? __ldg(dst + __ldg(ord + k0 + lane)) : 0;
#pragma unroll
        for (int i = 0; i < TILE / 8; i++) {
            const unsigned il = warp * (TILE / 8) + i;
            const uint64_t inst = tile * TILE + il;
            if (!live || inst >= n_inst) continue;
            const uint4* q = reinterpret_cast<const uint4*>(sm + lane * SC_STRIDE + il * 8);
            const uint4 a = q[0], b = q[1];
            const u32 c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            st256_cs(out + ((inst0 + inst) * cells_per_inst + d) * 8, c);
        }
        __syncthreads();
    }
}
#endif

// ---- host launchers of this variant ----
namespace h2e {

cudaError_t H2E_CAT(vm_upload_consts_w, H2E_SFX)(const DeviceConsts* c) { return cudaMemcpyToSymbol(g_consts, c, sizeof(DeviceConsts)); }

cudaError_t H2E_CAT(vm_launch_w, H2E_SFX)(const VmLaunch& L) {
    if (L.mode != 0) {
        // team mode: warps spin on each other's progress, so every CTA of the grid must be resident at the same
        // time. A cooperative launch makes the driver guarantee that (or fail), also against concurrent kernels.
        VmLaunch a = L;
        void* args[] = {&a.prog, &a.vals, &a.inputs, &a.cpool, &a.tables, &a.status, &a.progress, &a.scratch, &a.n_scratch,
                        &a.tile_words, &a.n_in_cells, &a.n_inst, &a.n_tiles, &a.mode};
        return cudaLaunchCooperativeKernel((const void*)h2e_vm_kernel, dim3(L.grid), dim3(L.block), args, 0, L.stream);
    }
    h2e_vm_kernel<<<L.grid, L.block, 0, L.stream>>>(L.prog, L.vals, L.inputs, L.cpool, L.tables, L.status, L.progress, L.scratch, L.n_scratch, L.tile_words, L.n_in_cells,
                                                    L.n_inst, L.n_tiles, L.mode);
    return cudaGetLastError();
}

#if H2E_TEAM_WARPS == 8 && !defined(H2E_WIDTH_PROBE)
cudaError_t vm_pack(cudaStream_t stream, unsigned blocks, const u32* rec, u32* out, const u32* sel, const u32* uoff, const u32* coff,
                    uint64_t tile_words, uint32_t i0, uint32_t n_i, uint64_t n_tiles, uint64_t out_tile_words) {
    h2e_pack_kernel<<<blocks, 256, 0, stream>>>(rec, out, sel, uoff, coff, tile_words, i0, n_i, n_tiles, out_tile_words);
    return cudaGetLastError();
}
cudaError_t vm_expand(cudaStream_t stream, unsigned blocks, const u32* rec, u32* out, const u32* coff, uint64_t tile_words, uint64_t s0, uint64_t n_s,
                      uint64_t n_tiles, uint64_t out_tile_words) {
    h2e_expand_kernel<<<blocks, 256, 0, stream>>>(rec, out, coff, tile_words, s0, n_s, n_tiles, out_tile_words);
    return cudaGetLastError();
}
cudaError_t vm_scatter(cudaStream_t stream, unsigned blocks, const u32* rec, u32* out, const u32* dst, const u32* ord, const u32* coff,
                       uint64_t tile_words, uint64_t n_slots, uint64_t inst0, uint64_t n_inst, uint64_t cells_per_inst, int mont) {
    // 6 CTAs x 33.5 KB of shared memory per SM (a per-device attribute: set on every launch, it is cheap)
    cudaFuncSetAttribute(h2e_scatter_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    h2e_scatter_kernel<<<blocks, 256, 0, stream>>>(rec, out, dst, ord, coff, tile_words, n_slots, inst0, n_inst, cells_per_inst, mont);
    return cudaGetLastError();
}
cudaError_t vm_imad_probe(cudaStream_t stream, unsigned blocks, u64* out, uint32_t iters) {
    H2E_CAT(h2e_imad_probe_w, H2E_SFX)<<<blocks, 256, 0, stream>>>(out, iters, 12345u);
    return cudaGetLastError();
}
#endif

cudaError_t H2E_CAT(vm_montgomery_w, H2E_SFX)(cudaStream_t stream, unsigned blocks, u32* cells, uint64_t n_cells) {
    h2e_montgomery_kernel<<<blocks, 256, 0, stream>>>(cells, n_cells);
    return cudaGetLastError();
}

}  // namespace h2e
