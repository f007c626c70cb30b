// Witness-VM macro-ops: one function per chip call of the reference, each computing every advice
// cell that call assigns, in the reference's assignment order, for ONE instance (one GPU thread).
// Cell layouts follow SURVEY Appendix A; per-function comments cite the reference lines.
//
// Portable: compiled by nvcc for sm_100a (product path) and, for the CPU test-suite only, as plain
// C++ by the host emulator (tests/emu). The product never runs the host build.
#pragma once
#include "bigint.cuh"
#include "h2e_program.h"
#include "modinv30.cuh"

namespace h2e {

// ------------------------------- compile-time field traits ----------------------------------
template <int FID>
struct FT;
template <>
struct FT<F_BN256_FQ> {
    static constexpr int L = 3, M = 3, R = 1, P = 1, NBITS = 254, NW = 8, NXA = 9, KBITS = 520, NX = 17, ND = 9;
    static constexpr int WLEAD = 38, DLEAD = 51, WDEC = 3, DDEC = 3;
};
template <>
struct FT<F_BLS12_381_FQ> {
    static constexpr int L = 4, M = 5, R = 2, P = 2, NBITS = 381, NW = 12, NXA = 13, KBITS = 774, NX = 25, ND = 13;
    static constexpr int WLEAD = 57, DLEAD = 70, WDEC = 4, DDEC = 4;
};
template <>
struct FT<F_BLS12_381_FR> {
    static constexpr int L = 3, M = 3, R = 1, P = 1, NBITS = 255, NW = 8, NXA = 9, KBITS = 522, NX = 17, ND = 9;
    static constexpr int WLEAD = 39, DLEAD = 52, WDEC = 3, DDEC = 3;
};

// Field constants live in the constant bank on the device; referring to the symbol directly (not
// through a pointer carried in LaneCtx) lets ptxas fold them into c[bank][offset] operands instead
// of issuing ~190 generic loads per int_mul.
#if defined(__CUDA_ARCH__)
extern __constant__ DeviceConsts g_consts;
#define H2E_CONSTS g_consts
#else
extern const DeviceConsts* g_host_consts;
#define H2E_CONSTS (*g_host_consts)
#endif

// ------------------------------- per-lane memory view ---------------------------------------
// Record layout of the VM (layout.h: COMPACT). A tile holds, for every slot, its static width class w in {1, 4, 8}
// words per lane: words [32 * off(slot) + lane * w + k] of the tile's block. A warp (lane = instance) therefore
// writes one whole 128-byte line per 1-word cell, 512 contiguous bytes per limb cell and 1 KiB per field element,
// and the VM moves only the significant words (2.4x fewer bytes than 32 per cell). Slot references in the
// DEVICE copy of a program are pre-translated by the host: ref = off(slot) << 2 | class (0: w = 1, 1: w = 4,
// 2: w = 8), so an operand load knows where its cell starts and how wide it is.
// H2E_WIDTH_PROBE build (host emulator and one device variant; checks the static width table, layout.h): plain
// layout, one 8-word cell per slot, references are slot numbers, and every store writes the width class of its
// call site instead of the value.
struct LaneCtx {
    u32* vals;           // compact: base of this tile's block (uniform over the warp); probe: cell 0
    u32 lane;            // instance within the tile
    const u32* inputs;   // this instance's inputs, instance-major: input cell i at inputs[i*8]
    const u32* cpool;    // constant pool, 8 words per entry (shared by all instances)
    const u32* tables;   // slot tables (OP_SELECT_INT), entries are references
    u32* scratch;        // team mode: this lane's view of the tile's scratch entries (16 words each, entry stride 32*16)
    u32 status;
};

#if defined(H2E_WIDTH_PROBE)
#define H2E_ADV(w) 8   // words between consecutive cells
#else
#define H2E_ADV(w) ((w) * TILE)
#endif

// One advice cell = 32 bytes per lane. sm_100a has 256-bit global stores/loads (STG.E.ENL2.256):
// one instruction per cell writes whole 32-byte sectors, 1 KiB contiguous per warp. Measured on
// B200 for this exact pattern: 6.7-7.4 TB/s with 256-bit stores vs 2.6-3.2 TB/s when the cell is
// split into two 128-bit stores (each then covers only half of every sector).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void st256(u32* p, u32 a, u32 b, u32 c, u32 d, u32 e, u32 f, u32 g, u32 h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f),
                 "r"(g), "r"(h)
                 : "memory");
}
#endif
#if defined(__CUDA_ARCH__)
// evict-first variant (STG.E.EF.ENL2.256) for cells nobody reads back: the record stream of the tail
// warps must not push the operand cells of the critical path out of L2
__device__ __forceinline__ void st256_stream(u32* p, u32 a, u32 b, u32 c, u32 d, u32 e, u32 f, u32 g, u32 h) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f),
                 "r"(g), "r"(h)
                 : "memory");
}
#endif
// H2E_WIDTH_PROBE build (compact export, h2e_compact_*): every store writes the WIDTH CLASS of the cell
// (1, 4 or 8 significant words, fixed by the call site) instead of its value. One pass of a shape's
// program over a dummy tile then yields the static width of every slot.
H2E_HD void st_probe(u32* p, u32 width_class) {
#if defined(__CUDA_ARCH__)
    st256(p, width_class, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
#else
    p[0] = width_class;
    for (int k = 1; k < 8; k++) p[k] = 0;
#endif
}
// raw 8-word store (scratch entries)
H2E_HD void st_raw8(u32* p, const u32* w) {
#if defined(__CUDA_ARCH__)
    st256(p, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
#else
    for (int k = 0; k < 8; k++) p[k] = w[k];
#endif
}
H2E_HD void ld8(u32* w, const u32* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p)
                 : "memory");
#else
    for (int k = 0; k < 8; k++) w[k] = p[k];
#endif
}
H2E_HD void ld4(u32* w, const u32* p) {
#if defined(__CUDA_ARCH__)
    uint4 a = reinterpret_cast<const uint4*>(p)[0];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
#else
    for (int k = 0; k < 4; k++) w[k] = p[k];
#endif
}

// plain stores of 1 / 4 words (compact cells narrower than 32 bytes); STREAM = evict-first
template <int STREAM>
H2E_HD void st_w4(u32* p, u32 a, u32 b, u32 c, u32 d) {
#if defined(__CUDA_ARCH__)
    if (STREAM) __stcs(reinterpret_cast<uint4*>(p), make_uint4(a, b, c, d));
    else *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
#else
    p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}
template <int STREAM>
H2E_HD void st_w1(u32* p, u32 a) {
#if defined(__CUDA_ARCH__)
    if (STREAM) __stcs(p, a);
    else *p = a;
#else
    p[0] = a;
#endif
}
template <int STREAM>
H2E_HD void st_w8(u32* p, u32 a, u32 b, u32 c, u32 d, u32 e, u32 f, u32 g, u32 h) {
#if defined(__CUDA_ARCH__)
    if (STREAM) st256_stream(p, a, b, c, d, e, f, g, h);
    else st256(p, a, b, c, d, e, f, g, h);
#else
    u32 w[8] = {a, b, c, d, e, f, g, h};
    for (int k = 0; k < 8; k++) p[k] = w[k];
#endif
}

// Output cursor: cells are written at consecutive slots, each at its width class. `p` is uniform over the warp
// (start of the next cell's 32-lane block); a lane's words sit at p + lane * w. STREAM = 1 uses evict-first stores.
template <int STREAM>
struct OutT {
    u32* p;
    u32 lane;
    H2E_HD OutT(u32* base, u32 lane_) : p(base), lane(lane_) {}
    H2E_HD OutT at_words(u32 cells, u32 words) const {  // cursor `cells` cells = `words` words per lane further on
#if defined(H2E_WIDTH_PROBE)
        (void)words;
        return OutT(p + (size_t)cells * 8, lane);
#else
        (void)cells;
        return OutT(p + (size_t)words * TILE, lane);
#endif
    }
#if defined(H2E_WIDTH_PROBE)
    H2E_HD void c8(const u32*) { st_probe(p, 8u); p += 8; }
    H2E_HD void c4(const u32*) { st_probe(p, 4u); p += 8; }
    H2E_HD void c1(u32) { st_probe(p, 1u); p += 8; }
    H2E_HD void r8(const u32*) { st_probe(p, 8u); p += 8; }
    H2E_HD void r4(const u32*) { st_probe(p, 4u); p += 8; }
#else
    H2E_HD void c8(const u32* w) {
        st_w8<STREAM>(p + lane * 8, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
        p += 8 * TILE;
    }
    H2E_HD void c4(const u32* w) {
        st_w4<STREAM>(p + lane * 4, w[0], w[1], w[2], w[3]);
        p += 4 * TILE;
    }
    H2E_HD void c1(u32 v) {
        st_w1<STREAM>(p + lane, v);
        p += TILE;
    }
    // cells that later macro-ops read back (limb accumulators, natives): never evict-first
    H2E_HD void r8(const u32* w) {
        st_w8<0>(p + lane * 8, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
        p += 8 * TILE;
    }
    H2E_HD void r4(const u32* w) {
        st_w4<0>(p + lane * 4, w[0], w[1], w[2], w[3]);
        p += 4 * TILE;
    }
#endif
};
typedef OutT<0> Out;
typedef OutT<1> OutStream;
// cursor of the unsplit int_mul / reduce / div_core blocks (thread mode)
#if defined(H2E_STREAM_ALL)
typedef OutStream OutBulk;
#else
typedef Out OutBulk;
#endif

// ---- slot references ----
// start of the referenced cell's 32-lane block (compact) / of the cell (probe)
H2E_HD u32* ref_base(const LaneCtx& ln, u32 ref) {
#if defined(H2E_WIDTH_PROBE)
    return ln.vals + (size_t)ref * 8;
#else
    return ln.vals + (size_t)(ref >> 2) * TILE;
#endif
}
template <class O>
H2E_HD O out_at(const LaneCtx& ln, u32 ref) { return O(ref_base(ln, ref), ln.lane); }
// operand loads: the referenced cell zero-extended (or truncated: limbs stored as full cells) to 8 / 4 words.
// Device: one predicated load per width class in a single asm block -- no branches, a dozen instructions per site
// (the macro-ops are I-cache bound in team mode, and operand loads sit at the top of every one of them).
H2E_HD void ld_slot8(const LaneCtx& ln, u32 ref, u32* w) {
#if defined(H2E_WIDTH_PROBE)
    ld8(w, ref_base(ln, ref));
#elif defined(__CUDA_ARCH__)
    const u32 wc = ref & 3u;
    const u32* q = ref_base(ln, ref) + ln.lane * (wc == 2u ? 8u : (wc == 1u ? 4u : 1u));
    H2E_UNROLL
    for (int k = 0; k < 8; k++) w[k] = 0;
    asm volatile(
        "{\n\t.reg .pred p0, p1, p2;\n\t"
        "setp.eq.u32 p2, %9, 2;\n\tsetp.eq.u32 p1, %9, 1;\n\tsetp.eq.u32 p0, %9, 0;\n\t"
        "@p2 ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
        "@p1 ld.global.v4.b32 {%0,%1,%2,%3}, [%8];\n\t"
        "@p0 ld.global.b32 %0, [%8];\n\t}"
        : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7])
        : "l"(q), "r"(wc)
        : "memory");
#else
    const u32* p = ref_base(ln, ref);
    const u32 wc = ref & 3u, n = wc == 2u ? 8u : (wc == 1u ? 4u : 1u);
    for (u32 k = 0; k < 8; k++) w[k] = k < n ? p[ln.lane * n + k] : 0;
#endif
}
H2E_HD void ld_slot4(const LaneCtx& ln, u32 ref, u32* w) {
#if defined(H2E_WIDTH_PROBE)
    ld4(w, ref_base(ln, ref));
#elif defined(__CUDA_ARCH__)
    const u32 wc = ref & 3u;
    const u32* q = ref_base(ln, ref) + ln.lane * (wc == 2u ? 8u : (wc == 1u ? 4u : 1u));
    H2E_UNROLL
    for (int k = 0; k < 4; k++) w[k] = 0;
    asm volatile(
        "{\n\t.reg .pred p0;\n\t"
        "setp.eq.u32 p0, %5, 0;\n\t"
        "@!p0 ld.global.v4.b32 {%0,%1,%2,%3}, [%4];\n\t"
        "@p0 ld.global.b32 %0, [%4];\n\t}"
        : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3])
        : "l"(q), "r"(wc)
        : "memory");
#else
    const u32* p = ref_base(ln, ref);
    const u32 wc = ref & 3u, n = wc == 2u ? 8u : (wc == 1u ? 4u : 1u);
    for (u32 k = 0; k < 4; k++) w[k] = k < n ? p[ln.lane * n + k] : 0;
#endif
}
// single-cell stores / loads at (reference of a block's first cell) + (cells, words per lane) inside the block
H2E_HD u32* blk_cell(const LaneCtx& ln, u32 ref, u32 cells, u32 words, u32 w) {
#if defined(H2E_WIDTH_PROBE)
    (void)words; (void)w;
    return ref_base(ln, ref) + (size_t)cells * 8;
#else
    (void)cells;
    return ref_base(ln, ref) + (size_t)words * TILE + ln.lane * w;
#endif
}
H2E_HD void st_cell8(u32* p, const u32* w) {
#if defined(H2E_WIDTH_PROBE)
    st_probe(p, 8u);
#else
    st_w8<0>(p, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
#endif
}
H2E_HD void st_cell4(u32* p, const u32* w) {
#if defined(H2E_WIDTH_PROBE)
    st_probe(p, 4u);
#else
    st_w4<0>(p, w[0], w[1], w[2], w[3]);
#endif
}
H2E_HD void st_cell1(u32* p, u32 v) {
#if defined(H2E_WIDTH_PROBE)
    st_probe(p, 1u);
#else
    st_w1<0>(p, v);
#endif
}
H2E_HD void ld_input8(const LaneCtx& ln, u32 idx, u32* w) { ld8(w, ln.inputs + (size_t)idx * 8); }

// ------------------------------- Fr helpers (canonical form) ---------------------------------
H2E_HD void fr_add(const FrConst& F, u32* r, const u32* a, const u32* b) {
    u32 s[8], d[8];
    u32 c = bn_add<8>(s, a, b);
    u32 br = bn_sub<8>(d, s, F.r);
    bool ge = c || !br;
    H2E_UNROLL
    for (int i = 0; i < 8; i++) r[i] = ge ? d[i] : s[i];
}
H2E_HD void fr_sub(const FrConst& F, u32* r, const u32* a, const u32* b) {
    u32 d[8], e[8];
    u32 br = bn_sub<8>(d, a, b);
    bn_add<8>(e, d, F.r);
    H2E_UNROLL
    for (int i = 0; i < 8; i++) r[i] = br ? e[i] : d[i];
}
// x (NX words, < 2^512) mod r
template <int NX>
H2E_HD void fr_reduce(const FrConst& F, u32* r, const u32* x) {
    typedef Barrett<NX, 8, 254, 512> B;
    u32 q[B::NQ];
    B::divrem(x, F.r, F.mu, q, r);
}
H2E_HD void fr_mul(const FrConst& F, u32* r, const u32* a, const u32* b) {
    u32 p[16];
    bn_mul<8, 8>(p, a, b);
    fr_reduce<16>(F, r, p);
}
// out-of-line copy for the macro-ops that multiply many times (is_int_zero's batched inversion): the ~600
// instruction body stays resident in the instruction cache instead of being streamed once per call site
static H2E_HDN void fr_mul_call(const FrConst& F, u32* r, const u32* a, const u32* b) {
    fr_mul(F, r, a, b);
}
static H2E_HDN void fr_inverse(const FrConst& F, u32* r, const u32* a) {
    ModInv30<8>::inverse(r, a, F.r);
}
// signed 256-bit two's complement -> canonical Fr (|x| << r)
H2E_HD void signed_to_fr(const FrConst& F, u32* r, const u32* x) {
    u32 e[8];
    bn_add<8>(e, x, F.r);
    bool neg = (x[7] >> 31) != 0;
    H2E_UNROLL
    for (int i = 0; i < 8; i++) r[i] = neg ? e[i] : x[i];
}

// ------------------------------- range rows (range_chip.rs:270-347, context.rs:835-997) ------
H2E_HD u32 chunk18(const u32* l, int j) {
    int bit = 18 * j, w = bit >> 5, s = bit & 31;
    u32 lo = l[w] >> s;
    u32 hi = (s != 0 && w + 1 < 4) ? (l[w + 1] << (32 - s)) : 0;
    return (lo | hi) & 0x3ffffu;
}
// assign_nonleading_limb: 3-line range value, 7 cells: common v0,v1,v2; tagged v3,v4,v5; acc.
template <class O>
H2E_HD void emit_limb3(O& o, const u32* l, u32& status) {
    H2E_UNROLL
    for (int j = 0; j < 6; j++) o.c1(chunk18(l, j));
    o.r4(l);
    if ((l[3] >> 12) != 0) status |= ST_RANGE;
}
// assign_{w_ceil,d}_leading_limb: 2-line range value, 5 cells: common v0,v1; tagged v2,v3; acc.
// `dec` chunks are decomposed, the rest are the zero padding of `v.resize(4)` (context.rs:987).
template <int DEC, int BITS, class O>
H2E_HD void emit_lead2(O& o, const u32* l, u32& status) {
    H2E_UNROLL
    for (int j = 0; j < 4; j++) o.c1(j < DEC ? chunk18(l, j) : 0u);
    o.r4(l);
    u32 t[4];
    bn_shr<4, 4, BITS>(t, l);
    if (!bn_is_zero<4>(t)) status |= ST_RANGE;
}
// assign_common: 1-line range value, 2 cells: tagged v, acc v.
template <class O>
H2E_HD void emit_common(O& o, u32 v, u32& status) {
    o.c1(v);
    o.c1(v);
    if (v >> 18) status |= ST_RANGE;
}

// limbs[i] (4 words each) of x (< 2^(108*L))
template <int NXW, int L>
H2E_HD void split_limbs(u32 (*limbs)[4], const u32* x) {
    bn_shr<NXW, 4, 0>(limbs[0], x);
    bn_mask<4, 108>(limbs[0]);
    bn_shr<NXW, 4, 108>(limbs[1], x);
    bn_mask<4, 108>(limbs[1]);
    bn_shr<NXW, 4, 216>(limbs[2], x);
    bn_mask<4, 108>(limbs[2]);
    if (L > 3) {
        bn_shr<NXW, 4, 324>(limbs[L > 3 ? 3 : 0], x);
        bn_mask<4, 108>(limbs[L > 3 ? 3 : 0]);
    }
}
// x = sum limbs[i] << (108 i); limbs may be overflowed (up to 4 full words)
template <int NXW, int L>
H2E_HD void gather_limbs(u32* x, const u32 (*limbs)[4]) {
    u32 t[NXW];
    bn_shl<4, NXW, 0>(x, limbs[0]);
    bn_shl<4, NXW, 108>(t, limbs[1]);
    bn_add<NXW>(x, x, t);
    bn_shl<4, NXW, 216>(t, limbs[2]);
    bn_add<NXW>(x, x, t);
    if (L > 3) {
        bn_shl<4, NXW, 324>(t, limbs[L > 3 ? 3 : 0]);
        bn_add<NXW>(x, x, t);
    }
}

// assign_w / assign_d (integer_chip.rs:236-281): range rows for each limb, then the native row
// sum_with_constant(limbs x limb_coeffs) = [limb_0..limb_{L-1}] last(native).
// x: NXW words. Outputs limbs and native (= x mod r).
template <class T, int NXW, int LDEC, int LBITS, class O>
H2E_HD void emit_assign_int(const DeviceConsts& C, O& o, const u32* x, u32 (*limbs)[4], u32* native, u32& status) {
    split_limbs<NXW, T::L>(limbs, x);
    H2E_UNROLL
    for (int i = 0; i < T::L - 1; i++) emit_limb3(o, limbs[i], status);
    emit_lead2<LDEC, LBITS>(o, limbs[T::L - 1], status);
    {
        // the leading limb must also hold every bit of x above 108*(L-1)+LBITS
        u32 t[NXW];
        bn_shr<NXW, NXW, 108 * (T::L - 1) + LBITS>(t, x);
        if (!bn_is_zero<NXW>(t)) status |= ST_RANGE;
    }
    fr_reduce<NXW>(C.fr, native, x);
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) o.c4(limbs[i]);
    o.r8(native);
}

// Slot offsets, inside an assign_w / assign_d block, of the cells later ops read: the limb
// accumulator cells and the native cell (block = (L-1) 3-line limbs, one 2-line limb, native row).
template <class T>
struct IntBlock {
    static constexpr int SIZE = 8 * T::L - 1;    // cells
    static constexpr int SIZE_W = 14 * T::L + 6;  // words per lane: (L-1) x 10 (3-line limb) + 8 (2-line limb) + 4L + 8
    H2E_HD static constexpr int acc(int i) { return i < T::L - 1 ? 7 * i + 6 : 7 * (T::L - 1) + 4; }
    H2E_HD static constexpr int acc_w(int i) { return i < T::L - 1 ? 10 * i + 6 : 10 * (T::L - 1) + 4; }
    static constexpr int NATIVE = 8 * T::L - 2;
    static constexpr int NATIVE_W = 14 * T::L - 2;
    // block number `blk` of a run of assign blocks starting at reference `ref`
    H2E_HD static u32* acc_ptr(const LaneCtx& ln, u32 ref, int blk, int i) { return blk_cell(ln, ref, blk * SIZE + acc(i), blk * SIZE_W + acc_w(i), 4); }
    H2E_HD static u32* native_ptr(const LaneCtx& ln, u32 ref, int blk) { return blk_cell(ln, ref, blk * SIZE + NATIVE, blk * SIZE_W + NATIVE_W, 8); }
};
// assign_w / assign_d cells when limbs and native are already known
template <class T, int LDEC, int LBITS, class O>
H2E_HD void emit_assign_int_known(O& o, const u32 (*limbs)[4], const u32* native, u32& status) {
    H2E_UNROLL
    for (int i = 0; i < T::L - 1; i++) emit_limb3(o, limbs[i], status);
    emit_lead2<LDEC, LBITS>(o, limbs[T::L - 1], status);
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) o.c4(limbs[i]);
    o.c8(native);
}

// native row of a linear limb op: [s_0..s_{L-1}] last(sum s_i * 2^(108 i) mod r)
template <class T, class O>
H2E_HD void emit_native_row(const DeviceConsts& C, O& o, const u32 (*s)[4]) {
    constexpr int NXW = T::L * 4 + 2;  // 108*(L-1)+128 bits
    u32 x[NXW];
    gather_limbs<NXW, T::L>(x, s);
    u32 native[8];
    fr_reduce<NXW>(C.fr, native, x);
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) o.c4(s[i]);
    o.c8(native);
}

// ------------------------------- mul equation (integer_chip.rs:73-215) -----------------------
// Constraint rows for  a * b = d * w + rem  on limbs and on native.
template <class T, class O>
H2E_HD void emit_mul_constraints(const DeviceConsts& C, const FieldConst& fc, O& o3, const u32 (*al)[4], const u32 (*bl)[4],
                                 const u32 (*dl)[4], const u32 (*rl)[4], const u32* an, const u32* bn, const u32* dn, const u32* rn,
                                 u32& status) {
    constexpr int L = T::L, M = T::M;
    // size of the mul_add_with_next_line block: pos with n terms -> n == 1 ? 4 cells (3 limbs + 1 field element = 20 words)
    // : 4n + 1 cells (n x (3 limbs + 1 field element) + 1 field element = 20n + 8 words)
    int stage3 = 0, stage3_w = 0;
    H2E_UNROLL
    for (int pos = 0; pos < M; pos++) {
        int hi = pos + 1 < L ? pos + 1 : L, lo = pos >= L - 1 ? pos - (L - 1) : 0;
        int n = hi - lo;
        stage3 += (n == 1) ? 4 : 4 * n + 1;
        stage3_w += (n == 1) ? 20 : 20 * n + 8;
    }
    O o4 = o3.at_words(stage3, stage3_w);

    // borrow = L*B + 2 ; c0 = B*borrow = L*2^216 + 2^109 ; c1 = c0 - borrow
    u32 c0[8], c1[8];
    bn_zero<8>(c0);
    c0[6] = (u32)L << 24;   // L * 2^216
    c0[3] |= 1u << 13;      // 2^109
    {
        u32 bw[8];
        bn_zero<8>(bw);
        bw[0] = 2;
        bw[3] = (u32)L << 12;  // L * 2^108
        bn_sub<8>(c1, c0, bw);
    }

    u32 vprev[8];  // v of the previous limb (v_h * B + v_l), < 2^126
    bn_zero<8>(vprev);
    u32 vh_prev = 0;
    u32 vl_prev[4] = {0, 0, 0, 0};

    H2E_UNROLL
    for (int pos = 0; pos < M; pos++) {
        const int hi = pos + 1 < L ? pos + 1 : L, lo = pos >= L - 1 ? pos - (L - 1) : 0;
        const int n = hi - lo;
        u32 acc[8];
        bn_zero<8>(acc);
        H2E_UNROLL
        for (int i = lo; i < hi; i++) {
            // row: [a_i, b_{pos-i}, d_i] last(t_prev) (base_chip.rs:259-273); single term: mul_add row
            o3.c4(al[i]);
            o3.c4(bl[pos - i]);
            o3.c4(dl[i]);
            if (n > 1) {
                u32 t[8];
                signed_to_fr(C.fr, t, acc);
                o3.c8(t);
            }
            u32 p[8];
            bn_mul<4, 4>(p, al[i], bl[pos - i]);
            bn_add<8>(acc, acc, p);
            bn_mul<4, 4>(p, dl[i], fc.w_limbs[pos - i]);
            bn_sub<8>(acc, acc, p);
        }
        u32 lfr[8];
        signed_to_fr(C.fr, lfr, acc);
        o3.c8(lfr);  // mul_add result / tail row (base_chip.rs:276-278)

        // ---- limb check `pos` (integer_chip.rs:114-192) ----
        u32 u[8];
        if (pos < L) {
            u32 r8[8] = {rl[pos < L ? pos : 0][0], rl[pos < L ? pos : 0][1], rl[pos < L ? pos : 0][2], rl[pos < L ? pos : 0][3], 0, 0, 0, 0};
            bn_sub<8>(u, acc, r8);
        } else {
            bn_copy<8>(u, acc);
        }
        if (pos == 0) {
            bn_add<8>(u, u, c0);
        } else {
            bn_add<8>(u, u, vprev);
            bn_add<8>(u, u, c1);
        }
        // sum row
        o4.c8(lfr);
        if (pos < L) o4.c4(rl[pos < L ? pos : 0]);
        if (pos > 0) {
            o4.c1(vh_prev);
            o4.c4(vl_prev);
        }
        u32 ufr[8];
        signed_to_fr(C.fr, ufr, u);
        o4.c8(ufr);
        if (u[7] >> 31) status |= ST_NEGATIVE;
        // v = u / B exactly
        {
            u32 lowbits[4] = {u[0], u[1], u[2], u[3] & 0xfffu};
            if (!bn_is_zero<4>(lowbits)) status |= ST_NONZERO_REMAINDER;
        }
        bn_shr<8, 8, 108>(vprev, u);
        u32 vh4[4];
        bn_shr<8, 4, 108>(vh4, vprev);
        bn_shr<8, 4, 0>(vl_prev, vprev);
        bn_mask<4, 108>(vl_prev);
        vh_prev = vh4[0];
        if (vh4[1] | vh4[2] | vh4[3]) status |= ST_RANGE;
        emit_common(o4, vh_prev, status);
        emit_limb3(o4, vl_prev, status);
        // tie row [v_h : 2^216, v_l : 2^108] last(u : -1)
        o4.c1(vh_prev);
        o4.c4(vl_prev);
        o4.c8(ufr);
    }
    // native row (integer_chip.rs:195-215)
    o4.c8(an);
    o4.c8(bn);
    o4.c8(dn);
    o4.c8(rn);
    o3.p = o4.p;
}

// ------------------------------- macro-ops ---------------------------------------------------
template <class T>
H2E_HD void load_int_limbs(const LaneCtx& ln, const u32* slots, u32 (*limbs)[4]) {
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) ld_slot4(ln, slots[i], limbs[i]);
}

// OP_LOAD_INT (test/bench harness prelude): L `assign` rows for the limbs + one for the native. a0 = first input cell;
// a1 = 0: limb i in the low 16 bytes of logical input i (64 bytes each), 1: limbs packed back to back, 16 bytes each.
template <int FID>
H2E_HD void op_load_int(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    Out o = out_at<Out>(ln, in.out);
    u32 limbs[T::L][4];
    // all input loads before the first store (loads and stores are ordered asm volatile: interleaved they would pay
    // one DRAM latency per limb)
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) ld4(limbs[i], ln.inputs + (size_t)in.a[0] * 8 + (in.a[1] ? 4 * i : 16 * i));  // a1: limbs packed, 16 bytes each
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) o.c4(limbs[i]);
    constexpr int NXW = T::L * 4 + 2;
    u32 x[NXW], native[8];
    gather_limbs<NXW, T::L>(x, limbs);
    fr_reduce<NXW>(H2E_CONSTS.fr, native, x);
    o.c8(native);
}

// OP_ASSIGN_W (integer_chip.rs:236-258). Input value occupies 1 (NW=8) or 2 (NW=12) input cells.
template <int FID>
H2E_HD void op_assign_w(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    u32 x[16];
    if (in.a[1] == 0) {
        ld_input8(ln, in.a[0], x);
        if (T::NW > 8) ld_input8(ln, in.a[0] + 1, x + 8);
    } else {
        H2E_UNROLL
        for (int k = 0; k < 8; k++) x[k] = ln.cpool[(size_t)in.a[0] * 8 + k];
        if (T::NW > 8) {
            H2E_UNROLL
            for (int k = 0; k < 8; k++) x[8 + k] = ln.cpool[(size_t)(in.a[0] + 1) * 8 + k];
        }
    }
    Out o = out_at<Out>(ln, in.out);
    u32 limbs[T::L][4], native[8];
    emit_assign_int<T, T::NW, T::WDEC, T::WLEAD>(H2E_CONSTS, o, x, limbs, native, ln.status);
}

// OP_ASSIGN_INT_CONST (integer_chip.rs:580-598): assign_constant rows for each limb and the native.
template <int FID>
H2E_HD void op_assign_int_const(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    u32 x[16];
    if (in.a[0] == 0) {
        ld_input8(ln, in.a[1], x);
        if (T::NW > 8) ld_input8(ln, in.a[1] + 1, x + 8);
    } else {
        H2E_UNROLL
        for (int k = 0; k < 8; k++) x[k] = ln.cpool[(size_t)in.a[1] * 8 + k];
        if (T::NW > 8) {
            H2E_UNROLL
            for (int k = 0; k < 8; k++) x[8 + k] = ln.cpool[(size_t)(in.a[1] + 1) * 8 + k];
        }
    }
    Out o = out_at<Out>(ln, in.out);
    u32 limbs[T::L][4], native[8];
    split_limbs<T::NW, T::L>(limbs, x);
    fr_reduce<T::NW>(H2E_CONSTS.fr, native, x);
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) o.c4(limbs[i]);
    o.c8(native);
}

// OP_INT_ADD / OP_INT_SUB / OP_INT_NEG / OP_MUL_SMALL (integer_chip.rs:384-464, 618-658)
// kind: 0 add, 1 sub, 2 neg, 3 mul-small
// The native cell of the result (a sum_with_constant row over the new limbs, integer_chip.rs:397-399)
// equals the same linear combination of the operands' natives mod r, because every AssignedInteger
// satisfies native = sum limb_i * 2^(108 i) mod r; that is one Fr add/sub instead of a wide reduction.
// All operand loads are issued before the first store (loads and stores are ordered asm volatile).
template <int FID, int KIND>
H2E_HD void op_int_linear(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const FieldConst& fc = H2E_CONSTS.f[FID];
    const FrConst& F = H2E_CONSTS.fr;
    Out o = out_at<Out>(ln, in.out);
    u32 al[T::L][4], bl[T::L][4], an[8], bn[8];
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) {
        ld_slot4(ln, in.a[i], al[i]);
        if (KIND <= 1) ld_slot4(ln, in.a[T::L + i], bl[i]);
    }
    u32 native[8];
    if (KIND == 0) {
        ld_slot8(ln, in.a[2 * T::L], an);
        ld_slot8(ln, in.a[2 * T::L + 1], bn);
        fr_add(F, native, an, bn);
    } else if (KIND == 1) {
        u32 t[8];
        ld_slot8(ln, in.a[2 * T::L + 1], an);
        ld_slot8(ln, in.a[2 * T::L + 2], bn);
        fr_sub(F, t, an, bn);
        fr_add(F, native, t, fc.upper_native[in.a[2 * T::L] & 63]);
    } else if (KIND == 2) {
        ld_slot8(ln, in.a[T::L + 1], an);
        fr_sub(F, native, fc.upper_native[in.a[T::L] & 63], an);
    } else {
        u32 p[9];
        ld_slot8(ln, in.a[T::L + 1], an);
        u32 k[1] = {in.a[T::L]};
        bn_mul<8, 1>(p, an, k);
        fr_reduce<9>(F, native, p);
    }
    u32 s[T::L][4];
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) {
        const u32* a = al[i];
        const u32* b = bl[i];
        if (KIND == 0) {
            o.c4(a);
            o.c4(b);
            if (bn_add<4>(s[i], a, b)) ln.status |= ST_RANGE;
        } else if (KIND == 1) {
            o.c4(a);
            o.c4(b);
            u32 t[4];
            bn_add<4>(t, a, fc.upper[in.a[2 * T::L] & 63][i]);
            if (bn_sub<4>(s[i], t, b)) ln.status |= ST_NEGATIVE;
        } else if (KIND == 2) {
            o.c4(a);
            if (bn_sub<4>(s[i], fc.upper[in.a[T::L] & 63][i], a)) ln.status |= ST_NEGATIVE;
        } else {
            o.c4(a);
            u32 k[1] = {in.a[T::L]};
            u32 p[5];
            bn_mul<4, 1>(p, a, k);
            if (p[4]) ln.status |= ST_RANGE;
            H2E_UNROLL
            for (int k2 = 0; k2 < 4; k2++) s[i][k2] = p[k2];
        }
        o.c4(s[i]);
    }
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) o.c4(s[i]);
    o.c8(native);
}

// OP_REDUCE (integer_chip.rs:283-373). Rows after assign_w(rem): assign_common(d), the native row
// [d : w_native, rem.native : 1] last(a.native : -1), then R limb rows.
template <class T, class O>
H2E_HD void emit_reduce_rest(const FieldConst& fc, O& o, const u32 (*al)[4], const u32* an, const u32 (*rl)[4], const u32* rn, u32 d,
                             u32& status) {
    emit_common(o, d, status);
    o.c1(d);
    o.c8(rn);
    o.c8(an);
    // limb rows: u_i = d*w_i + rem_i + 64*B - a_i + v_{i-1} - (i ? 64 : 0); v_i = u_i / B
    u32 vprev[4] = {0, 0, 0, 0};
    H2E_UNROLL
    for (int i = 0; i < T::R; i++) {
        u32 u[8];
        u32 d1[1] = {d};
        u32 p[5];
        bn_mul<4, 1>(p, fc.w_limbs[i], d1);
        u[0] = p[0]; u[1] = p[1]; u[2] = p[2]; u[3] = p[3]; u[4] = p[4]; u[5] = 0; u[6] = 0; u[7] = 0;
        u32 t8[8] = {rl[i][0], rl[i][1], rl[i][2], rl[i][3], 0, 0, 0, 0};
        bn_add<8>(u, u, t8);
        u32 k8[8] = {0, 0, 0, 64u << 12, 0, 0, 0, 0};  // 64 * 2^108
        bn_add<8>(u, u, k8);
        u32 a8[8] = {al[i][0], al[i][1], al[i][2], al[i][3], 0, 0, 0, 0};
        if (bn_sub<8>(u, u, a8)) status |= ST_NEGATIVE;
        if (i > 0) {
            u32 v8[8] = {vprev[0], vprev[1], vprev[2], vprev[3], 0, 0, 0, 0};
            bn_add<8>(u, u, v8);
            u32 b8[8] = {64, 0, 0, 0, 0, 0, 0, 0};
            if (bn_sub<8>(u, u, b8)) status |= ST_NEGATIVE;
        }
        {
            u32 lowbits[4] = {u[0], u[1], u[2], u[3] & 0xfffu};
            if (!bn_is_zero<4>(lowbits)) status |= ST_NONZERO_REMAINDER;
        }
        u32 v[4];
        bn_shr<8, 4, 108>(v, u);
        u32 vlast[4] = {vprev[0], vprev[1], vprev[2], vprev[3]};
        emit_limb3(o, v, status);
        // row [d : w_i, rem_i : 1, a_i : -1, (v_{i-1} : 1 | 0 : 0)] last(v_i : -B) const
        o.c1(d);
        o.c4(rl[i]);
        o.c4(al[i]);
        o.c4(vlast);  // raw 0 on the first row
        o.c4(v);
        H2E_UNROLL
        for (int k = 0; k < 4; k++) vprev[k] = v[k];
    }
}
template <int FID>
H2E_HDN void op_reduce(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    u32 al[T::L][4], an[8];
    load_int_limbs<T>(ln, in.a, al);
    ld_slot8(ln, in.a[T::L], an);
    u32 x[T::NXA];
    gather_limbs<T::NXA, T::L>(x, al);
    typedef Barrett<T::NXA, T::NW, T::NBITS, T::KBITS> B;
    u32 q[B::NQ], rem[T::NW];
    B::divrem(x, fc.w, fc.mu, q, rem);
    OutBulk o = out_at<OutBulk>(ln, in.out);
    u32 rl[T::L][4], rn[8];
    emit_assign_int<T, T::NW, T::WDEC, T::WLEAD>(C, o, rem, rl, rn, ln.status);
    u32 d = q[0];
    {
        u32 hi = 0;
        H2E_UNROLL
        for (int i = 1; i < B::NQ; i++) hi |= q[i];
        if (hi) ln.status |= ST_RANGE;
    }
    emit_reduce_rest<T>(fc, o, al, an, rl, rn, d, ln.status);
}
// Team-mode split of OP_REDUCE: HEAD stores the cells later ops read (limb accumulators and native
// of rem) plus the quotient cell for the TAIL; TAIL re-reads them and writes the whole block.
template <int FID>
H2E_HDN void op_reduce_head(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    u32 al[T::L][4];
    load_int_limbs<T>(ln, in.a, al);
    u32 x[T::NXA];
    gather_limbs<T::NXA, T::L>(x, al);
    typedef Barrett<T::NXA, T::NW, T::NBITS, T::KBITS> B;
    u32 q[B::NQ], rem[T::NW];
    B::divrem(x, fc.w, fc.mu, q, rem);
    u32 rl[T::L][4], rn[8];
    split_limbs<T::NW, T::L>(rl, rem);
    fr_reduce<T::NW>(C.fr, rn, rem);
    {
        u32 hi = 0;
        H2E_UNROLL
        for (int i = 1; i < B::NQ; i++) hi |= q[i];
        if (hi) ln.status |= ST_RANGE;
    }
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) st_cell4(IntBlock<T>::acc_ptr(ln, in.out, 0, i), rl[i]);
    st_cell8(IntBlock<T>::native_ptr(ln, in.out, 0), rn);
    st_cell1(blk_cell(ln, in.out, IntBlock<T>::SIZE, IntBlock<T>::SIZE_W, 1), q[0]);  // first cell of assign_common(d)
}
template <int FID>
H2E_HDN void op_reduce_tail(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const FieldConst& fc = H2E_CONSTS.f[FID];
    u32 al[T::L][4], an[8], rl[T::L][4], rn[8], dw[4];
    load_int_limbs<T>(ln, in.a, al);
    ld_slot8(ln, in.a[T::L], an);
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) ld4(rl[i], IntBlock<T>::acc_ptr(ln, in.out, 0, i));
    ld8(rn, IntBlock<T>::native_ptr(ln, in.out, 0));
    dw[0] = *blk_cell(ln, in.out, IntBlock<T>::SIZE, IntBlock<T>::SIZE_W, 1);
    OutStream o = out_at<OutStream>(ln, in.out);
    emit_assign_int_known<T, T::WDEC, T::WLEAD>(o, rl, rn, ln.status);
    emit_reduce_rest<T>(fc, o, al, an, rl, rn, dw[0], ln.status);
}

// OP_INT_MUL (integer_chip.rs:466-483)
template <int FID>
H2E_HDN void op_int_mul(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    constexpr int L = T::L;
    u32 al[L][4], bl[L][4], an[8], bn[8];
    load_int_limbs<T>(ln, in.a, al);
    ld_slot8(ln, in.a[L], an);
    load_int_limbs<T>(ln, in.a + L + 1, bl);
    ld_slot8(ln, in.a[2 * L + 1], bn);
    u32 q[T::ND], rem[T::NW];
    {
        u32 xa[T::NXA], xb[T::NXA];
        gather_limbs<T::NXA, L>(xa, al);
        gather_limbs<T::NXA, L>(xb, bl);
        u32 x[2 * T::NXA];
        bn_mul<T::NXA, T::NXA>(x, xa, xb);
        typedef Barrett<2 * T::NXA, T::NW, T::NBITS, T::KBITS> B;
        static_assert(B::NQ == T::ND, "quotient width");
        B::divrem(x, fc.w, fc.mu, q, rem);
    }
    OutBulk o = out_at<OutBulk>(ln, in.out);
    u32 rl[L][4], rn[8], dl[L][4], dn[8];
    emit_assign_int<T, T::NW, T::WDEC, T::WLEAD>(C, o, rem, rl, rn, ln.status);
    emit_assign_int<T, T::ND, T::DDEC, T::DLEAD>(C, o, q, dl, dn, ln.status);
    emit_mul_constraints<T>(C, fc, o, al, bl, dl, rl, an, bn, dn, rn, ln.status);
}

// Team-mode split of OP_INT_MUL. HEAD computes quotient and remainder and stores only the cells
// other macro-ops can depend on (limb accumulators and natives of rem, and of d for the TAIL);
// TAIL re-reads them and writes every cell of the block (range chunks, copies, constraint rows).
// Only HEAD sits on the program's critical path.
template <int FID>
H2E_HDN void op_int_mul_head(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    constexpr int L = T::L;
    u32 al[L][4], bl[L][4];
    load_int_limbs<T>(ln, in.a, al);
    load_int_limbs<T>(ln, in.a + L + 1, bl);
    u32 q[T::ND], rem[T::NW];
    {
        u32 xa[T::NXA], xb[T::NXA];
        gather_limbs<T::NXA, L>(xa, al);
        gather_limbs<T::NXA, L>(xb, bl);
        u32 x[2 * T::NXA];
        bn_mul<T::NXA, T::NXA>(x, xa, xb);
        typedef Barrett<2 * T::NXA, T::NW, T::NBITS, T::KBITS> B;
        B::divrem(x, fc.w, fc.mu, q, rem);
    }
    u32 limbs[L][4], native[8];
    split_limbs<T::NW, L>(limbs, rem);
    fr_reduce<T::NW>(C.fr, native, rem);
    H2E_UNROLL
    for (int i = 0; i < L; i++) st_cell4(IntBlock<T>::acc_ptr(ln, in.out, 0, i), limbs[i]);
    st_cell8(IntBlock<T>::native_ptr(ln, in.out, 0), native);
    split_limbs<T::ND, L>(limbs, q);
    fr_reduce<T::ND>(C.fr, native, q);
    H2E_UNROLL
    for (int i = 0; i < L; i++) st_cell4(IntBlock<T>::acc_ptr(ln, in.out, 1, i), limbs[i]);
    st_cell8(IntBlock<T>::native_ptr(ln, in.out, 1), native);
}
template <int FID>
H2E_HDN void op_int_mul_tail(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    constexpr int L = T::L;
    u32 al[L][4], bl[L][4], an[8], bn[8];
    load_int_limbs<T>(ln, in.a, al);
    ld_slot8(ln, in.a[L], an);
    load_int_limbs<T>(ln, in.a + L + 1, bl);
    ld_slot8(ln, in.a[2 * L + 1], bn);
    u32 rl[L][4], rn[8], dl[L][4], dn[8];
    H2E_UNROLL
    for (int i = 0; i < L; i++) {
        ld4(rl[i], IntBlock<T>::acc_ptr(ln, in.out, 0, i));
        ld4(dl[i], IntBlock<T>::acc_ptr(ln, in.out, 1, i));
    }
    ld8(rn, IntBlock<T>::native_ptr(ln, in.out, 0));
    ld8(dn, IntBlock<T>::native_ptr(ln, in.out, 1));
    // in.flags: bit 0 = the two assign blocks (range chunks, copies), bit 1 = the constraint rows;
    // the scheduler issues them as two instructions so that neither is longer than the HEAD
    if (in.flags & 1) {
        OutStream o = out_at<OutStream>(ln, in.out);
        emit_assign_int_known<T, T::WDEC, T::WLEAD>(o, rl, rn, ln.status);
        emit_assign_int_known<T, T::DDEC, T::DLEAD>(o, dl, dn, ln.status);
    }
    if (in.flags & 2) {
        OutStream o = out_at<OutStream>(ln, in.out).at_words(2 * IntBlock<T>::SIZE, 2 * IntBlock<T>::SIZE_W);
        emit_mul_constraints<T>(C, fc, o, al, bl, dl, rl, an, bn, dn, rn, ln.status);
    }
}

// OP_DIV_CORE (integer_chip.rs:522-535): c = a / b in W (0 if b == 0), d = (b*c - a) / w, then
// the mul equation b * c = d * w + a. HAVE_INV: b^-1 mod w was computed by OP_DIV_INV (team mode runs
// it concurrently with is_int_zero(b), which only shares the operand b) and is read from scratch.
static const int SCRATCH_STRIDE = TILE * 16;  // words between consecutive scratch entries of one lane
// a * b mod w (a, b < w); out of line: the merged inversion below multiplies up to six times
template <int FID>
H2E_HDN void w_mulmod(u32* r, const u32* a, const u32* b) {
    typedef FT<FID> T;
    const FieldConst& fc = H2E_CONSTS.f[FID];
    constexpr int NW = T::NW;
    u32 p[2 * NW];
    bn_mul<NW, NW>(p, a, b);
    typedef Barrett<2 * NW, NW, T::NBITS, T::KBITS> B2;
    u32 q2[B2::NQ];
    B2::divrem(p, fc.w, fc.mu, q2, r);
}
// OP_DIV_INV: b_j^-1 mod w for K = flags & 3 denominators (the int_divs of one dependency level, merged by the scheduler)
// with ONE inversion: the product of the (non-zero) denominators is inverted and each inverse recovered with 3 products
// (Montgomery's trick). A zero denominator takes no part in the product and gets the inverse 0 (integer_chip.rs:527: c = 0).
template <int FID>
H2E_HDN void op_div_inv(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const FieldConst& fc = H2E_CONSTS.f[FID];
    constexpr int L = T::L, NW = T::NW, KMAX = 14 / (L + 1);
    const int K = (in.flags & 3u) ? (int)(in.flags & 3u) : 1;
    u32 bm[KMAX][NW], pre[KMAX][NW];
    bool z[KMAX];
    H2E_UNROLL
    for (int j = 0; j < KMAX; j++) {
        z[j] = true;
        if (j < K) {
            u32 bl[L][4];
            load_int_limbs<T>(ln, in.a + j * (L + 1), bl);
            u32 xb[T::NXA];
            gather_limbs<T::NXA, L>(xb, bl);
            typedef Barrett<T::NXA, NW, T::NBITS, T::KBITS> B1;
            u32 q1[B1::NQ];
            B1::divrem(xb, fc.w, fc.mu, q1, bm[j]);
            z[j] = bn_is_zero<NW>(bm[j]);
        }
        H2E_UNROLL
        for (int w = 0; w < NW; w++) bm[j][w] = z[j] ? (w == 0 ? 1u : 0u) : bm[j][w];
        if (j == 0) {
            bn_copy<NW>(pre[0], bm[0]);
        } else if (j < K) {
            w_mulmod<FID>(pre[j], pre[j - 1], bm[j]);
        } else {
            bn_copy<NW>(pre[j], pre[j - 1]);
        }
    }
    u32 acc[NW];
    ModInv30<NW>::inverse(acc, pre[KMAX - 1], fc.w);
    H2E_UNROLL
    for (int j = KMAX - 1; j >= 0; j--) {
        if (j >= K) continue;
        u32 inv[16];
        if (j > 0) {
            w_mulmod<FID>(inv, acc, pre[j - 1]);
            u32 t[NW];
            w_mulmod<FID>(t, acc, bm[j]);
            bn_copy<NW>(acc, t);
        } else {
            bn_copy<NW>(inv, acc);
        }
        H2E_UNROLL
        for (int i = 0; i < 16; i++) inv[i] = (i < NW && !z[j]) ? inv[i] : 0;
        u32* sp = ln.scratch + (size_t)in.a[j * (L + 1) + L] * SCRATCH_STRIDE;
        st_raw8(sp, inv);
        st_raw8(sp + 8, inv + 8);
    }
}
template <int FID, bool HAVE_INV>
H2E_HD void div_core_body(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    constexpr int L = T::L, NW = T::NW;
    u32 al[L][4], bl[L][4], an[8], bn[8];
    load_int_limbs<T>(ln, in.a, al);
    ld_slot8(ln, in.a[L], an);
    load_int_limbs<T>(ln, in.a + L + 1, bl);
    ld_slot8(ln, in.a[2 * L + 1], bn);
    u32 binv[16];
    if (HAVE_INV) {
        const u32* sp = ln.scratch + (size_t)in.a[2 * L + 2] * SCRATCH_STRIDE;
        ld8(binv, sp);
        ld8(binv + 8, sp + 8);
    }
    u32 xa[T::NXA], xb[T::NXA];
    gather_limbs<T::NXA, L>(xa, al);
    gather_limbs<T::NXA, L>(xb, bl);
    typedef Barrett<T::NXA, NW, T::NBITS, T::KBITS> B1;
    u32 c[NW];
    {
        u32 q1[B1::NQ], am[NW];
        B1::divrem(xa, fc.w, fc.mu, q1, am);
        if (!HAVE_INV) {
            u32 bm[NW];
            B1::divrem(xb, fc.w, fc.mu, q1, bm);
            ModInv30<NW>::inverse(binv, bm, fc.w);
        }
        // c = am * binv mod w
        u32 p[2 * NW];
        bn_mul<NW, NW>(p, am, binv);
        typedef Barrett<2 * NW, NW, T::NBITS, T::KBITS> B2;
        u32 q2[B2::NQ];
        B2::divrem(p, fc.w, fc.mu, q2, c);
    }
    // d = (b_bn * c - a_bn) / w
    u32 q[T::ND];
    {
        constexpr int NT = T::NXA + NW;
        u32 t[NT];
        bn_mul<T::NXA, NW>(t, xb, c);
        u32 a_ext[NT];
        H2E_UNROLL
        for (int i = 0; i < NT; i++) a_ext[i] = i < T::NXA ? xa[i] : 0;
        if (bn_sub<NT>(t, t, a_ext)) ln.status |= ST_NEGATIVE;
        typedef Barrett<NT, NW, T::NBITS, T::KBITS> B3;
        u32 q3[B3::NQ], rem3[NW];
        B3::divrem(t, fc.w, fc.mu, q3, rem3);
        H2E_UNROLL
        for (int i = 0; i < T::ND; i++) q[i] = i < B3::NQ ? q3[i] : 0;
    }
    OutBulk o = out_at<OutBulk>(ln, in.out);
    u32 cl[L][4], cn[8], dl[L][4], dn[8];
    emit_assign_int<T, NW, T::WDEC, T::WLEAD>(C, o, c, cl, cn, ln.status);
    emit_assign_int<T, T::ND, T::DDEC, T::DLEAD>(C, o, q, dl, dn, ln.status);
    emit_mul_constraints<T>(C, fc, o, bl, cl, dl, al, bn, cn, dn, an, ln.status);
}
template <int FID>
H2E_HDN void op_div_core(LaneCtx& ln, const Instr& in) {
    div_core_body<FID, false>(ln, in);
}
template <int FID>
H2E_HDN void op_div_core_s(LaneCtx& ln, const Instr& in) {
    div_core_body<FID, true>(ln, in);
}
// Team-mode split of OP_DIV_CORE_S. HEAD: c = a * b^-1 mod w, stored only where later macro-ops read it
// (limb accumulators and native of the c block); TAIL re-reads c, computes d = (b*c - a) / w and writes
// every cell of the block (range chunks of c and d, copies, the constraint rows). Only HEAD sits on the
// critical path of a chain of point additions.
template <int FID>
H2E_HDN void op_div_head_s(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    constexpr int L = T::L, NW = T::NW;
    u32 al[L][4];
    load_int_limbs<T>(ln, in.a, al);
    u32 binv[16];
    const u32* sp = ln.scratch + (size_t)in.a[2 * L + 2] * SCRATCH_STRIDE;
    ld8(binv, sp);
    ld8(binv + 8, sp + 8);
    u32 xa[T::NXA];
    gather_limbs<T::NXA, L>(xa, al);
    u32 c[NW];
    {
        typedef Barrett<T::NXA, NW, T::NBITS, T::KBITS> B1;
        u32 q1[B1::NQ], am[NW];
        B1::divrem(xa, fc.w, fc.mu, q1, am);
        u32 p[2 * NW];
        bn_mul<NW, NW>(p, am, binv);
        typedef Barrett<2 * NW, NW, T::NBITS, T::KBITS> B2;
        u32 q2[B2::NQ];
        B2::divrem(p, fc.w, fc.mu, q2, c);
    }
    u32 limbs[L][4], native[8];
    split_limbs<NW, L>(limbs, c);
    fr_reduce<NW>(C.fr, native, c);
    H2E_UNROLL
    for (int i = 0; i < L; i++) st_cell4(IntBlock<T>::acc_ptr(ln, in.out, 0, i), limbs[i]);
    st_cell8(IntBlock<T>::native_ptr(ln, in.out, 0), native);
}
template <int FID>
H2E_HDN void op_div_tail(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    constexpr int L = T::L, NW = T::NW;
    u32 al[L][4], bl[L][4], an[8], bn[8];
    load_int_limbs<T>(ln, in.a, al);
    ld_slot8(ln, in.a[L], an);
    load_int_limbs<T>(ln, in.a + L + 1, bl);
    ld_slot8(ln, in.a[2 * L + 1], bn);
    u32 cl[L][4], cn[8];
    H2E_UNROLL
    for (int i = 0; i < L; i++) ld4(cl[i], IntBlock<T>::acc_ptr(ln, in.out, 0, i));
    ld8(cn, IntBlock<T>::native_ptr(ln, in.out, 0));
    u32 xa[T::NXA], xb[T::NXA], xc[T::NXA], c[NW];
    gather_limbs<T::NXA, L>(xa, al);
    gather_limbs<T::NXA, L>(xb, bl);
    gather_limbs<T::NXA, L>(xc, cl);
    H2E_UNROLL
    for (int i = 0; i < NW; i++) c[i] = xc[i];
    // d = (b_bn * c - a_bn) / w
    u32 q[T::ND];
    {
        constexpr int NT = T::NXA + NW;
        u32 t[NT];
        bn_mul<T::NXA, NW>(t, xb, c);
        u32 a_ext[NT];
        H2E_UNROLL
        for (int i = 0; i < NT; i++) a_ext[i] = i < T::NXA ? xa[i] : 0;
        if (bn_sub<NT>(t, t, a_ext)) ln.status |= ST_NEGATIVE;
        typedef Barrett<NT, NW, T::NBITS, T::KBITS> B3;
        u32 q3[B3::NQ], rem3[NW];
        B3::divrem(t, fc.w, fc.mu, q3, rem3);
        H2E_UNROLL
        for (int i = 0; i < T::ND; i++) q[i] = i < B3::NQ ? q3[i] : 0;
    }
    OutStream o = out_at<OutStream>(ln, in.out);
    u32 dl[L][4], dn[8];
    emit_assign_int_known<T, T::WDEC, T::WLEAD>(o, cl, cn, ln.status);
    emit_assign_int<T, T::ND, T::DDEC, T::DLEAD>(C, o, q, dl, dn, ln.status);
    emit_mul_constraints<T>(C, fc, o, bl, cl, dl, al, bn, cn, dn, an, ln.status);
}

// is_zero / invert rows (base_chip.rs:298-325) given a and its inverse (0 for a = 0):
// [a, c] then [a, b] last(c) with c = 1 - a*b. Returns the condition (0/1).
template <class O>
H2E_HD u32 emit_is_zero_rows(O& o, const u32* a, const u32* inv) {
    u32 c = bn_is_zero<8>(a) ? 1u : 0u;
    o.c8(a);
    o.c1(c);
    o.c8(a);
    o.c8(inv);
    o.c1(c);
    return c;
}
template <class O>
H2E_HD u32 emit_is_zero(const DeviceConsts& C, O& o, const u32* a) {
    // Every lane calls the inversion: its loop votes with a full-warp mask, and lanes are different circuit
    // instances (a tile may hold zero and non-zero values side by side). inverse(0) = 0, which is what
    // invert().unwrap_or(0) yields (base_chip.rs:301).
    u32 inv[8];
    bool z = bn_is_zero<8>(a);
    fr_inverse(C.fr, inv, a);
    u32 c = z ? 1u : 0u;
    o.c8(a);
    o.c1(c);
    o.c8(a);
    o.c8(inv);
    o.c1(c);
    return c;
}

// OP_IS_INT_ZERO on a reduced integer (integer_chip.rs:540-578)
template <int FID, class O>
H2E_HD void is_int_zero_body(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    constexpr int L = T::L;
    u32 al[L][4], an[8];
    load_int_limbs<T>(ln, in.a, al);
    ld_slot8(ln, in.a[L], an);
    O o = out_at<O>(ln, in.out);
    // values whose inverses the rows need: sum of limbs, native - w_native, limb_i - w_i (i < P)
    constexpr int K = 2 + T::P;
    u32 val[K][8], inv[K][8];
    bn_zero<8>(val[0]);
    H2E_UNROLL
    for (int i = 0; i < L; i++) {
        u32 t[8] = {al[i][0], al[i][1], al[i][2], al[i][3], 0, 0, 0, 0};
        bn_add<8>(val[0], val[0], t);
    }
    fr_add(C.fr, val[1], an, fc.neg_w_native);
    H2E_UNROLL
    for (int i = 0; i < T::P; i++) {
        u32 t[8] = {al[i][0], al[i][1], al[i][2], al[i][3], 0, 0, 0, 0};
        fr_add(C.fr, val[2 + i], t, fc.neg_w_limbs[i]);
    }
    // one inversion for all K values (Montgomery's trick); zeros are replaced by 1 and their
    // "inverse" forced back to 0, which is what invert().unwrap_or(0) yields (base_chip.rs:301)
    {
        bool z[K];
        u32 nz[K][8], pre[K][8];
        H2E_UNROLL
        for (int k = 0; k < K; k++) {
            z[k] = bn_is_zero<8>(val[k]);
            H2E_UNROLL
            for (int w = 0; w < 8; w++) nz[k][w] = z[k] ? (w == 0 ? 1u : 0u) : val[k][w];
        }
        bn_copy<8>(pre[0], nz[0]);
        H2E_UNROLL
        for (int k = 1; k < K; k++) fr_mul(C.fr, pre[k], pre[k - 1], nz[k]);
        u32 acc[8];
        fr_inverse(C.fr, acc, pre[K - 1]);
        H2E_UNROLL
        for (int k = K - 1; k >= 1; k--) {
            u32 t[8];
            fr_mul(C.fr, t, acc, pre[k - 1]);
            fr_mul(C.fr, acc, acc, nz[k]);
            H2E_UNROLL
            for (int w = 0; w < 8; w++) inv[k][w] = z[k] ? 0u : t[w];
        }
        H2E_UNROLL
        for (int w = 0; w < 8; w++) inv[0][w] = z[0] ? 0u : acc[w];
    }
    // is_pure_zero: sum row + is_zero rows
    H2E_UNROLL
    for (int i = 0; i < L; i++) o.c4(al[i]);
    o.c8(val[0]);
    u32 is_zero = emit_is_zero_rows(o, val[0], inv[0]);
    // is_pure_w_modulus
    o.c8(an);
    o.c8(val[1]);
    u32 is_eq = emit_is_zero_rows(o, val[1], inv[1]);
    H2E_UNROLL
    for (int i = 0; i < T::P; i++) {
        o.c4(al[i]);
        o.c8(val[2 + i]);
        u32 is_limb_eq = emit_is_zero_rows(o, val[2 + i], inv[2 + i]);
        o.c1(is_eq);
        o.c1(is_limb_eq);
        is_eq = is_eq & is_limb_eq;
        o.c1(is_eq);
    }
    // or
    o.c1(is_zero);
    o.c1(is_eq);
    o.c1(is_zero | is_eq);
}
template <int FID>
H2E_HDN void op_is_int_zero(LaneCtx& ln, const Instr& in) {
    is_int_zero_body<FID, Out>(ln, in);
}
// Team-mode split of OP_IS_INT_ZERO. Later macro-ops only read the condition cell (the last cell of the
// block), and the condition needs no inversion: HEAD computes it from zero tests and stores that one cell
// (slot a[13]); the Fr inversions only fill record cells, so the TAIL -- the whole block -- is deferred work.
template <int FID>
H2E_HDN void op_is_int_zero_head(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    constexpr int L = T::L;
    u32 al[L][4], an[8];
    load_int_limbs<T>(ln, in.a, al);
    ld_slot8(ln, in.a[L], an);
    u32 sum[8], t8[8];
    bn_zero<8>(sum);
    H2E_UNROLL
    for (int i = 0; i < L; i++) {
        u32 t[8] = {al[i][0], al[i][1], al[i][2], al[i][3], 0, 0, 0, 0};
        bn_add<8>(sum, sum, t);
    }
    u32 is_zero = bn_is_zero<8>(sum) ? 1u : 0u;
    fr_add(C.fr, t8, an, fc.neg_w_native);
    u32 is_eq = bn_is_zero<8>(t8) ? 1u : 0u;
    H2E_UNROLL
    for (int i = 0; i < T::P; i++) {
        u32 t[8] = {al[i][0], al[i][1], al[i][2], al[i][3], 0, 0, 0, 0};
        fr_add(C.fr, t8, t, fc.neg_w_limbs[i]);
        is_eq &= bn_is_zero<8>(t8) ? 1u : 0u;
    }
    st_cell1(blk_cell(ln, in.a[13], 0, 0, 1), is_zero | is_eq);
}
// TAIL: the whole block of K = flags & 3 is_int_zero calls (operands of call j at a[j(L+1) ..], first slot of
// its block in in.out / a[11 + j]) with one Fr inversion for all K * (2 + P) values.
template <int FID>
H2E_HDN void op_is_int_zero_tail(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    const DeviceConsts& C = H2E_CONSTS;
    const FieldConst& fc = C.f[FID];
    constexpr int L = T::L, NV = 2 + T::P, KMAX = L == 3 ? 3 : 2;
    const int K = in.flags & 3;
    u32 val[KMAX * NV][8], pre[KMAX * NV][8];  // indexed at run time: lives in local memory (this is deferred work)
    u32 acc[8] = {1u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j = 0; j < K; j++) {
        u32 al[L][4], an[8];
        load_int_limbs<T>(ln, in.a + j * (L + 1), al);
        ld_slot8(ln, in.a[j * (L + 1) + L], an);
        u32 v[NV][8];
        bn_zero<8>(v[0]);
        H2E_UNROLL
        for (int i = 0; i < L; i++) {
            u32 t[8] = {al[i][0], al[i][1], al[i][2], al[i][3], 0, 0, 0, 0};
            bn_add<8>(v[0], v[0], t);
        }
        fr_add(C.fr, v[1], an, fc.neg_w_native);
        H2E_UNROLL
        for (int i = 0; i < T::P; i++) {
            u32 t[8] = {al[i][0], al[i][1], al[i][2], al[i][3], 0, 0, 0, 0};
            fr_add(C.fr, v[2 + i], t, fc.neg_w_limbs[i]);
        }
        H2E_UNROLL
        for (int k = 0; k < NV; k++) {
            // zeros are replaced by 1 in the product; their "inverse" is forced back to 0 below
            const bool z = bn_is_zero<8>(v[k]);
            u32 nz[8];
            H2E_UNROLL
            for (int w = 0; w < 8; w++) {
                nz[w] = z ? (w == 0 ? 1u : 0u) : v[k][w];
                val[j * NV + k][w] = v[k][w];
                pre[j * NV + k][w] = acc[w];
            }
            fr_mul_call(C.fr, acc, acc, nz);
        }
    }
    u32 ia[8];
    fr_inverse(C.fr, ia, acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j = K - 1; j >= 0; j--) {
        u32 v[NV][8], inv[NV][8];
        H2E_UNROLL
        for (int k = NV - 1; k >= 0; k--) {
            u32 pk[8], nz[8];
            H2E_UNROLL
            for (int w = 0; w < 8; w++) {
                v[k][w] = val[j * NV + k][w];
                pk[w] = pre[j * NV + k][w];
            }
            const bool z = bn_is_zero<8>(v[k]);
            H2E_UNROLL
            for (int w = 0; w < 8; w++) nz[w] = z ? (w == 0 ? 1u : 0u) : v[k][w];
            u32 t[8];
            fr_mul_call(C.fr, t, ia, pk);
            fr_mul_call(C.fr, ia, ia, nz);
            H2E_UNROLL
            for (int w = 0; w < 8; w++) inv[k][w] = z ? 0u : t[w];
        }
        u32 al[L][4], an[8];
        load_int_limbs<T>(ln, in.a + j * (L + 1), al);
        ld_slot8(ln, in.a[j * (L + 1) + L], an);
        OutStream o = out_at<OutStream>(ln, j == 0 ? in.out : in.a[11 + j]);
        // is_pure_zero: sum row + is_zero rows
        H2E_UNROLL
        for (int i = 0; i < L; i++) o.c4(al[i]);
        o.c8(v[0]);
        u32 is_zero = emit_is_zero_rows(o, v[0], inv[0]);
        // is_pure_w_modulus
        o.c8(an);
        o.c8(v[1]);
        u32 is_eq = emit_is_zero_rows(o, v[1], inv[1]);
        H2E_UNROLL
        for (int i = 0; i < T::P; i++) {
            o.c4(al[i]);
            o.c8(v[2 + i]);
            u32 is_limb_eq = emit_is_zero_rows(o, v[2 + i], inv[2 + i]);
            o.c1(is_eq);
            o.c1(is_limb_eq);
            is_eq = is_eq & is_limb_eq;
            o.c1(is_eq);
        }
        // or
        o.c1(is_zero);
        o.c1(is_eq);
        o.c1(is_zero | is_eq);
    }
}

// OP_MASK_INT (integer_chip.rs:511-520): mul(a_i, cond) rows. cond is boolean.
template <int FID>
H2E_HD void op_mask_int(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    Out o = out_at<Out>(ln, in.out);
    u32 cond[8];
    ld_slot8(ln, in.a[T::L + 1], cond);
    bool keep = cond[0] != 0;
    u32 av[T::L + 1][8];
    H2E_UNROLL
    for (int i = 0; i <= T::L; i++) ld_slot8(ln, in.a[i], av[i]);
    H2E_UNROLL
    for (int i = 0; i <= T::L; i++) {
        u32 z[8];
        const u32* a = av[i];
        H2E_UNROLL
        for (int k = 0; k < 8; k++) z[k] = keep ? a[k] : 0;
        o.c8(a);
        o.c8(cond);
        o.c8(z);
    }
}

// bisec row (base_chip.rs:574-598): [cond, a, cond, b] last(c)
template <class O>
H2E_HD void emit_bisec(O& o, const u32* cond, const u32* a, const u32* b) {
    bool pick_a = cond[0] != 0;
    u32 c[8];
    H2E_UNROLL
    for (int k = 0; k < 8; k++) c[k] = pick_a ? a[k] : b[k];
    o.c8(cond);
    o.c8(a);
    o.c8(cond);
    o.c8(b);
    o.c8(c);
}

// OP_BISEC_INT (integer_chip.rs:660-681)
template <int FID>
H2E_HD void op_bisec_int(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    Out o = out_at<Out>(ln, in.out);
    u32 cond[8];
    ld_slot8(ln, in.a[0], cond);
    u32 av[T::L + 1][8], bv[T::L + 1][8];
    H2E_UNROLL
    for (int i = 0; i <= T::L; i++) {
        ld_slot8(ln, in.a[1 + i], av[i]);
        ld_slot8(ln, in.a[2 + T::L + i], bv[i]);
    }
    H2E_UNROLL
    for (int i = 0; i <= T::L; i++) emit_bisec(o, cond, av[i], bv[i]);
}

// OP_SUM_ASSERT_ZERO (integer_chip.rs:607-611): sum of limbs row + assert_constant(sum, 0)
template <int FID>
H2E_HD void op_sum_assert_zero(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    Out o = out_at<Out>(ln, in.out);
    u32 sum[8];
    bn_zero<8>(sum);
    u32 al[T::L][4];
    load_int_limbs<T>(ln, in.a, al);
    H2E_UNROLL
    for (int i = 0; i < T::L; i++) {
        const u32* a = al[i];
        u32 t[8] = {a[0], a[1], a[2], a[3], 0, 0, 0, 0};
        bn_add<8>(sum, sum, t);
        o.c4(a);
    }
    o.c8(sum);
    o.c8(sum);
    if (!bn_is_zero<8>(sum)) ln.status |= ST_ASSERT_VALUE;
}

// ------------------------------- base chip macro-ops ----------------------------------------
H2E_HD void ld_const8(const LaneCtx& ln, u32 idx, u32* w) {
    H2E_UNROLL
    for (int k = 0; k < 8; k++) w[k] = ln.cpool[(size_t)idx * 8 + k];
}

H2E_HD void op_assign(LaneCtx& ln, const Instr& in) {
    u32 w[8];
    ld_input8(ln, in.a[0], w);
    st_cell8(blk_cell(ln, in.out, 0, 0, 8), w);
}
H2E_HD void op_assign_const(LaneCtx& ln, const Instr& in) {
    u32 w[8];
    if (in.a[0] == 0)
        ld_input8(ln, in.a[1], w);
    else
        ld_const8(ln, in.a[1], w);
    st_cell8(blk_cell(ln, in.out, 0, 0, 8), w);
}
// assign_bit (base_chip.rs:357-367): [a, a]
H2E_HD void op_assign_bit(LaneCtx& ln, const Instr& in) {
    u32 w[8];
    ld_input8(ln, in.a[0], w);
    Out o = out_at<Out>(ln, in.out);
    o.c8(w);
    o.c8(w);
}
// sum_with_constant_in_one_line (base_chip.rs:110-132): [x_i ...] last(sum)
static H2E_HDN void op_linsum(LaneCtx& ln, const Instr& in) {
    const FrConst& F = H2E_CONSTS.fr;
    Out o = out_at<Out>(ln, in.out);
    u32 n = in.a[0];
    u32 sum[8];
    if (in.a[1] != NONE)
        ld_const8(ln, in.a[1], sum);
    else
        bn_zero<8>(sum);
    for (u32 i = 0; i < n; i++) {
        u32 x[8], c[8], t[8];
        ld_slot8(ln, in.a[2 + 2 * i], x);
        ld_const8(ln, in.a[3 + 2 * i], c);
        o.c8(x);
        fr_mul(F, t, x, c);
        fr_add(F, sum, sum, t);
    }
    o.c8(sum);
}
// mul (base_chip.rs:176-193): [a, b] last(ab)
H2E_HD void op_mul(LaneCtx& ln, const Instr& in) {
    u32 a[8], b[8], c[8];
    ld_slot8(ln, in.a[0], a);
    ld_slot8(ln, in.a[1], b);
    fr_mul(H2E_CONSTS.fr, c, a, b);
    Out o = out_at<Out>(ln, in.out);
    o.c8(a);
    o.c8(b);
    o.c8(c);
}
// or / xor / xnor / not_and (base_chip.rs:405-467): [a, b] last(c), computed in Fr
H2E_HD void op_bool(LaneCtx& ln, const Instr& in) {
    const FrConst& F = H2E_CONSTS.fr;
    u32 a[8], b[8], ab[8], c[8], t[8];
    ld_slot8(ln, in.a[0], a);
    ld_slot8(ln, in.a[1], b);
    {
        // both operands are bits (the keccak chip: ~150k of these rows per permutation): no field product needed
        u32 hi = (a[0] | b[0]) >> 1;
        H2E_UNROLL
        for (int k = 1; k < 8; k++) hi |= a[k] | b[k];
        if (hi == 0) {
            const u32 x = a[0], y = b[0];
            const u32 kind = in.a[2];
            const u32 r = kind == 1 ? (x | y) : (kind == 2 ? (x ^ y) : (kind == 3 ? (1u ^ x ^ y) : (y & ~x)));
            u32 cz[8] = {r, 0, 0, 0, 0, 0, 0, 0};
            Out o = out_at<Out>(ln, in.out);
            o.c8(a);
            o.c8(b);
            o.c8(cz);
            return;
        }
    }
    fr_mul(F, ab, a, b);
    u32 one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    switch (in.a[2]) {
        case 1:  // or: a + b - ab
            fr_add(F, t, a, b);
            fr_sub(F, c, t, ab);
            break;
        case 2:  // xor: a + b - 2ab
            fr_add(F, t, a, b);
            fr_sub(F, t, t, ab);
            fr_sub(F, c, t, ab);
            break;
        case 3:  // xnor: 1 - a - b + 2ab
            fr_sub(F, t, one, a);
            fr_sub(F, t, t, b);
            fr_add(F, t, t, ab);
            fr_add(F, c, t, ab);
            break;
        default:  // not_and: b - ab
            fr_sub(F, c, b, ab);
            break;
    }
    Out o = out_at<Out>(ln, in.out);
    o.c8(a);
    o.c8(b);
    o.c8(c);
}
// one or / xor / xnor / not_and value (base_chip.rs:405-467): bits take the fast path, anything else is computed in Fr
H2E_HD void bool_value(const FrConst& F, u32 kind, const u32* a, const u32* b, u32* c) {
    u32 hi = (a[0] | b[0]) >> 1;
    H2E_UNROLL
    for (int k = 1; k < 8; k++) hi |= a[k] | b[k];
    if (hi == 0) {
        const u32 x = a[0], y = b[0];
        c[0] = kind == 1 ? (x | y) : (kind == 2 ? (x ^ y) : (kind == 3 ? (1u ^ x ^ y) : (y & ~x)));
        H2E_UNROLL
        for (int k = 1; k < 8; k++) c[k] = 0;
        return;
    }
    u32 ab[8], t[8], one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    fr_mul_call(F, ab, a, b);
    switch (kind) {
        case 1:
            fr_add(F, t, a, b);
            fr_sub(F, c, t, ab);
            break;
        case 2:
            fr_add(F, t, a, b);
            fr_sub(F, t, t, ab);
            fr_sub(F, c, t, ab);
            break;
        case 3:
            fr_sub(F, t, one, a);
            fr_sub(F, t, t, b);
            fr_add(F, t, t, ab);
            fr_add(F, c, t, ab);
            break;
        default: fr_sub(F, c, b, ab); break;
    }
}
// The vector forms work on AssignedCondition cells (bits by construction: assign_bit rows and the outputs of boolean rows),
// so their cells are stored at one word each; an operand or result that does not fit raises ST_RANGE (records unspecified).
H2E_HD u32 one_word(const u32* x, u32& status) {
    u32 hi = 0;
    H2E_UNROLL
    for (int k = 1; k < 8; k++) hi |= x[k];
    if (hi) status |= ST_RANGE;
    return x[0];
}
// value of a cell that must fit one word: a 1-word cell (the common case: the reference carries its class) is one 32-bit load
H2E_HD u32 ld_word(LaneCtx& ln, u32 ref) {
#if !defined(H2E_WIDTH_PROBE)
    if ((ref & 3u) == 0u) return ref_base(ln, ref)[ln.lane];
#endif
    u32 w[8];
    ld_slot8(ln, ref, w);
    return one_word(w, ln.status);
}
// or / xor / xnor / not_and of two one-word values: bits by bit operations, anything else through the field (the result then
// does not fit a word in general: ST_RANGE)
H2E_HD u32 bool_word(LaneCtx& ln, u32 kind, u32 x, u32 y) {
    if (((x | y) >> 1) == 0u) return kind == 1 ? (x | y) : (kind == 2 ? (x ^ y) : (kind == 3 ? (1u ^ x ^ y) : (y & ~x)));
    u32 a[8] = {x, 0, 0, 0, 0, 0, 0, 0}, b[8] = {y, 0, 0, 0, 0, 0, 0, 0}, c[8];
    bool_value(H2E_CONSTS.fr, kind, a, b, c);
    return one_word(c, ln.status);
}
// OP_BOOLV: n independent rows [a_i, b_i] last(c_i); the operands of eight elements are in flight at a time
static H2E_HDN void op_boolv(LaneCtx& ln, const Instr& in) {
    const u32 kind = in.a[0], n = in.a[1];
    const u32* tab = ln.tables + in.a[2];
    Out o = out_at<Out>(ln, in.out);
    u32 i = 0;
    for (; i + 8 <= n; i += 8) {
        u32 a[8], b[8];
        H2E_UNROLL
        for (int e = 0; e < 8; e++) {
            a[e] = ld_word(ln, tab[2 * (i + e)]);
            b[e] = ld_word(ln, tab[2 * (i + e) + 1]);
        }
        H2E_UNROLL
        for (int e = 0; e < 8; e++) {
            o.c1(a[e]);
            o.c1(b[e]);
            o.c1(bool_word(ln, kind, a[e], b[e]));
        }
    }
    for (; i < n; i++) {
        const u32 a = ld_word(ln, tab[2 * i]), b = ld_word(ln, tab[2 * i + 1]);
        o.c1(a);
        o.c1(b);
        o.c1(bool_word(ln, kind, a, b));
    }
}
// OP_CHIV: per element t = not_and(u, v) = v - u v: row [u, v] last(t); then xor(s, t): row [s, t] last(out)
static H2E_HDN void op_chiv(LaneCtx& ln, const Instr& in) {
    const u32 n = in.a[1];
    const u32* tab = ln.tables + in.a[2];
    Out o = out_at<Out>(ln, in.out);
    u32 i = 0;
    for (; i + 4 <= n; i += 4) {
        u32 u[4], v[4], s[4];
        H2E_UNROLL
        for (int e = 0; e < 4; e++) {
            u[e] = ld_word(ln, tab[3 * (i + e)]);
            v[e] = ld_word(ln, tab[3 * (i + e) + 1]);
            s[e] = ld_word(ln, tab[3 * (i + e) + 2]);
        }
        H2E_UNROLL
        for (int e = 0; e < 4; e++) {
            const u32 t = bool_word(ln, 4u, u[e], v[e]);
            o.c1(u[e]);
            o.c1(v[e]);
            o.c1(t);
            o.c1(s[e]);
            o.c1(t);
            o.c1(bool_word(ln, 2u, s[e], t));
        }
    }
    for (; i < n; i++) {
        const u32 u = ld_word(ln, tab[3 * i]), v = ld_word(ln, tab[3 * i + 1]), s = ld_word(ln, tab[3 * i + 2]);
        const u32 t = bool_word(ln, 4u, u, v);
        o.c1(u);
        o.c1(v);
        o.c1(t);
        o.c1(s);
        o.c1(t);
        o.c1(bool_word(ln, 2u, s, t));
    }
}
// bisec (base_chip.rs:574-598) in Fr: c = cond*a + (1-cond)*b
H2E_HD void op_bisec(LaneCtx& ln, const Instr& in) {
    const FrConst& F = H2E_CONSTS.fr;
    u32 cond[8], a[8], b[8];
    ld_slot8(ln, in.a[0], cond);
    ld_slot8(ln, in.a[1], a);
    ld_slot8(ln, in.a[2], b);
    Out o = out_at<Out>(ln, in.out);
    u32 hi = 0;
    H2E_UNROLL
    for (int k = 1; k < 8; k++) hi |= cond[k];
    if (hi == 0 && cond[0] <= 1) {
        emit_bisec(o, cond, a, b);
    } else {
        u32 t1[8], t2[8], one[8] = {1, 0, 0, 0, 0, 0, 0, 0}, c[8];
        fr_mul(F, t1, cond, a);
        fr_sub(F, t2, one, cond);
        fr_mul(F, t2, t2, b);
        fr_add(F, c, t1, t2);
        o.c8(cond);
        o.c8(a);
        o.c8(cond);
        o.c8(b);
        o.c8(c);
    }
}
H2E_HD void op_is_zero(LaneCtx& ln, const Instr& in) {
    u32 a[8];
    ld_slot8(ln, in.a[0], a);
    Out o = out_at<Out>(ln, in.out);
    emit_is_zero(H2E_CONSTS, o, a);
}
// assert_constant (base_chip.rs:375-379): value check + [a]
H2E_HD void op_assert_const(LaneCtx& ln, const Instr& in) {
    u32 a[8], c[8];
    ld_slot8(ln, in.a[0], a);
    ld_const8(ln, in.a[1], c);
    u32 d = 0;
    H2E_UNROLL
    for (int k = 0; k < 8; k++) d |= a[k] ^ c[k];
    if (d) ln.status |= in.a[2] ? in.a[2] : (u32)ST_ASSERT_VALUE;
    st_cell8(blk_cell(ln, in.out, 0, 0, 8), a);
}
// assert_equal (base_chip.rs:369-373): [a, b]
H2E_HD void op_assert_equal(LaneCtx& ln, const Instr& in) {
    u32 a[8], b[8];
    ld_slot8(ln, in.a[0], a);
    ld_slot8(ln, in.a[1], b);
    Out o = out_at<Out>(ln, in.out);
    o.c8(a);
    o.c8(b);
}

// ------------------------------- scalar decomposition / select chip --------------------------
// OP_DECOMPOSE_NATIVE (native_scalar_ecc_chip.rs:110-151): per 2 bits: assign_bit(b0) [b0,b0],
// assign_bit(b1) [b1,b1], row [v_next:4, b1:2, b0:1] last(v:-1); then assert_constant(v, 0).
static H2E_HDN void op_decompose_native(LaneCtx& ln, const Instr& in) {
    u32 s[8];
    ld_slot8(ln, in.a[0], s);
    Out o = out_at<Out>(ln, in.out);
    u32 v[8];
    bn_copy<8>(v, s);
    for (u32 i = 0; i < in.a[1]; i++) {
        u32 b0 = v[0] & 1u, b1 = (v[0] >> 1) & 1u;
        u32 vn[8];
        bn_shr<8, 8, 2>(vn, v);
        o.c1(b0);
        o.c1(b0);
        o.c1(b1);
        o.c1(b1);
        o.c8(vn);
        o.c1(b1);
        o.c1(b0);
        o.c8(v);
        bn_copy<8>(v, vn);
    }
    o.c8(v);
    if (!bn_is_zero<8>(v)) ln.status |= ST_ASSERT_VALUE;
}
// OP_DECOMPOSE_LIMB (general_scalar_ecc_chip.rs:108-128): per bit: assign_bit(b) [b,b], row
// [rest:-1, b:1] last(v:2) with v = (rest - b)/2; then assert_constant(rest, 0).
static H2E_HDN void op_decompose_limb(LaneCtx& ln, const Instr& in) {
    u32 rest[8];
    ld_slot8(ln, in.a[0], rest);
    Out o = out_at<Out>(ln, in.out);
    for (u32 j = 0; j < in.a[1]; j++) {
        u32 b = rest[0] & 1u;
        u32 v[8];
        bn_shr<8, 8, 1>(v, rest);
        o.c1(b);
        o.c1(b);
        o.c8(rest);
        o.c1(b);
        o.c8(v);
        bn_copy<8>(rest, v);
    }
    o.c8(rest);
    if (!bn_is_zero<8>(rest)) ln.status |= ST_ASSERT_VALUE;
}
// OP_CACHE_INT (ecc_chip.rs:734-751): L+1 cache rows, value column = copy of the source cell.
template <int FID>
H2E_HD void op_cache_int(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    Out o = out_at<Out>(ln, in.out);
    u32 av[T::L + 1][8];
    H2E_UNROLL
    for (int i = 0; i <= T::L; i++) ld_slot8(ln, in.a[i], av[i]);
    H2E_UNROLL
    for (int i = 0; i <= T::L; i++) o.c8(av[i]);
}
// OP_SELECT_INT (ecc_chip.rs:753-777 + 935-953): the candidate is chosen by byte 0 of the index
// cell; each select row holds [value copied from that candidate, selector = index].
template <int FID>
H2E_HD void op_select_int(LaneCtx& ln, const Instr& in) {
    typedef FT<FID> T;
    u32 idx[8];
    ld_slot8(ln, in.a[0], idx);
    u32 c = idx[0] & 0xffu;
    if (c >= in.a[2]) {
        ln.status |= ST_RANGE;
        c = 0;
    }
    const u32* tab = ln.tables + in.a[1] + (size_t)c * (T::L + 1);
    Out o = out_at<Out>(ln, in.out);
    u32 av[T::L + 1][8];
    H2E_UNROLL
    for (int i = 0; i <= T::L; i++) ld_slot8(ln, tab[i], av[i]);
    H2E_UNROLL
    for (int i = 0; i <= T::L; i++) {
        o.c8(av[i]);
        o.c8(idx);
    }
}

// ------------------------------- dispatch ----------------------------------------------------
template <int FID>
H2E_HD void exec_field_op(LaneCtx& ln, const Instr& in) {
    switch (in.op) {
        case OP_LOAD_INT: op_load_int<FID>(ln, in); break;
        case OP_ASSIGN_W: op_assign_w<FID>(ln, in); break;
        case OP_ASSIGN_INT_CONST: op_assign_int_const<FID>(ln, in); break;
        case OP_INT_ADD: op_int_linear<FID, 0>(ln, in); break;
        case OP_INT_SUB: op_int_linear<FID, 1>(ln, in); break;
        case OP_INT_NEG: op_int_linear<FID, 2>(ln, in); break;
        case OP_MUL_SMALL: op_int_linear<FID, 3>(ln, in); break;
        case OP_REDUCE: op_reduce<FID>(ln, in); break;
        case OP_INT_MUL: op_int_mul<FID>(ln, in); break;
        case OP_INT_MUL_HEAD: op_int_mul_head<FID>(ln, in); break;
        case OP_INT_MUL_TAIL: op_int_mul_tail<FID>(ln, in); break;
        case OP_REDUCE_HEAD: op_reduce_head<FID>(ln, in); break;
        case OP_REDUCE_TAIL: op_reduce_tail<FID>(ln, in); break;
        case OP_DIV_INV: op_div_inv<FID>(ln, in); break;
        case OP_DIV_CORE_S: op_div_core_s<FID>(ln, in); break;
        case OP_DIV_CORE: op_div_core<FID>(ln, in); break;
        case OP_IS_INT_ZERO: op_is_int_zero<FID>(ln, in); break;
        case OP_IS_INT_ZERO_HEAD: op_is_int_zero_head<FID>(ln, in); break;
        case OP_IS_INT_ZERO_TAIL: op_is_int_zero_tail<FID>(ln, in); break;
        case OP_DIV_HEAD_S: op_div_head_s<FID>(ln, in); break;
        case OP_DIV_TAIL: op_div_tail<FID>(ln, in); break;
        case OP_MASK_INT: op_mask_int<FID>(ln, in); break;
        case OP_BISEC_INT: op_bisec_int<FID>(ln, in); break;
        case OP_SUM_ASSERT_ZERO: op_sum_assert_zero<FID>(ln, in); break;
        case OP_CACHE_INT: op_cache_int<FID>(ln, in); break;
        case OP_SELECT_INT: op_select_int<FID>(ln, in); break;
        default: break;
    }
}

static H2E_HDN void exec_instr(LaneCtx& ln, const Instr& in) {
    switch (in.op) {
        case OP_NOP: break;
        case OP_ASSIGN: op_assign(ln, in); break;
        case OP_ASSIGN_CONST: op_assign_const(ln, in); break;
        case OP_ASSIGN_BIT: op_assign_bit(ln, in); break;
        case OP_LINSUM: op_linsum(ln, in); break;
        case OP_MUL: op_mul(ln, in); break;
        case OP_BOOL: op_bool(ln, in); break;
        case OP_BISEC: op_bisec(ln, in); break;
        case OP_IS_ZERO: op_is_zero(ln, in); break;
        case OP_ASSERT_CONST: op_assert_const(ln, in); break;
        case OP_ASSERT_EQUAL: op_assert_equal(ln, in); break;
        case OP_DECOMPOSE_NATIVE: op_decompose_native(ln, in); break;
        case OP_DECOMPOSE_LIMB: op_decompose_limb(ln, in); break;
        case OP_BOOLV: op_boolv(ln, in); break;
        case OP_CHIV: op_chiv(ln, in); break;
        default:
            switch (in.field) {
                case F_BN256_FQ: exec_field_op<F_BN256_FQ>(ln, in); break;
                case F_BLS12_381_FQ: exec_field_op<F_BLS12_381_FQ>(ln, in); break;
                case F_BLS12_381_FR: exec_field_op<F_BLS12_381_FR>(ln, in); break;
                default: break;
            }
    }
}

}  // namespace h2e
