// C-ABI implementation (include/h2ecc_b200.h): shapes, schedules, device state and launches. The kernels
// live in vm_kernel.cu.
#include <cuda_runtime.h>

#include <immintrin.h>

#include <atomic>
#include <functional>
#include <mutex>
#include <string>
#include <thread>

#include "../../include/h2ecc_b200.h"
#include "circuits.h"
#include "schedule.h"
#include "script_builder.h"
#include "fieldinfo.h"
#include "layout.h"
#include "vm_kernel.h"

using namespace h2e;
struct h2e_stream;
static void stream_destroy(h2e_stream* st);

#ifndef H2E_BLOCK
#define H2E_BLOCK 128
#endif

// -----------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches(0);

struct DeviceState {
    Instr* d_prog = nullptr;
    u32* d_cpool = nullptr;
    u32* d_tables = nullptr;
    struct Team {  // device copy of the TeamStreams built for one (CTAs per tile, warps per CTA) pair
        void* blob = nullptr;
        TeamProg prog = {};
    };
    std::map<int, Team> team;
    int sm_count = 0;
    bool consts_uploaded = false;
    // record export tables (layout.h): per format the words-per-lane prefix over the selected slots, and the
    // list of selected slots of the UNIQUE form
    u32* d_off_compact = nullptr;
    u32* d_off_unique = nullptr;   // UNIQUE: prefix (words per lane) over the stored slots
    u32* d_sel_unique = nullptr;   //         the stored slots
    u32* d_off_primary = nullptr;  // PRIMARY: same
    u32* d_sel_primary = nullptr;
    const u32* d_off_packed(int format) const { return format == REC_PRIMARY ? d_off_primary : d_off_unique; }
    const u32* d_sel_packed(int format) const { return format == REC_PRIMARY ? d_sel_primary : d_sel_unique; }
    u32* d_scatter_dst[2] = {nullptr, nullptr};  // slot -> cell index, column-major / row-major (h2e_records_scatter)
    u32* d_scatter_ord[2] = {nullptr, nullptr};  // the slots sorted by that cell index
    // pipelines of the host-buffer entry points (h2e_batch_run_host*), one per record format, kept across calls
    h2e_stream* host_pipe[REC_FORMATS] = {nullptr, nullptr, nullptr, nullptr};
};

struct h2e_shape {
    Context ctx;
    Schedule sched;
    bool sched_ready = false;
    int force_mode = 0;  // 0 auto, 1 thread-per-instance, 2 team
    int force_ctas = 0;
    int force_crit = 0;  // critical warps per CTA (0 = by estimated work)
    int force_warps = 0;  // 8 or 16 warps per CTA (0 = by shape and batch size)
    int export_format = 0;  // H2E_EXPORT_* applied by the host-buffer entry point
    Layout lay;              // static record layouts (width classes, copy classes), built on first use
    bool lay_ready = false;
    bool probe_checked = false;  // the width table has been compared with the device code's own widths (h2e_compact_prepare)
    // consumer-side expansion plans (h2e_records_expand, modes 1 / 2): the slots in the order of their destination cell
    struct ExpandItem {
        uint32_t dst;    // cell index inside an instance's dense array
        uint32_t slot;
    };
    std::vector<ExpandItem> expand_plan[2];
    std::mutex mu;
    std::map<int, DeviceState> dev;
};

#define CUDA_OK(call)                                                             \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e__);          \
            return -2;                                                            \
        }                                                                         \
    } while (0)

static int ensure_layout(h2e_shape* s);

// device copy of a slot-numbered program: slot numbers -> references into the COMPACT records (layout.h)
static std::vector<Instr> device_program(const h2e_shape* s, const std::vector<Instr>& prog) {
    std::vector<Instr> p(prog);
    translate_program(p.data(), p.size(), s->lay.off_compact.data(), s->lay.width.data());
    return p;
}

static int ensure_device(h2e_shape* s, int device, DeviceState** out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_err = "no CUDA device available: the witness VM has no CPU fallback";
        return -3;
    }
    int rc = ensure_layout(s);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(s->mu);
    CUDA_OK(cudaSetDevice(device));
    DeviceState& d = s->dev[device];
    if (!d.d_prog) {
        const Shape& sh = s->ctx.shape;
        size_t np = std::max<size_t>(sh.program.size(), 1), nc = std::max<size_t>(sh.consts.size(), 1);
        CUDA_OK(cudaMalloc(&d.d_prog, np * sizeof(Instr)));
        CUDA_OK(cudaMalloc(&d.d_cpool, nc * 32));
        CUDA_OK(cudaMalloc(&d.d_tables, std::max<size_t>(sh.tables.size(), 1) * 4));
        if (!sh.tables.empty()) {
            std::vector<uint32_t> tab(sh.tables.size());
            for (size_t i = 0; i < tab.size(); i++) tab[i] = sh.tables[i] < sh.slot_cell.size() ? slot_ref(sh.tables[i], s->lay.off_compact.data(), s->lay.width.data()) : 0;
            CUDA_OK(cudaMemcpy(d.d_tables, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
        }
        if (!sh.program.empty()) {
            std::vector<Instr> p = device_program(s, sh.program);
            CUDA_OK(cudaMemcpy(d.d_prog, p.data(), p.size() * sizeof(Instr), cudaMemcpyHostToDevice));
        }
        if (!sh.consts.empty()) CUDA_OK(cudaMemcpy(d.d_cpool, sh.consts.data(), sh.consts.size() * 32, cudaMemcpyHostToDevice));
        CUDA_OK(vm_upload_consts_w8(&host_consts()));
        CUDA_OK(vm_upload_consts_w16(&host_consts()));
        CUDA_OK(vm_upload_consts_wprobe(&host_consts()));
        CUDA_OK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device));
    }
    *out = &d;
    return 0;
}

static uint64_t pad_tiles(uint64_t n) { return (n + TILE - 1) / TILE * TILE; }

// critical / tail warp split of a CTA, in proportion to the estimated work of the two instruction
// classes, biased towards the critical warps (the critical path, not tail throughput, bounds the run)
static double crit_fraction(h2e_shape* s) {
    double wc = 0, wt = 0;
    for (const Instr& in : s->sched.program) ((in.flags & 0x80) ? wt : wc) += instr_cost(in);
    return wc / std::max(wc + wt, 1.0);
}
static int pick_n_crit(h2e_shape* s, int warps) {
    int n_crit = (int)(warps * crit_fraction(s) + 0.5) + warps / 8;
    n_crit = std::min(std::max(n_crit, 1), warps - 1);
    if (s->force_crit > 0) n_crit = std::min(s->force_crit, warps - 1);
    return n_crit;
}

// Kernel variant: 8 warps per CTA at 255 registers, or 16 warps at 128. The inversion-heavy shapes
// (MSM: one W and one Fr inversion per point addition) need the registers; the tower arithmetic of
// the pairing shapes gains more from twice the resident warps once many tiles share the GPU
// (bn256 pairing, 896 instances: 60 -> 49.5 ms; MSM n=1000 x 128: 167 -> 206 ms with 16 warps).
static int pick_warps(h2e_shape* s, uint64_t tiles) {
    if (s->force_warps == 8 || s->force_warps == 16) return s->force_warps;
    double inv = 0, all = 0;
    for (const Instr& in : s->sched.program) {
        if (in.flags & 0x80) continue;
        double c = instr_cost(in);
        all += c;
        if (in.op == OP_IS_INT_ZERO || in.op == OP_DIV_CORE || in.op == OP_DIV_INV || in.op == OP_IS_ZERO) inv += c;
    }
    // measured on bn256 pairing (8 / 16 warps per CTA, split layout): 1 tile 21.1 / 22.7 ms, 16 tiles 37.1 / 32.8, 28 tiles 58.7 / 53.5;
    // 8 tiles 24.9 / 25.5; bls12_381 8 tiles 31.1 / 33.0, 16 tiles 51.6 / 50.4; MSM n=1000, 4 tiles 162 / 235
    return (tiles >= 16 && inv < 0.3 * all) ? 16 : 8;
}

// Build (once per shape) the levelised schedule and (once per device and CTAs-per-tile) upload the
// per-warp instruction streams.
static int ensure_team(h2e_shape* s, DeviceState* d, unsigned G, uint64_t tiles, TeamProg* out, int* warps_out) {
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->sched_ready) {
        s->sched = levelise(s->ctx.shape);
        s->sched_ready = true;
    }
    const int warps = pick_warps(s, tiles);
    *warps_out = warps;
    DeviceState::Team& t = d->team[(int)G * 32 + warps];
    if (!t.blob) {
        int n_crit = pick_n_crit(s, warps);
        // split layout when a tile has at least two CTAs: whole CTAs are critical or tail, in the same proportion
        TeamLayout lay = {G, (uint32_t)n_crit, G, (uint32_t)(warps - n_crit), 0};
        uint32_t g_crit = 0;
        if (G >= 2 && !getenv("H2E_MIXED")) {
            // measured (bn256 pairing, 16 tiles, G = 9): 5 critical + 4 tail CTAs 33.3 ms, 6 + 3 36.2, 4 + 5 41.1, 7 + 2 45.2
            double f = s->force_crit > 0 ? (double)n_crit / warps : crit_fraction(s);
            // at least one CTA in eight stays a tail CTA (MSM n=1000, G = 37: 32 + 5 CTAs 162 ms, 36 + 1 CTAs 245 ms)
            g_crit = (uint32_t)std::min<int>(std::max<int>((int)(G * f + 0.5), 1), (int)G - (int)((G + 7) / 8));
            lay = TeamLayout{g_crit, (uint32_t)warps, G - g_crit, (uint32_t)warps, g_crit};
            dedicate_inversion_ctas(s->sched, lay);
        }
        TeamStreams ts;
        try {
            const char *hl = getenv("H2E_HOP_LOCAL"), *hg = getenv("H2E_HOP_GLOBAL");  // tuning of the scheduler's latency model
            ts = build_team_streams(s->sched, lay, hl ? atof(hl) : 2000.0, hg ? atof(hg) : 2000.0);
        } catch (std::exception& e) {
            g_err = e.what();
            return -1;
        }
        translate_program(ts.crit.data(), ts.crit.size(), s->lay.off_compact.data(), s->lay.width.data());
        translate_program(ts.tail.data(), ts.tail.size(), s->lay.off_compact.data(), s->lay.width.data());
        auto pad = [](size_t x) { return (x + 255) / 256 * 256; };
        size_t sizes[7] = {std::max<size_t>(ts.crit.size(), 1) * sizeof(Instr), std::max<size_t>(ts.crit_dep.size(), 1) * sizeof(DepRec),
                           ts.crit_off.size() * 4, std::max<size_t>(ts.tail.size(), 1) * sizeof(Instr),
                           std::max<size_t>(ts.tail_dep.size(), 1) * sizeof(DepRec), ts.tail_off.size() * 4, ts.extra.size() * 4};
        const void* src[7] = {ts.crit.data(), ts.crit_dep.data(), ts.crit_off.data(), ts.tail.data(), ts.tail_dep.data(), ts.tail_off.data(),
                              ts.extra.data()};
        size_t srcsz[7] = {ts.crit.size() * sizeof(Instr), ts.crit_dep.size() * sizeof(DepRec), ts.crit_off.size() * 4, ts.tail.size() * sizeof(Instr),
                           ts.tail_dep.size() * sizeof(DepRec), ts.tail_off.size() * 4, ts.extra.size() * 4};
        size_t offs[8] = {0};
        for (int i = 0; i < 7; i++) offs[i + 1] = offs[i] + pad(sizes[i]);
        std::vector<uint8_t> host(offs[7], 0);
        for (int i = 0; i < 7; i++)
            if (srcsz[i]) memcpy(&host[offs[i]], src[i], srcsz[i]);
        CUDA_OK(cudaMalloc(&t.blob, offs[7]));
        CUDA_OK(cudaMemcpy(t.blob, host.data(), offs[7], cudaMemcpyHostToDevice));
        char* base = (char*)t.blob;
        t.prog.crit = (const Instr*)(base + offs[0]);
        t.prog.crit_dep = (const DepRec*)(base + offs[1]);
        t.prog.crit_off = (const uint32_t*)(base + offs[2]);
        t.prog.tail = (const Instr*)(base + offs[3]);
        t.prog.tail_dep = (const DepRec*)(base + offs[4]);
        t.prog.tail_off = (const uint32_t*)(base + offs[5]);
        t.prog.extra = (const uint32_t*)(base + offs[6]);
        t.prog.n_levels = 0;
        t.prog.n_crit = (uint32_t)n_crit;
        t.prog.G = G;
        t.prog.twc = ts.twc;
        t.prog.g_crit = g_crit;
    }
    *out = t.prog;
    return 0;
}

// Launch one pass of the VM over `tiles` tiles. Chooses thread-per-instance (many instances, short
// program) or team mode (few instances, long program).
static int launch_vm_group(h2e_shape* s, DeviceState* d, cudaStream_t stream, u32* d_rec, const u32* d_inputs, u32* d_status, uint64_t n_inst);
static uint64_t compact_tile_words(const h2e_shape* s) { return (uint64_t)s->lay.off_compact.back() * TILE; }

// A program is "long" when one thread per instance would leave the GPU nearly empty for the instance counts that
// fit in HBM (a pairing check is 175k macro-ops and 197 MB of cells per instance): such shapes always run in team
// mode, in groups of at most SMs / 4 tiles per launch.
static bool long_program(const Shape& sh) { return sh.program.size() >= 4096; }
// Tiles per cooperative launch of a large batch: SMs / 4, i.e. 4 CTAs per tile (2 critical + 2 tail). Measured on the pairing
// shapes (ms per tile and 148 SMs): 4 CTAs per tile 1.13, 2 CTAs 1.43 (a quarter of the SMs idle), 3 CTAs 1.31, 5 CTAs 1.51.
static uint64_t team_group_tiles(const DeviceState* d) { return (uint64_t)std::max(1, (d->sm_count > 0 ? d->sm_count : 148) / 4); }

// Launch the VM over n_inst instances: d_rec receives the COMPACT records (whole tiles).
static int launch_vm(h2e_shape* s, DeviceState* d, cudaStream_t stream, u32* d_rec, const u32* d_inputs, u32* d_status, uint64_t n_inst) {
    const Shape& sh = s->ctx.shape;
    const uint64_t tiles = pad_tiles(n_inst) / TILE, group = team_group_tiles(d);
    // Also grouped: shorter programs of large macro-ops (the keccak hash: 2 k vector macro-ops, 470 k cells per instance) at batch
    // sizes that would give one thread per instance only a few warps per SM. Measured on that shape: team groups 240 k hashes/s at
    // any batch size; one thread per instance 60 k/s at 8192 instances, 893 k/s at 65536 (crossover near 18 k instances = 4 tiles
    // per SM).
    const int sms = d->sm_count > 0 ? d->sm_count : 148;
    const bool grouped = long_program(sh) || (sh.slot_cell.size() >= 100000 && sh.program.size() >= 64 && tiles < (uint64_t)4 * sms);
    if (s->force_mode != 0 || !grouped || tiles <= group) return launch_vm_group(s, d, stream, d_rec, d_inputs, d_status, n_inst);
    // more tiles than one cooperative launch can hold: equal groups, back to back on the stream
    const uint64_t n_groups = (tiles + group - 1) / group, per = (tiles + n_groups - 1) / n_groups;
    for (uint64_t t0 = 0; t0 < tiles; t0 += per) {
        const uint64_t nt = std::min(per, tiles - t0), i0 = t0 * TILE, ni = std::min<uint64_t>(n_inst - i0, nt * TILE);
        int rc = launch_vm_group(s, d, stream, d_rec + t0 * compact_tile_words(s), d_inputs + i0 * sh.n_inputs * 8, d_status + i0, ni);
        if (rc) return rc;
    }
    return 0;
}

static int launch_vm_group(h2e_shape* s, DeviceState* d, cudaStream_t stream, u32* d_vals, const u32* d_inputs, u32* d_status, uint64_t n_inst) {
    const Shape& sh = s->ctx.shape;
    const uint64_t tile_words = compact_tile_words(s);
    uint64_t padded = pad_tiles(n_inst), tiles = padded / TILE;
    int sms = d->sm_count > 0 ? d->sm_count : 148;
    bool team = sh.program.size() >= 64 && tiles * 2 <= (uint64_t)sms;
    if (s->force_mode == 1) team = false;
    if (s->force_mode >= 2) team = true;
    if (team && tiles > (uint64_t)sms) {
        g_err = "team mode needs every CTA resident: at most one tile per SM";
        return -1;
    }
    if (!team) {
        const int block = H2E_BLOCK;
        uint64_t grid = (padded + block - 1) / block;
        TeamProg flat = {};
        flat.crit = d->d_prog;
        flat.n_levels = (uint32_t)sh.program.size();
        VmLaunch L = {(unsigned)grid, (unsigned)block, stream, flat, d_vals, d_inputs, d->d_cpool, d->d_tables, d_status, nullptr, nullptr, 0,
                      tile_words, (uint32_t)sh.n_inputs, n_inst, tiles, 0};
        g_launches++;
        // Variant: since the macro-ops write the compact layout, thread mode is no longer bound by the store stream but by
        // latency at low occupancy; twice the resident warps (128 registers, a few spills) measured 0.552 vs 0.610 ms per step
        // of configs[1]. H2E_THREAD_W8=1 selects the 255-register build.
        CUDA_OK(getenv("H2E_THREAD_W8") ? vm_launch_w8(L) : vm_launch_w16(L));
        return 0;
    }
    // CTAs per tile: all CTAs of the grid must be resident at once (one CTA per SM at 255 registers x 256 threads)
    unsigned G = (unsigned)std::max<uint64_t>(1, (uint64_t)sms / tiles);
    if (s->force_ctas > 0) G = (unsigned)std::min<uint64_t>((uint64_t)s->force_ctas, std::max<uint64_t>(1, (uint64_t)sms / tiles));
    TeamProg prog;
    int warps = 8;
    int rc = ensure_team(s, d, G, tiles, &prog, &warps);
    if (rc) return rc;
    CUDA_OK(cudaMemsetAsync(d_status, 0, padded * 4, stream));
    // progress counters of this launch (stream-ordered allocation: concurrent launches never share them)
    u32* d_progress = nullptr;
    // + the scratch entries (64 bytes per instance each: inverses handed from OP_DIV_INV to OP_DIV_CORE_S)
    const size_t pbytes = ((size_t)tiles * prog.twc * 4 + 255) / 256 * 256;
    const size_t sbytes = (size_t)tiles * s->sched.n_scratch * TILE * 64;
    CUDA_OK(cudaMallocAsync((void**)&d_progress, pbytes + sbytes, stream));
    CUDA_OK(cudaMemsetAsync(d_progress, 0, pbytes, stream));
    u32* d_scratch = (u32*)((char*)d_progress + pbytes);
    VmLaunch L = {(unsigned)(tiles * G), (unsigned)warps * 32, stream, prog, d_vals, d_inputs, d->d_cpool, d->d_tables, d_status, d_progress, d_scratch,
                  s->sched.n_scratch, tile_words, (uint32_t)sh.n_inputs, n_inst, tiles, (s->force_mode >= 3 ? s->force_mode : 1) | (getenv("H2E_ACQ_POLL") ? 0x100 : 0)};
    g_launches++;
    CUDA_OK(warps == 16 ? vm_launch_w16(L) : vm_launch_w8(L));
    CUDA_OK(cudaFreeAsync(d_progress, stream));
    return 0;
}

// Static record layouts of the shape (host only).
static int ensure_layout(h2e_shape* s) {
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->lay_ready) return 0;
    try {
        s->lay = build_layout(s->ctx.shape);
    } catch (std::exception& e) {
        g_err = e.what();
        return -1;
    }
    s->lay_ready = true;
    return 0;
}

// Device copies of the export tables.
static int ensure_layout_device(h2e_shape* s, DeviceState* d) {
    int rc = ensure_layout(s);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(s->mu);
    if (d->d_off_compact) return 0;
    const Layout& lay = s->lay;
    for (int format : {(int)REC_UNIQUE, (int)REC_PRIMARY}) {
        const std::vector<uint32_t>& sel = lay.stored_slots(format);
        std::vector<uint32_t> uoff(sel.size() + 1, 0);  // prefix over the stored slots only
        for (size_t i = 0; i < sel.size(); i++) uoff[i + 1] = uoff[i] + lay.width[sel[i]];
        u32 *&d_off = format == REC_PRIMARY ? d->d_off_primary : d->d_off_unique, *&d_sel = format == REC_PRIMARY ? d->d_sel_primary : d->d_sel_unique;
        CUDA_OK(cudaMalloc(&d_off, uoff.size() * 4));
        CUDA_OK(cudaMemcpy(d_off, uoff.data(), uoff.size() * 4, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&d_sel, std::max<size_t>(sel.size(), 1) * 4));
        if (!sel.empty()) CUDA_OK(cudaMemcpy(d_sel, sel.data(), sel.size() * 4, cudaMemcpyHostToDevice));
    }
    CUDA_OK(cudaMalloc(&d->d_off_compact, lay.off_compact.size() * 4));
    CUDA_OK(cudaMemcpy(d->d_off_compact, lay.off_compact.data(), lay.off_compact.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}

// Cross-check of the width table (layout.h restates the width class of every store of the macro-op code): the
// width-probe build of the VM runs the program once, in thread mode, with every store writing the width class of
// its cell instead of its value; the two must agree for every slot.
static int check_widths_on_device(h2e_shape* s, DeviceState* d) {
    const Shape& sh = s->ctx.shape;
    const size_t n_slots = sh.slot_cell.size();
    if (s->probe_checked) return 0;
    u32 *d_cells = nullptr, *d_in = nullptr, *d_status = nullptr;
    CUDA_OK(cudaMalloc(&d_cells, std::max<size_t>(n_slots, 1) * 32));
    CUDA_OK(cudaMemset(d_cells, 0, std::max<size_t>(n_slots, 1) * 32));
    CUDA_OK(cudaMalloc(&d_in, std::max<size_t>(sh.n_inputs, 1) * 32));
    CUDA_OK(cudaMemset(d_in, 0, std::max<size_t>(sh.n_inputs, 1) * 32));
    CUDA_OK(cudaMalloc(&d_status, TILE * 4));
    // (the probe build addresses cells by slot number: it runs the program and the slot tables as traced)
    Instr* d_raw = nullptr;
    u32* d_rawtab = nullptr;
    CUDA_OK(cudaMalloc(&d_raw, std::max<size_t>(sh.program.size(), 1) * sizeof(Instr)));
    CUDA_OK(cudaMalloc(&d_rawtab, std::max<size_t>(sh.tables.size(), 1) * 4));
    if (!sh.program.empty()) CUDA_OK(cudaMemcpy(d_raw, sh.program.data(), sh.program.size() * sizeof(Instr), cudaMemcpyHostToDevice));
    if (!sh.tables.empty()) CUDA_OK(cudaMemcpy(d_rawtab, sh.tables.data(), sh.tables.size() * 4, cudaMemcpyHostToDevice));
    TeamProg flat = {};
    flat.crit = d_raw;
    flat.n_levels = (uint32_t)sh.program.size();
    VmLaunch L = {1u, (unsigned)TILE, 0, flat, d_cells, d_in, d->d_cpool, d_rawtab, d_status, nullptr, nullptr, 0, 0, (uint32_t)sh.n_inputs,
                  1, 1, 0};
    g_launches++;
    CUDA_OK(vm_launch_wprobe(L));
    std::vector<uint32_t> cells(n_slots * 8);
    CUDA_OK(cudaMemcpy(cells.data(), d_cells, n_slots * 32, cudaMemcpyDeviceToHost));
    cudaFree(d_cells);
    cudaFree(d_in);
    cudaFree(d_status);
    cudaFree(d_raw);
    cudaFree(d_rawtab);
    for (size_t i = 0; i < n_slots; i++)
        if (cells[8 * i] != s->lay.width[i]) {
            g_err = "width table mismatch at slot " + std::to_string(i) + ": layout.h says " + std::to_string(s->lay.width[i]) +
                    " words, the device code stores " + std::to_string(cells[8 * i]);
            return -1;
        }
    s->probe_checked = true;
    return 0;
}

static int launch_montgomery(DeviceState* d, cudaStream_t stream, u32* d_cells, uint64_t n_cells) {
    if (n_cells == 0) return 0;
    int sms = d->sm_count > 0 ? d->sm_count : 148;
    uint64_t blocks = std::min<uint64_t>((n_cells + 255) / 256, (uint64_t)sms * 8);
    g_launches++;
    CUDA_OK(vm_montgomery_w8(stream, (unsigned)blocks, d_cells, n_cells));
    return 0;
}

extern "C" {

const char* h2e_last_error(void) { return g_err.c_str(); }
int h2e_version(void) { return 1; }
uint64_t h2e_launch_count(void) { return g_launches.load(); }

h2e_shape* h2e_shape_from_script(int field, const uint32_t* script, size_t n_words, const uint8_t* statics64, size_t n_statics) {
    try {
        if (field < 0 || field >= F_COUNT) throw std::runtime_error("bad field id");
        std::vector<Big> st;
        for (size_t i = 0; i < n_statics; i++) {
            uint32_t w[16];
            memcpy(w, statics64 + 64 * i, 64);
            st.push_back(Big::from_words(w, 16));
        }
        h2e_shape* s = new h2e_shape();
        try {
            run_script(s->ctx, (Field)field, script, n_words, st);
        } catch (...) {
            delete s;
            throw;
        }
        return s;
    } catch (std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

h2e_shape* h2e_shape_build(int circuit_kind, const uint64_t* params, size_t n_params) {
    try {
        h2e_shape* s = new h2e_shape();
        try {
            build_circuit(s->ctx, circuit_kind, params, n_params);
        } catch (...) {
            delete s;
            throw;
        }
        return s;
    } catch (std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void h2e_shape_free(h2e_shape* s) {
    if (!s) return;
    for (auto& kv : s->dev) {
        if (cudaSetDevice(kv.first) == cudaSuccess) {
            cudaFree(kv.second.d_prog);
            cudaFree(kv.second.d_cpool);
            cudaFree(kv.second.d_tables);
            for (auto& t : kv.second.team) cudaFree(t.second.blob);
            for (int k = 0; k < REC_FORMATS; k++)
                if (kv.second.host_pipe[k]) stream_destroy(kv.second.host_pipe[k]);
            cudaFree(kv.second.d_off_compact);
            cudaFree(kv.second.d_off_unique);
            cudaFree(kv.second.d_sel_unique);
            cudaFree(kv.second.d_off_primary);
            cudaFree(kv.second.d_sel_primary);
            cudaFree(kv.second.d_scatter_dst[0]);
            cudaFree(kv.second.d_scatter_dst[1]);
            cudaFree(kv.second.d_scatter_ord[0]);
            cudaFree(kv.second.d_scatter_ord[1]);
        }
    }
    delete s;
}

int h2e_shape_query(const h2e_shape* s, uint64_t out[16]) {
    const Shape& sh = s->ctx.shape;
    for (int i = 0; i < 3; i++) {
        out[i] = sh.height[i];
        out[3 + i] = sh.offset[i];
    }
    out[6] = sh.slot_cell.size();
    out[7] = sh.fixed.size();
    out[8] = sh.perms.size();
    out[9] = sh.program.size();
    out[10] = sh.consts.size();
    out[11] = sh.n_inputs;
    out[12] = sh.tables.size();
    return 0;
}
int h2e_shape_slot_cells(const h2e_shape* s, uint32_t* out) {
    const Shape& sh = s->ctx.shape;
    for (size_t i = 0; i < sh.slot_cell.size(); i++) {
        out[3 * i] = sh.slot_cell[i].region;
        out[3 * i + 1] = sh.slot_cell[i].col;
        out[3 * i + 2] = sh.slot_cell[i].row;
    }
    return 0;
}
int h2e_shape_fixed(const h2e_shape* s, uint32_t* out) {
    const Shape& sh = s->ctx.shape;
    for (size_t i = 0; i < sh.fixed.size(); i++) {
        out[4 * i] = sh.fixed[i].region;
        out[4 * i + 1] = sh.fixed[i].col;
        out[4 * i + 2] = sh.fixed[i].row;
        out[4 * i + 3] = sh.fixed[i].cidx;
    }
    return 0;
}
int h2e_shape_consts(const h2e_shape* s, uint8_t* out) {
    const Shape& sh = s->ctx.shape;
    if (!sh.consts.empty()) memcpy(out, sh.consts.data(), sh.consts.size() * 32);
    return 0;
}
int h2e_shape_program(const h2e_shape* s, uint8_t* out) {
    const Shape& sh = s->ctx.shape;
    if (!sh.program.empty()) memcpy(out, sh.program.data(), sh.program.size() * sizeof(Instr));
    return 0;
}
int h2e_shape_schedule(h2e_shape* s, uint64_t* n_levels, uint64_t* n_instr, uint8_t* program_out, uint32_t* level_start_out) {
    try {
        std::lock_guard<std::mutex> lk(s->mu);
        if (!s->sched_ready) {
            s->sched = levelise(s->ctx.shape);
            s->sched_ready = true;
        }
    } catch (std::exception& e) {
        g_err = e.what();
        return -1;
    }
    const Schedule& sc = s->sched;
    if (n_levels) *n_levels = sc.level_start.size() - 1;
    if (n_instr) *n_instr = sc.program.size();
    if (program_out && !sc.program.empty()) memcpy(program_out, sc.program.data(), sc.program.size() * sizeof(Instr));
    if (level_start_out) memcpy(level_start_out, sc.level_start.data(), sc.level_start.size() * 4);
    return 0;
}
int h2e_shape_team_order(h2e_shape* s, int ctas_per_tile, uint64_t* n_instr, uint8_t* program_out, double* est_cycles) {
    try {
        std::lock_guard<std::mutex> lk(s->mu);
        if (!s->sched_ready) {
            s->sched = levelise(s->ctx.shape);
            s->sched_ready = true;
        }
        if (n_instr) *n_instr = s->sched.program.size();
        if (!program_out && !est_cycles) return 0;
        if (ctas_per_tile < 1) throw std::runtime_error("ctas_per_tile must be >= 1");
        int n_crit = pick_n_crit(s, 8);
        uint32_t G = (uint32_t)ctas_per_tile;
        TeamLayout lay = {G, (uint32_t)n_crit, G, (uint32_t)(8 - n_crit), 0};
        if (G >= 2) {
            uint32_t g_crit = (uint32_t)std::min<int>(std::max<int>((int)(G * crit_fraction(s) + 0.5), 1), (int)G - (int)((G + 7) / 8));
            lay = TeamLayout{g_crit, 8, G - g_crit, 8, g_crit};
            dedicate_inversion_ctas(s->sched, lay);
        }
        TeamStreams ts = build_team_streams(s->sched, lay);
        if (est_cycles) *est_cycles = ts.est_cycles;
        if (program_out) {
            std::vector<Instr> order = simulate_team_order(ts);
            memcpy(program_out, order.data(), order.size() * sizeof(Instr));
        }
    } catch (std::exception& e) {
        g_err = e.what();
        return -1;
    }
    return 0;
}
int h2e_shape_tables(const h2e_shape* s, uint32_t* out) {
    const Shape& sh = s->ctx.shape;
    if (!sh.tables.empty()) memcpy(out, sh.tables.data(), sh.tables.size() * 4);
    return 0;
}
int h2e_shape_perms(const h2e_shape* s, uint32_t* out) {
    const Shape& sh = s->ctx.shape;
    for (size_t i = 0; i < sh.perms.size(); i++)
        for (int k = 0; k < 2; k++) {
            out[6 * i + 3 * k] = sh.perms[i][k].region;
            out[6 * i + 3 * k + 1] = sh.perms[i][k].col;
            out[6 * i + 3 * k + 2] = sh.perms[i][k].row;
        }
    return 0;
}

size_t h2e_vals_bytes(const h2e_shape* s, uint64_t n_inst) { return (size_t)pad_tiles(n_inst) * s->ctx.shape.slot_cell.size() * 32; }
size_t h2e_inputs_bytes(const h2e_shape* s, uint64_t n_inst) { return (size_t)n_inst * s->ctx.shape.n_inputs * 32; }

// COMPACT records -> WIDE cells on the device (all slots of `tiles` tiles)
static int launch_expand(h2e_shape* s, DeviceState* d, cudaStream_t stream, const u32* d_rec, u32* d_vals, uint64_t tiles) {
    const uint64_t n_slots = s->ctx.shape.slot_cell.size();
    if (!n_slots || !tiles) return 0;
    const int sms = d->sm_count > 0 ? d->sm_count : 148;
    g_launches++;
    CUDA_OK(vm_expand(stream, (unsigned)sms * 8, d_rec, d_vals, d->d_off_compact, compact_tile_words(s), 0, n_slots, tiles, n_slots * TILE * 8));
    return 0;
}
// COMPACT records -> UNIQUE / PRIMARY records on the device (all stored slots of `tiles` tiles)
static int launch_pack(h2e_shape* s, DeviceState* d, cudaStream_t stream, int format, const u32* d_rec, u32* d_out, uint64_t tiles) {
    const uint32_t n_sel = (uint32_t)s->lay.stored_slots(format).size();
    if (!n_sel || !tiles) return 0;
    const int sms = d->sm_count > 0 ? d->sm_count : 148;
    g_launches++;
    CUDA_OK(vm_pack(stream, (unsigned)sms * 8, d_rec, d_out, d->d_sel_packed(format), d->d_off_packed(format), d->d_off_compact, compact_tile_words(s), 0, n_sel,
                    tiles, (uint64_t)s->lay.off(format).back() * TILE));
    return 0;
}
static bool bad_format(int format) { return format < REC_WIDE || format >= REC_FORMATS; }

int h2e_batch_run_records(h2e_shape* s, int device, void* stream, int format, uint64_t n_inst, const void* d_inputs, void* d_records, uint32_t* d_status) {
    if (bad_format(format)) {
        g_err = "unknown record format";
        return -1;
    }
    if (n_inst == 0) return 0;
    DeviceState* d;
    int rc = ensure_device(s, device, &d);
    if (rc) return rc;
    rc = ensure_layout_device(s, d);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (format == REC_COMPACT) return launch_vm(s, d, st, (u32*)d_records, (const u32*)d_inputs, d_status, n_inst);
    // the VM's own layout is COMPACT: the other forms are derived from a stream-ordered temporary
    const uint64_t tiles = pad_tiles(n_inst) / TILE;
    u32* d_tmp = nullptr;
    CUDA_OK(cudaMallocAsync((void**)&d_tmp, tiles * compact_tile_words(s) * 4, st));
    rc = launch_vm(s, d, st, d_tmp, (const u32*)d_inputs, d_status, n_inst);
    if (!rc) rc = format == REC_WIDE ? launch_expand(s, d, st, d_tmp, (u32*)d_records, tiles) : launch_pack(s, d, st, format, d_tmp, (u32*)d_records, tiles);
    CUDA_OK(cudaFreeAsync(d_tmp, st));
    return rc;
}
int h2e_batch_run(h2e_shape* s, int device, void* stream, uint64_t n_inst, const void* d_inputs, void* d_vals, uint32_t* d_status) {
    return h2e_batch_run_records(s, device, stream, REC_WIDE, n_inst, d_inputs, d_vals, d_status);
}

// ---- record layouts ----------------------------------------------------------------------------
int h2e_shape_layout(h2e_shape* s, int format, uint32_t* off_out, uint8_t* width_out, uint32_t* root_out) {
    if (bad_format(format)) {
        g_err = "unknown record format";
        return -1;
    }
    int rc = ensure_layout(s);
    if (rc) return rc;
    const Layout& lay = s->lay;
    const size_t n = s->ctx.shape.slot_cell.size();
    if (off_out) {
        if (format == REC_WIDE)
            for (size_t i = 0; i <= n; i++) off_out[i] = (uint32_t)(8 * i);
        else
            memcpy(off_out, lay.off(format).data(), (n + 1) * 4);
    }
    if (width_out) {
        for (size_t i = 0; i < n; i++) width_out[i] = format == REC_WIDE ? 8 : lay.width[i];
    }
    if (root_out && n) memcpy(root_out, lay.root.data(), n * 4);
    return 0;
}
int h2e_shape_layout_derived(h2e_shape* s, uint32_t* src_out, uint8_t* shift_out) {
    int rc = ensure_layout(s);
    if (rc) return rc;
    const size_t n = s->ctx.shape.slot_cell.size();
    if (src_out && n) memcpy(src_out, s->lay.der_src.data(), n * 4);
    if (shift_out && n) memcpy(shift_out, s->lay.der_shift.data(), n);
    return 0;
}
size_t h2e_records_bytes(h2e_shape* s, int format, uint64_t n_inst) {
    if (bad_format(format) || ensure_layout(s)) return 0;
    return (size_t)(pad_tiles(n_inst) / TILE) * s->lay.words_per_lane(format, s->ctx.shape.slot_cell.size()) * TILE * 4;
}

// ---- compact export (round-1 names, kept) ------------------------------------------------------
int h2e_compact_prepare(h2e_shape* s, int device) {
    DeviceState* d;
    int rc = ensure_device(s, device, &d);
    if (rc) return rc;
    rc = ensure_layout_device(s, d);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(s->mu);
    return check_widths_on_device(s, d);
}
size_t h2e_compact_bytes(const h2e_shape* s, uint64_t n_inst) { return h2e_records_bytes(const_cast<h2e_shape*>(s), REC_COMPACT, n_inst); }
int h2e_compact_widths(const h2e_shape* s, uint8_t* out) { return h2e_shape_layout(const_cast<h2e_shape*>(s), REC_COMPACT, nullptr, out, nullptr); }

}  // extern "C"

// ---- streaming pipeline ------------------------------------------------------------------------
// One chunk = up to `chunk_tiles` whole tiles: inputs go up, the VM fills the chunk's value tiles, the export
// kernel packs them into the requested record format, and the records stream into the caller's (pinned) host
// buffer -- all on one of two CUDA streams, so that chunk k+1 computes while chunk k is on the bus. The caller
// owns the host buffers (a ring of them, reused as tickets complete): the library holds no host memory and no
// caller pointer past the completion of the ticket. Memory is bounded per chunk, not per batch
// (the reference bounds it per Context: context.rs:254-292).
struct h2e_stream {
    h2e_shape* s = nullptr;
    DeviceState* d = nullptr;
    int device = 0, format = REC_WIDE;
    uint64_t chunk_tiles = 0;   // tiles per submit (at most)
    uint64_t piece_words = 0;   // != 0: a tile's records do not fit the staging buffer; packed in pieces of at most this many words per lane
    int n_buf = 0;
    cudaStream_t st[2] = {nullptr, nullptr};
    void* d_vals[2] = {nullptr, nullptr};
    void* d_stage[2] = {nullptr, nullptr};
    void* d_in[2] = {nullptr, nullptr};
    u32* d_status[2] = {nullptr, nullptr};
    static const int RING = 64;
    cudaEvent_t ev[RING] = {};
    // The status words come down into pinned staging owned by the stream (one slot per ticket) and are handed to the
    // caller's array when the ticket is reaped: the caller's status array is ordinary pageable memory, and an
    // "asynchronous" copy into pageable memory blocks the submitting thread until the whole chunk has drained, which
    // would serialise the pipeline.
    uint32_t* h_status_stage = nullptr;  // [RING][chunk_tiles * TILE], pinned
    uint32_t* user_status[RING] = {};
    uint64_t user_status_n[RING] = {};
    uint64_t n_submitted = 0;
    uint64_t tile_bytes = 0, tile_wide_bytes = 0;
    bool shared_vals = false;            // both pipeline slots compute into d_vals[0]; a chunk's VM waits for the previous chunk's export kernel
    cudaEvent_t ev_packed[2] = {nullptr, nullptr};
};

static void stream_destroy(h2e_stream* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (int k = 0; k < 2; k++) {
        if (p->st[k]) {
            cudaStreamSynchronize(p->st[k]);
            cudaStreamDestroy(p->st[k]);
        }
        if (k == 0 || !p->shared_vals) cudaFree(p->d_vals[k]);
        if (p->ev_packed[k]) cudaEventDestroy(p->ev_packed[k]);
        cudaFree(p->d_stage[k]);
        cudaFree(p->d_in[k]);
        cudaFree(p->d_status[k]);
    }
    for (int i = 0; i < h2e_stream::RING; i++)
        if (p->ev[i]) cudaEventDestroy(p->ev[i]);
    if (p->h_status_stage) cudaFreeHost(p->h_status_stage);
    delete p;
}

static int stream_open(h2e_shape* s, int device, int format, size_t chunk_bytes_hint, h2e_stream** out) {
    if (bad_format(format)) {
        g_err = "unknown record format";
        return -1;
    }
    DeviceState* d;
    int rc = ensure_device(s, device, &d);
    if (rc) return rc;
    rc = ensure_layout_device(s, d);
    if (rc) return rc;
    const Shape& sh = s->ctx.shape;
    const uint64_t n_slots = sh.slot_cell.size();
    h2e_stream* p = new h2e_stream();
    p->s = s;
    p->d = d;
    p->device = device;
    p->format = format;
    p->tile_wide_bytes = n_slots * TILE * 32;
    p->tile_bytes = s->lay.words_per_lane(format, n_slots) * TILE * 4;
    const uint64_t tile_compact = compact_tile_words(s) * 4;          // the VM writes COMPACT records
    const uint64_t stage_tile = format == REC_COMPACT ? 0 : p->tile_bytes;  // UNIQUE / WIDE are derived into a staging buffer
    // team mode keeps per-launch scratch (inverses handed between macro-ops): 2 KiB per int_div per tile
    uint64_t n_div = 0;
    if (long_program(sh))
        for (const Instr& in : sh.program) n_div += in.op == OP_DIV_CORE;
    const uint64_t per_tile = tile_compact + stage_tile + n_div * TILE * 64 + (uint64_t)sh.n_inputs * TILE * 32 + 4096;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        g_err = "cudaMemGetInfo failed";
        delete p;
        return -2;
    }
    const uint64_t budget = (uint64_t)(free_b * 0.88);
    // chunk size: long programs as many tiles as one team launch takes (SMs / 2); short programs 128 MiB of cells (the pipeline
    // fills in a fraction of a millisecond, and a chunk still is ~25 MB on the bus)
    const char* cm = getenv("H2E_HOST_CHUNK_MB");  // tuning: default chunk of a short program, in MiB of 32-byte cells
    const uint64_t dflt_chunk = (uint64_t)(cm ? std::max(1, atoi(cm)) : 128) << 20;
    uint64_t want = long_program(sh) ? team_group_tiles(d) : std::max<uint64_t>(1, (chunk_bytes_hint ? chunk_bytes_hint : dflt_chunk) / std::max<uint64_t>(p->tile_wide_bytes, 1));
    if (chunk_bytes_hint && long_program(sh)) want = std::min<uint64_t>(want, std::max<uint64_t>(1, chunk_bytes_hint / std::max<uint64_t>(p->tile_wide_bytes, 1)));
    p->n_buf = 2;
    uint64_t fit = budget / (2 * per_tile);
    uint64_t stage_bytes = 0;
    if (fit == 0 && stage_tile && !getenv("H2E_STREAM_SINGLE") && budget >= per_tile + stage_tile + (uint64_t)sh.n_inputs * TILE * 32 + 4096) {
        // Two chunks with their record tiles do not fit (a 4096-point MSM tile: 69 GB of records + 25 GB packed), but one record
        // buffer and TWO staging buffers do: the next tile computes into the shared record buffer as soon as this tile has been
        // packed, while this tile's packed records cross the bus from its own staging buffer. (One buffer of each serialises
        // compute and copy: 0.65 s per tile on that shape = 220 ms + 433 ms. Two record buffers with a small staging buffer do
        // not help: the export kernels cannot run beside the VM, whose CTAs hold every SM's registers.)
        fit = 1;
        p->shared_vals = true;
    }
    if (fit == 0) {
        p->n_buf = 1;
        fit = budget / per_tile;
    }
    if (fit == 0) {
        // not even one tile plus its packed records (4096-point MSM: 155 GB of cells per tile): pack in pieces through a 1 GiB staging buffer
        if (budget < tile_compact + (1ull << 30) + n_div * TILE * 64) {
            g_err = "one 32-instance tile of this shape does not fit in device memory";
            delete p;
            return -1;
        }
        fit = 1;
        if (format != REC_COMPACT) {
            p->piece_words = (1ull << 30) / (TILE * 4);
            stage_bytes = 1ull << 30;
        }
    }
    p->chunk_tiles = std::max<uint64_t>(1, std::min(want, fit));
    if (getenv("H2E_STREAM_SHARED") && stage_tile && p->n_buf == 2) p->shared_vals = true;  // test hook: the shared-record-buffer pipeline on any shape
    if (const char* pw = getenv("H2E_STREAM_PIECE_WORDS")) {
        // test hook: export every chunk in slot-range pieces through a staging buffer of this many words per lane (the path a
        // tile too large for device memory takes), whatever the shape
        if (format != REC_COMPACT && !p->piece_words && !p->shared_vals) {
            p->piece_words = std::max<uint64_t>((uint64_t)atoll(pw), 8 * p->chunk_tiles);
            stage_bytes = p->piece_words * TILE * 4;
        }
    }
    if (!stage_bytes) stage_bytes = p->chunk_tiles * stage_tile;
    auto fail = [&](const char* what) {
        g_err = std::string("stream_open: ") + what + ": " + cudaGetErrorString(cudaGetLastError());
        stream_destroy(p);
        return -2;
    };
    for (int k = 0; k < p->n_buf; k++) {
        if (cudaStreamCreateWithFlags(&p->st[k], cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
        if (k == 1 && p->shared_vals) p->d_vals[1] = p->d_vals[0];
        else if (cudaMalloc(&p->d_vals[k], p->chunk_tiles * tile_compact) != cudaSuccess) return fail("record tiles");
        if (p->shared_vals && cudaEventCreateWithFlags(&p->ev_packed[k], cudaEventDisableTiming) != cudaSuccess) return fail("event");
        if (stage_bytes && cudaMalloc(&p->d_stage[k], stage_bytes) != cudaSuccess) return fail("staging buffer");
        if (cudaMalloc(&p->d_in[k], std::max<uint64_t>(p->chunk_tiles * TILE * sh.n_inputs * 32, 32)) != cudaSuccess) return fail("inputs");
        if (cudaMalloc((void**)&p->d_status[k], p->chunk_tiles * TILE * 4) != cudaSuccess) return fail("status");
    }
    for (int i = 0; i < h2e_stream::RING; i++)
        if (cudaEventCreateWithFlags(&p->ev[i], cudaEventDisableTiming) != cudaSuccess) return fail("event");
    if (cudaHostAlloc((void**)&p->h_status_stage, (size_t)h2e_stream::RING * p->chunk_tiles * TILE * 4, cudaHostAllocDefault) != cudaSuccess)
        return fail("status staging");
    *out = p;
    return 0;
}

// hand a completed ticket's status words to the caller's array (once)
static void stream_reap(h2e_stream* p, uint64_t ticket) {
    const int slot = (int)(ticket % h2e_stream::RING);
    if (!p->user_status[slot]) return;
    memcpy(p->user_status[slot], p->h_status_stage + (size_t)slot * p->chunk_tiles * TILE, p->user_status_n[slot] * 4);
    p->user_status[slot] = nullptr;
}

static int stream_submit(h2e_stream* p, uint64_t n_inst, const void* h_inputs, void* h_records, uint32_t* h_status, uint64_t* ticket) {
    h2e_shape* s = p->s;
    DeviceState* d = p->d;
    const Shape& sh = s->ctx.shape;
    if (n_inst == 0 || n_inst > p->chunk_tiles * TILE) {
        g_err = "h2e_stream_submit: a chunk holds 1 .. " + std::to_string(p->chunk_tiles * TILE) + " instances";
        return -1;
    }
    CUDA_OK(cudaSetDevice(p->device));
    if (p->n_submitted >= (uint64_t)h2e_stream::RING) {
        // ticket ring: the slot being reused must have completed (it has, unless the caller never polled RING chunks back)
        CUDA_OK(cudaEventSynchronize(p->ev[p->n_submitted % h2e_stream::RING]));
        stream_reap(p, p->n_submitted - h2e_stream::RING);
    }
    const int k = (int)(p->n_submitted % p->n_buf);
    cudaStream_t st = p->st[k];
    const uint64_t n_slots = sh.slot_cell.size(), nt = (n_inst + TILE - 1) / TILE;
    const int sms = d->sm_count > 0 ? d->sm_count : 148;
    const size_t in_len = (size_t)n_inst * sh.n_inputs * 32;
    if (in_len) CUDA_OK(cudaMemcpyAsync(p->d_in[k], h_inputs, in_len, cudaMemcpyHostToDevice, st));
    // shared record buffer: the previous chunk (on the other stream) must have been packed into its staging buffer
    if (p->shared_vals && p->n_submitted > 0) CUDA_OK(cudaStreamWaitEvent(st, p->ev_packed[k ^ 1], 0));
    int rc = launch_vm(s, d, st, (u32*)p->d_vals[k], (const u32*)p->d_in[k], p->d_status[k], n_inst);
    if (rc) return rc;
    const u32* d_rec = (const u32*)p->d_vals[k];
    const uint64_t ctw = compact_tile_words(s);
    if (p->format == REC_COMPACT) {
        CUDA_OK(cudaMemcpyAsync(h_records, d_rec, nt * p->tile_bytes, cudaMemcpyDeviceToHost, st));
    } else if (!p->piece_words) {
        if (p->format == REC_WIDE) {
            rc = launch_expand(s, d, st, d_rec, (u32*)p->d_stage[k], nt);
            if (!rc && s->export_format == H2E_EXPORT_MONTGOMERY) rc = launch_montgomery(d, st, (u32*)p->d_stage[k], nt * p->tile_wide_bytes / 32);
        } else {
            rc = launch_pack(s, d, st, p->format, d_rec, (u32*)p->d_stage[k], nt);
        }
        if (rc) return rc;
        if (p->shared_vals) CUDA_OK(cudaEventRecord(p->ev_packed[k], st));
        CUDA_OK(cudaMemcpyAsync(h_records, p->d_stage[k], nt * p->tile_bytes, cudaMemcpyDeviceToHost, st));
    } else {
        // a tile's records do not fit the staging buffer: pieces of consecutive (selected) slots, at most piece_words words per
        // lane and tile each; every piece lands at its offset inside each tile's block of the host buffer (2-D copy: one row per tile)
        const bool uniq = p->format == REC_UNIQUE || p->format == REC_PRIMARY;
        const uint32_t n_sel = uniq ? (uint32_t)s->lay.stored_slots(p->format).size() : (uint32_t)n_slots;
        std::vector<uint32_t> poff(n_sel + 1, 0);  // words per lane before selected slot i, in the output format
        for (uint32_t i = 0; i < n_sel; i++) poff[i + 1] = poff[i] + (uniq ? s->lay.width[s->lay.stored_slots(p->format)[i]] : 8u);
        const uint32_t* off = poff.data();
        uint32_t i0 = 0;
        while (i0 < n_sel) {
            uint32_t i1 = (uint32_t)(std::upper_bound(off + i0, off + n_sel + 1, off[i0] + (uint32_t)std::min<uint64_t>(p->piece_words / nt, 0xffffffffu - off[i0])) - off) - 1;
            if (i1 <= i0) i1 = i0 + 1;
            const uint64_t piece_bytes = (uint64_t)(off[i1] - off[i0]) * TILE * 4;
            g_launches++;
            if (uniq) {
                CUDA_OK(vm_pack(st, (unsigned)sms * 8, d_rec, (u32*)p->d_stage[k], d->d_sel_packed(p->format), d->d_off_packed(p->format), d->d_off_compact, ctw, i0, i1 - i0, nt, piece_bytes / 4));
            } else {
                CUDA_OK(vm_expand(st, (unsigned)sms * 8, d_rec, (u32*)p->d_stage[k], d->d_off_compact, ctw, i0, i1 - i0, nt, piece_bytes / 4));
                if (s->export_format == H2E_EXPORT_MONTGOMERY) {
                    rc = launch_montgomery(d, st, (u32*)p->d_stage[k], nt * piece_bytes / 32);
                    if (rc) return rc;
                }
            }
            CUDA_OK(cudaMemcpy2DAsync((char*)h_records + (uint64_t)off[i0] * TILE * 4, p->tile_bytes, p->d_stage[k], piece_bytes, piece_bytes, nt,
                                      cudaMemcpyDeviceToHost, st));
            i0 = i1;
        }
    }
    {
        const int slot = (int)(p->n_submitted % h2e_stream::RING);
        CUDA_OK(cudaMemcpyAsync(p->h_status_stage + (size_t)slot * p->chunk_tiles * TILE, p->d_status[k], n_inst * 4, cudaMemcpyDeviceToHost, st));
        p->user_status[slot] = h_status;
        p->user_status_n[slot] = n_inst;
    }
    CUDA_OK(cudaEventRecord(p->ev[p->n_submitted % h2e_stream::RING], st));
    if (ticket) *ticket = p->n_submitted;
    p->n_submitted++;
    return 0;
}

static int stream_wait(h2e_stream* p, uint64_t ticket, bool block) {
    if (ticket >= p->n_submitted) {
        g_err = "unknown ticket";
        return -1;
    }
    if (ticket + h2e_stream::RING < p->n_submitted) return 0;  // its ring slot was reused, which waited for it
    cudaEvent_t e = p->ev[ticket % h2e_stream::RING];
    if (block) {
        CUDA_OK(cudaEventSynchronize(e));
        stream_reap(p, ticket);
        return 0;
    }
    cudaError_t q = cudaEventQuery(e);
    if (q == cudaSuccess) {
        stream_reap(p, ticket);
        return 0;
    }
    if (q == cudaErrorNotReady) return 1;
    g_err = std::string("stream poll: ") + cudaGetErrorString(q);
    return -2;
}

// whole batch through a cached pipeline (the one-call host entry points)
static int run_host_batch(h2e_shape* s, int device, int format, uint64_t n_inst, const void* h_inputs, void* h_records, uint32_t* h_status) {
    if (n_inst == 0) return 0;
    DeviceState* d;
    int rc = ensure_device(s, device, &d);
    if (rc) return rc;
    if (!d->host_pipe[format]) {
        rc = stream_open(s, device, format, 0, &d->host_pipe[format]);
        if (rc) return rc;
    }
    h2e_stream* p = d->host_pipe[format];
    const Shape& sh = s->ctx.shape;
    const uint64_t chunk = p->chunk_tiles * TILE;
    uint64_t last = 0;
    for (uint64_t i0 = 0; i0 < n_inst; i0 += chunk) {
        const uint64_t ni = std::min(chunk, n_inst - i0);
        rc = stream_submit(p, ni, (const char*)h_inputs + i0 * sh.n_inputs * 32, (char*)h_records + (i0 / TILE) * p->tile_bytes, h_status + i0, &last);
        if (rc) return rc;
    }
    for (int k = 0; k < p->n_buf; k++) CUDA_OK(cudaStreamSynchronize(p->st[k]));
    for (uint64_t t = p->n_submitted > (uint64_t)h2e_stream::RING ? p->n_submitted - h2e_stream::RING : 0; t < p->n_submitted; t++) stream_reap(p, t);
    return 0;
}

extern "C" {

h2e_stream* h2e_stream_open(h2e_shape* s, int device, int format, size_t chunk_bytes_hint) {
    h2e_stream* p = nullptr;
    if (stream_open(s, device, format, chunk_bytes_hint, &p)) return nullptr;
    return p;
}
int h2e_stream_query(const h2e_stream* p, uint64_t out[8]) {
    out[0] = p->chunk_tiles * TILE;              // instances per chunk (at most)
    out[1] = p->chunk_tiles * p->tile_bytes;     // bytes of a full chunk's records in the stream's format
    out[2] = p->tile_bytes;                      // bytes per 32-instance tile
    out[3] = (uint64_t)p->n_buf;                 // chunks in flight on the device
    out[4] = p->piece_words ? 1 : 0;             // 1: tiles are exported in slot-range pieces
    out[5] = (uint64_t)h2e_stream::RING;         // tickets that may be outstanding
    out[6] = p->n_submitted;
    out[7] = 0;
    return 0;
}
int h2e_stream_submit(h2e_stream* p, uint64_t n_inst, const void* h_inputs, void* h_records, uint32_t* h_status, uint64_t* ticket) {
    return stream_submit(p, n_inst, h_inputs, h_records, h_status, ticket);
}
int h2e_stream_poll(h2e_stream* p, uint64_t ticket) { return stream_wait(p, ticket, false); }
int h2e_stream_wait(h2e_stream* p, uint64_t ticket) { return stream_wait(p, ticket, true); }
int h2e_stream_close(h2e_stream* p) {
    stream_destroy(p);
    return 0;
}

int h2e_batch_run_host(h2e_shape* s, int device, uint64_t n_inst, const void* h_inputs, void* h_vals, uint32_t* h_status) {
    return run_host_batch(s, device, REC_WIDE, n_inst, h_inputs, h_vals, h_status);
}
int h2e_batch_run_host_compact(h2e_shape* s, int device, uint64_t n_inst, const void* h_inputs, void* h_compact, uint32_t* h_status) {
    return run_host_batch(s, device, REC_COMPACT, n_inst, h_inputs, h_compact, h_status);
}
int h2e_batch_run_host_records(h2e_shape* s, int device, int format, uint64_t n_inst, const void* h_inputs, void* h_records, uint32_t* h_status) {
    if (bad_format(format)) {
        g_err = "unknown record format";
        return -1;
    }
    return run_host_batch(s, device, format, n_inst, h_inputs, h_records, h_status);
}

// Consumer side (host): records in `format` -> plain cells. mode 0: vals[tile][slot][lane][32 bytes] (the WIDE
// layout); mode 1 / 2: one dense cell array per instance, out[instance][cell][32 bytes], cell = the slot's advice
// cell in column-major (1: region, column, row -- the prover's advice columns) or row-major (2: region, row,
// column -- RecordsInner, context.rs:241-252) order over the regions' heights; cells no slot maps to are zeroed.
// Copies are filled from their roots on the way. `n_threads` host threads, one contiguous range of tiles each.
int h2e_records_expand(h2e_shape* s, int format, int mode, uint64_t n_inst, const void* h_records, void* h_out, int n_threads) {
    if (bad_format(format) || mode < 0 || mode > 2) {
        g_err = "unknown record format / expansion mode";
        return -1;
    }
    if (ensure_layout(s)) return -1;
    const Shape& sh = s->ctx.shape;
    const Layout& lay = s->lay;
    const uint64_t n_slots = sh.slot_cell.size(), tiles = (n_inst + TILE - 1) / TILE;
    const uint64_t tile_words = lay.words_per_lane(format, n_slots) * TILE;
    std::vector<uint32_t> wide_off;
    const uint32_t* off;
    if (format == REC_WIDE) {
        wide_off.resize(n_slots + 1);
        for (uint64_t i = 0; i <= n_slots; i++) wide_off[i] = (uint32_t)(8 * i);
        off = wide_off.data();
    } else {
        off = lay.off(format).data();
    }
    std::vector<uint64_t> dst;
    uint64_t cells_per_inst = 0;
    if (mode != 0 && h2e_shape_dense_cells(s) > 0xffffffffull) {
        g_err = "dense cell index exceeds 32 bits";
        return -1;
    }
    if (mode != 0) {
        uint64_t base[3];
        for (int r = 0; r < 3; r++) {
            base[r] = cells_per_inst;
            cells_per_inst += (uint64_t)ADV_COLS[r] * sh.height[r];
        }
        dst.resize(n_slots);
        for (uint64_t i = 0; i < n_slots; i++) {
            const Cell& c = sh.slot_cell[i];
            dst[i] = base[c.region] + (mode == 1 ? (uint64_t)c.col * sh.height[c.region] + c.row : (uint64_t)c.row * ADV_COLS[c.region] + c.col);
        }
    }
    const uint32_t* src = (const uint32_t*)h_records;
    uint32_t* out = (uint32_t*)h_out;
    const bool packed = format == REC_UNIQUE || format == REC_PRIMARY;
    if (n_threads < 1) n_threads = 1;
    n_threads = (int)std::min<uint64_t>((uint64_t)n_threads, std::max<uint64_t>(tiles, 1));
    // where slot sl's value is stored in `format`, and how to rebuild it: (words per lane before it, width, bit-field shift)
    struct Src {
        uint32_t off;
        uint8_t w, derived, shift;
    };
    auto source_of = [&](uint64_t sl) {
        uint32_t r = packed ? lay.root[sl] : (uint32_t)sl;
        // PRIMARY: a range chunk is rebuilt from the stored cell it is a bit field of
        const bool derived = format == REC_PRIMARY && lay.der_src[r] != NONE;
        const uint8_t shift = derived ? lay.der_shift[r] : 0;
        if (derived) r = lay.der_src[r];
        return Src{off[r], (uint8_t)(format == REC_WIDE ? 8u : lay.width[r]), (uint8_t)derived, shift};
    };
    auto load_cell = [](uint32_t* q, const uint32_t* p, const Src& c) {  // p = the stored cell of this lane
        if (c.derived) {
            uint32_t v = 0;
            if (c.shift != DER_ZERO) {
                const uint32_t wi = c.shift >> 5, sh = c.shift & 31;
                const uint32_t lo = wi < c.w ? p[wi] : 0, hi = wi + 1u < c.w ? p[wi + 1] : 0;
                v = (uint32_t)((((uint64_t)hi << 32) | lo) >> sh) & 0x3ffffu;
            }
            q[0] = v;
            for (uint32_t k2 = 1; k2 < 8; k2++) q[k2] = 0;
            return;
        }
        for (uint32_t k2 = 0; k2 < c.w; k2++) q[k2] = p[k2];
        for (uint32_t k2 = c.w; k2 < 8; k2++) q[k2] = 0;
    };
    std::function<void(uint64_t, uint64_t)> work;
    if (mode == 0) {
        work = [&](uint64_t t0, uint64_t t1) {
            for (uint64_t tile = t0; tile < t1; tile++) {
                const uint32_t* tsrc = src + tile * tile_words;
                for (uint64_t sl = 0; sl < n_slots; sl++) {
                    const Src c = source_of(sl);
                    const uint32_t* p = tsrc + (uint64_t)c.off * TILE;
                    uint32_t* q = out + (tile * n_slots + sl) * TILE * 8;
                    for (unsigned lane = 0; lane < (unsigned)TILE; lane++) load_cell(q + lane * 8, p + lane * c.w, c);
                }
            }
        };
    } else {
        // Dense per-instance arrays. The records are instance-minor, the output instance-major: slots are taken in the order of
        // their destination cell, a block of them at a time (its source lines stay in cache while the 32 lanes pass over it),
        // so every instance receives contiguous runs, written with streaming stores; unassigned cells are zeroed on the way.
        std::vector<h2e_shape::ExpandItem>& plan = s->expand_plan[mode - 1];
        {
            std::lock_guard<std::mutex> lk(s->mu);
            if (plan.size() != n_slots) {
                plan.resize(n_slots);
                for (uint64_t i = 0; i < n_slots; i++) plan[i] = h2e_shape::ExpandItem{(uint32_t)dst[i], (uint32_t)i};
                std::sort(plan.begin(), plan.end(), [](const h2e_shape::ExpandItem& a, const h2e_shape::ExpandItem& b) { return a.dst < b.dst; });
            }
        }
        const bool aligned = ((uintptr_t)out % 16) == 0;
        work = [&, aligned](uint64_t t0, uint64_t t1) {
            const uint64_t BLOCK = 48;  // the block's source lines (<= 48 KB, typically ~16 KB) stay in L1 over the 32 lane passes
            std::vector<Src> srcs(BLOCK);
            const __m128i z = _mm_setzero_si128();
            auto put = [aligned](uint32_t* q, __m128i lo, __m128i hi) {
                if (aligned) {  // streaming stores: the dense arrays are written once and read by somebody else
                    _mm_stream_si128((__m128i*)q, lo);
                    _mm_stream_si128((__m128i*)q + 1, hi);
                } else {
                    _mm_storeu_si128((__m128i*)q, lo);
                    _mm_storeu_si128((__m128i*)q + 1, hi);
                }
            };
            for (uint64_t tile = t0; tile < t1; tile++) {
                const uint32_t* tsrc = src + tile * tile_words;
                const unsigned lanes = (unsigned)std::min<uint64_t>(TILE, n_inst - tile * TILE);
                for (uint64_t k0 = 0; k0 < n_slots; k0 += BLOCK) {
                    const uint64_t k1 = std::min(n_slots, k0 + BLOCK);
                    for (uint64_t k = k0; k < k1; k++) srcs[k - k0] = source_of(plan[k].slot);
                    const uint64_t first = k0 == 0 ? 0 : (uint64_t)plan[k0 - 1].dst + 1;  // cells before the block that nobody assigns
                    const uint64_t last = k1 == n_slots ? cells_per_inst : 0;               // ... and after the last slot
                    for (unsigned lane = 0; lane < lanes; lane++) {
                        uint32_t* base = out + (tile * TILE + lane) * cells_per_inst * 8;
                        uint64_t next = first;
                        for (uint64_t k = k0; k < k1; k++) {
                            const uint64_t d = plan[k].dst;
                            for (; next < d; next++) put(base + next * 8, z, z);
                            const Src& c = srcs[k - k0];
                            const uint32_t* p = tsrc + (uint64_t)c.off * TILE + lane * c.w;
                            __m128i lo, hi = z;
                            if (c.derived) {
                                uint32_t v = 0;
                                if (c.shift != DER_ZERO) {
                                    const uint32_t wi = c.shift >> 5, sh = c.shift & 31;
                                    const uint32_t l0 = wi < c.w ? p[wi] : 0, h0 = wi + 1u < c.w ? p[wi + 1] : 0;
                                    v = (uint32_t)((((uint64_t)h0 << 32) | l0) >> sh) & 0x3ffffu;
                                }
                                lo = _mm_cvtsi32_si128((int)v);
                            } else if (c.w == 1) {
                                lo = _mm_cvtsi32_si128((int)p[0]);
                            } else {
                                lo = _mm_loadu_si128((const __m128i*)p);
                                if (c.w == 8) hi = _mm_loadu_si128((const __m128i*)p + 1);
                            }
                            put(base + d * 8, lo, hi);
                            next = d + 1;
                        }
                        for (; next < last; next++) put(base + next * 8, z, z);
                    }
                }
            }
            _mm_sfence();
        };
    }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, tiles * t / n_threads, tiles * (t + 1) / n_threads);
    for (auto& t : th) t.join();
    return 0;
}
int h2e_expand_compact(const h2e_shape* s, uint64_t n_inst, const void* h_compact, void* h_vals, int n_threads) {
    return h2e_records_expand(const_cast<h2e_shape*>(s), REC_COMPACT, 0, n_inst, h_compact, h_vals, n_threads);
}
// cells of one instance's dense array in modes 1 / 2 of h2e_records_expand and of h2e_records_scatter
uint64_t h2e_shape_dense_cells(const h2e_shape* s) {
    const Shape& sh = s->ctx.shape;
    uint64_t n = 0;
    for (int r = 0; r < 3; r++) n += (uint64_t)ADV_COLS[r] * sh.height[r];
    return n;
}

// Prover hand-off on the DEVICE: COMPACT records (d_records, as filled by h2e_batch_run_records for n_inst instances) -> one
// dense cell array per instance in d_out, out[inst0 + instance][cell][32 bytes] (order 1 = column-major, 2 = row-major as
// above; encoding canonical or Montgomery). d_out must hold (inst0 + n_inst) * h2e_shape_dense_cells cells and be
// zeroed by the caller where unassigned cells matter. Asynchronous on `stream`.
int h2e_records_scatter(h2e_shape* s, int device, void* stream, uint64_t n_inst, const void* d_records, void* d_out, uint64_t inst0, int order, int encoding) {
    if (order != 1 && order != 2) {
        g_err = "order must be 1 (column-major) or 2 (row-major)";
        return -1;
    }
    if (n_inst == 0) return 0;
    DeviceState* d;
    int rc = ensure_device(s, device, &d);
    if (rc) return rc;
    rc = ensure_layout_device(s, d);
    if (rc) return rc;
    const Shape& sh = s->ctx.shape;
    const uint64_t n_slots = sh.slot_cell.size();
    {
        std::lock_guard<std::mutex> lk(s->mu);
        if (!d->d_scatter_dst[order - 1]) {
            uint64_t base[3], tot = 0;
            for (int r = 0; r < 3; r++) {
                base[r] = tot;
                tot += (uint64_t)ADV_COLS[r] * sh.height[r];
            }
            if (tot > 0xffffffffull) {
                g_err = "dense cell index exceeds 32 bits";
                return -1;
            }
            std::vector<uint32_t> dst(std::max<uint64_t>(n_slots, 1));
            for (uint64_t i = 0; i < n_slots; i++) {
                const Cell& c = sh.slot_cell[i];
                dst[i] = (uint32_t)(base[c.region] + (order == 1 ? (uint64_t)c.col * sh.height[c.region] + c.row : (uint64_t)c.row * ADV_COLS[c.region] + c.col));
            }
            std::vector<uint32_t> ord(dst.size());
            std::iota(ord.begin(), ord.end(), 0u);
            std::sort(ord.begin(), ord.begin() + n_slots, [&](uint32_t a, uint32_t b) { return dst[a] < dst[b]; });
            CUDA_OK(cudaMalloc(&d->d_scatter_ord[order - 1], ord.size() * 4));
            CUDA_OK(cudaMemcpy(d->d_scatter_ord[order - 1], ord.data(), ord.size() * 4, cudaMemcpyHostToDevice));
            CUDA_OK(cudaMalloc(&d->d_scatter_dst[order - 1], dst.size() * 4));
            CUDA_OK(cudaMemcpy(d->d_scatter_dst[order - 1], dst.data(), dst.size() * 4, cudaMemcpyHostToDevice));
        }
    }
    const int sms = d->sm_count > 0 ? d->sm_count : 148;
    g_launches++;
    // (one CTA = one 32-slot x 32-instance block through 33 KB of shared memory: 6 CTAs resident per SM)
    const uint64_t items = ((n_slots + 31) / 32) * ((n_inst + TILE - 1) / TILE);
    const char* sc = getenv("H2E_SCATTER_CTAS");  // tuning: CTAs per SM
    const uint64_t per_sm = sc ? (uint64_t)std::max(1, atoi(sc)) : 12;  // measured on 32 bn256 pairing instances: 2 CTAs per SM 5.5 ms, 6 4.6 ms, 12 4.0 ms
    CUDA_OK(vm_scatter((cudaStream_t)stream, (unsigned)std::min<uint64_t>(items, (uint64_t)sms * per_sm), (const u32*)d_records, (u32*)d_out, d->d_scatter_dst[order - 1],
                       d->d_scatter_ord[order - 1], d->d_off_compact, compact_tile_words(s), n_slots, inst0, n_inst, h2e_shape_dense_cells(s),
                       encoding == H2E_EXPORT_MONTGOMERY ? 1 : 0));
    return 0;
}

int h2e_shape_set_export(h2e_shape* s, int format) {
    if (format != H2E_EXPORT_CANONICAL && format != H2E_EXPORT_MONTGOMERY) {
        g_err = "unknown export format";
        return -1;
    }
    s->export_format = format;
    return 0;
}

int h2e_cells_to_montgomery(h2e_shape* s, int device, void* stream, void* d_cells, uint64_t n_cells) {
    DeviceState* d;
    int rc = ensure_device(s, device, &d);
    if (rc) return rc;
    return launch_montgomery(d, (cudaStream_t)stream, (u32*)d_cells, n_cells);
}

int h2e_measure_imad_peak(int device, double* imad_per_sec) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        g_err = "no CUDA device available";
        return -3;
    }
    CUDA_OK(cudaSetDevice(device));
    int sms = 0;
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    uint64_t* d_out = nullptr;
    CUDA_OK(cudaMalloc(&d_out, 8));
    const unsigned blocks = (unsigned)sms * 8;
    const uint32_t iters = 1 << 16;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_OK(cudaEventRecord(e0, 0));
        CUDA_OK(vm_imad_probe(0, blocks, d_out, iters));
        CUDA_OK(cudaEventRecord(e1, 0));
        CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::max(best, (double)blocks * 256.0 * 8.0 * iters / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    g_launches += 4;
    *imad_per_sec = best;
    return 0;
}

int h2e_shape_set_mode(h2e_shape* s, int mode, int ctas_per_tile) {
    s->force_mode = mode & 0xff;
    s->force_crit = (mode >> 8) & 0xff;  // tuning: bits 8..15 = critical warps per CTA
    s->force_warps = (mode >> 16) & 0xff;  // bits 16..23 = warps per CTA (8 or 16)
    s->force_ctas = ctas_per_tile;
    std::lock_guard<std::mutex> lk(s->mu);
    for (auto& kv : s->dev) {  // streams depend on the split: rebuild on next launch
        for (auto& t : kv.second.team) cudaFree(t.second.blob);
        kv.second.team.clear();
    }
    return 0;
}

}  // extern "C"
