// C-ABI implementation (include/h2ecc_b200.h): shapes, schedules, device state and launches. The kernels
// live in vm_kernel.cu.
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <string>
#include <thread>

#include "../../include/h2ecc_b200.h"
#include "circuits.h"
#include "schedule.h"
#include "script_builder.h"
#include "fieldinfo.h"
#include "vm_kernel.h"

using namespace h2e;

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
    // workspace of the host-buffer entry point (h2e_batch_run_host), kept across calls: two chunk
    // buffers + two streams (double buffering), inputs, status
    cudaStream_t ws_stream[2] = {nullptr, nullptr};
    void* ws_vals[2] = {nullptr, nullptr};
    size_t ws_vals_cap = 0;
    void* ws_in = nullptr;
    size_t ws_in_cap = 0;
    u32* ws_status = nullptr;
    size_t ws_status_cap = 0;
    // compact export
    u32* d_compact_off = nullptr;
    void* ws_compact[2] = {nullptr, nullptr};
    size_t ws_compact_cap = 0;
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
    std::vector<uint32_t> compact_off;  // [n_slots + 1] prefix sums of the slots' width classes (words per lane); empty until probed
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

static int ensure_device(h2e_shape* s, int device, DeviceState** out) {
    std::lock_guard<std::mutex> lk(s->mu);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_err = "no CUDA device available: the witness VM has no CPU fallback";
        return -3;
    }
    CUDA_OK(cudaSetDevice(device));
    DeviceState& d = s->dev[device];
    if (!d.d_prog) {
        const Shape& sh = s->ctx.shape;
        size_t np = std::max<size_t>(sh.program.size(), 1), nc = std::max<size_t>(sh.consts.size(), 1);
        CUDA_OK(cudaMalloc(&d.d_prog, np * sizeof(Instr)));
        CUDA_OK(cudaMalloc(&d.d_cpool, nc * 32));
        CUDA_OK(cudaMalloc(&d.d_tables, std::max<size_t>(sh.tables.size(), 1) * 4));
        if (!sh.tables.empty()) CUDA_OK(cudaMemcpy(d.d_tables, sh.tables.data(), sh.tables.size() * 4, cudaMemcpyHostToDevice));
        if (!sh.program.empty()) CUDA_OK(cudaMemcpy(d.d_prog, sh.program.data(), sh.program.size() * sizeof(Instr), cudaMemcpyHostToDevice));
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
static int launch_vm(h2e_shape* s, DeviceState* d, cudaStream_t stream, u32* d_vals, const u32* d_inputs, u32* d_status, uint64_t n_inst) {
    const Shape& sh = s->ctx.shape;
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
                      sh.slot_cell.size(), (uint32_t)sh.n_inputs, n_inst, tiles, 0};
        g_launches++;
        CUDA_OK(getenv("H2E_THREAD_W16") ? vm_launch_w16(L) : vm_launch_w8(L));  // (tuning switch; the 255-register build is the default)
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
                  s->sched.n_scratch, sh.slot_cell.size(), (uint32_t)sh.n_inputs, n_inst, tiles, s->force_mode >= 3 ? s->force_mode : 1};
    g_launches++;
    CUDA_OK(warps == 16 ? vm_launch_w16(L) : vm_launch_w8(L));
    CUDA_OK(cudaFreeAsync(d_progress, stream));
    return 0;
}

// Chunk buffers of the host entry points: two (double buffering) when they fit, one when a single chunk already
// takes more than half of the free memory (a 4096-point MSM tile is 155 GB).
static int ensure_ws_vals(DeviceState* d, size_t chunk_bytes) {
    if (d->ws_vals_cap >= chunk_bytes) return 0;
    for (int k = 0; k < 2; k++) {
        cudaFree(d->ws_vals[k]);
        d->ws_vals[k] = nullptr;
    }
    d->ws_vals_cap = 0;
    CUDA_OK(cudaMalloc(&d->ws_vals[0], chunk_bytes));
    if (cudaMalloc(&d->ws_vals[1], chunk_bytes) != cudaSuccess) {
        cudaGetLastError();  // clear the allocation failure: single-buffer mode
        d->ws_vals[1] = nullptr;
    }
    d->ws_vals_cap = chunk_bytes;
    return 0;
}

// Static width class of every slot (compact export). The widths are fixed by the macro-op code (which store
// a call site uses), so they are read off the device: the width-probe build of the VM runs the program once,
// in thread mode, with every store writing the width class of its cell instead of its value.
static int ensure_compact(h2e_shape* s, DeviceState* d) {
    const Shape& sh = s->ctx.shape;
    const size_t n_slots = sh.slot_cell.size();
    if (s->compact_off.empty()) {
        u32 *d_cells = nullptr, *d_in = nullptr, *d_status = nullptr;
        CUDA_OK(cudaMalloc(&d_cells, std::max<size_t>(n_slots, 1) * 32));
        CUDA_OK(cudaMemset(d_cells, 0, std::max<size_t>(n_slots, 1) * 32));
        CUDA_OK(cudaMalloc(&d_in, std::max<size_t>(sh.n_inputs, 1) * 32));
        CUDA_OK(cudaMemset(d_in, 0, std::max<size_t>(sh.n_inputs, 1) * 32));
        CUDA_OK(cudaMalloc(&d_status, TILE * 4));
        TeamProg flat = {};
        flat.crit = d->d_prog;
        flat.n_levels = (uint32_t)sh.program.size();
        VmLaunch L = {1u, (unsigned)TILE, 0, flat, d_cells, d_in, d->d_cpool, d->d_tables, d_status, nullptr, nullptr, 0, n_slots, (uint32_t)sh.n_inputs,
                      1, 1, 0};
        g_launches++;
        CUDA_OK(vm_launch_wprobe(L));
        std::vector<uint32_t> cells(n_slots * 8);
        CUDA_OK(cudaMemcpy(cells.data(), d_cells, n_slots * 32, cudaMemcpyDeviceToHost));
        cudaFree(d_cells);
        cudaFree(d_in);
        cudaFree(d_status);
        std::vector<uint32_t> off(n_slots + 1, 0);
        for (size_t i = 0; i < n_slots; i++) {
            uint32_t w = cells[8 * i];
            if (w != 1 && w != 4 && w != 8) {
                g_err = "width probe: slot " + std::to_string(i) + " was not written by the program";
                return -1;
            }
            off[i + 1] = off[i] + w;
        }
        s->compact_off.swap(off);
    }
    if (!d->d_compact_off) {
        CUDA_OK(cudaMalloc(&d->d_compact_off, s->compact_off.size() * 4));
        CUDA_OK(cudaMemcpy(d->d_compact_off, s->compact_off.data(), s->compact_off.size() * 4, cudaMemcpyHostToDevice));
    }
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
            for (int k = 0; k < 2; k++) {
                if (kv.second.ws_stream[k]) cudaStreamDestroy(kv.second.ws_stream[k]);
                cudaFree(kv.second.ws_vals[k]);
            }
            cudaFree(kv.second.ws_in);
            cudaFree(kv.second.ws_status);
            cudaFree(kv.second.d_compact_off);
            cudaFree(kv.second.ws_compact[0]);
            cudaFree(kv.second.ws_compact[1]);
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

int h2e_batch_run(h2e_shape* s, int device, void* stream, uint64_t n_inst, const void* d_inputs, void* d_vals, uint32_t* d_status) {
    if (n_inst == 0) return 0;
    DeviceState* d;
    int rc = ensure_device(s, device, &d);
    if (rc) return rc;
    return launch_vm(s, d, (cudaStream_t)stream, (u32*)d_vals, (const u32*)d_inputs, d_status, n_inst);
}

// ---- compact export -------------------------------------------------------------------------
int h2e_compact_prepare(h2e_shape* s, int device) {
    DeviceState* d;
    int rc = ensure_device(s, device, &d);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(s->mu);
    return ensure_compact(s, d);
}
size_t h2e_compact_bytes(const h2e_shape* s, uint64_t n_inst) {
    if (s->compact_off.empty()) return 0;
    return (size_t)(pad_tiles(n_inst) / TILE) * s->compact_off.back() * TILE * 4;
}
int h2e_compact_widths(const h2e_shape* s, uint8_t* out) {
    if (s->compact_off.empty()) {
        g_err = "call h2e_compact_prepare first";
        return -1;
    }
    for (size_t i = 0; i + 1 < s->compact_off.size(); i++) out[i] = (uint8_t)(s->compact_off[i + 1] - s->compact_off[i]);
    return 0;
}
int h2e_batch_run_host_compact(h2e_shape* s, int device, uint64_t n_inst, const void* h_inputs, void* h_compact, uint32_t* h_status) {
    if (n_inst == 0) return 0;
    DeviceState* d;
    int rc = ensure_device(s, device, &d);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        rc = ensure_compact(s, d);
        if (rc) return rc;
    }
    const Shape& sh = s->ctx.shape;
    const uint64_t n_slots = sh.slot_cell.size();
    const uint64_t tile_bytes = n_slots * TILE * 32, ctile_bytes = (uint64_t)s->compact_off.back() * TILE * 4;
    uint64_t tiles = (n_inst + TILE - 1) / TILE;
    uint64_t tiles_per_chunk = std::max<uint64_t>(1, std::min<uint64_t>(tiles, (256ull << 20) / std::max<uint64_t>(tile_bytes, 1)));
    const size_t chunk_bytes = tiles_per_chunk * tile_bytes, cchunk_bytes = tiles_per_chunk * ctile_bytes,
                 in_bytes = h2e_inputs_bytes(s, n_inst), st_bytes = pad_tiles(n_inst) * 4;
    for (int k = 0; k < 2; k++)
        if (!d->ws_stream[k]) CUDA_OK(cudaStreamCreateWithFlags(&d->ws_stream[k], cudaStreamNonBlocking));
    rc = ensure_ws_vals(d, chunk_bytes);
    if (rc) return rc;
    const int n_buf = d->ws_vals[1] ? 2 : 1;
    if (d->ws_compact_cap < cchunk_bytes) {
        for (int k = 0; k < 2; k++) {
            cudaFree(d->ws_compact[k]);
            d->ws_compact[k] = nullptr;
        }
        d->ws_compact_cap = 0;
        for (int k = 0; k < 2; k++) CUDA_OK(cudaMalloc(&d->ws_compact[k], cchunk_bytes));
        d->ws_compact_cap = cchunk_bytes;
    }
    if (d->ws_in_cap < std::max<size_t>(in_bytes, 32)) {
        cudaFree(d->ws_in);
        d->ws_in = nullptr;
        d->ws_in_cap = 0;
        CUDA_OK(cudaMalloc(&d->ws_in, std::max<size_t>(in_bytes, 32)));
        d->ws_in_cap = std::max<size_t>(in_bytes, 32);
    }
    if (d->ws_status_cap < st_bytes) {
        cudaFree(d->ws_status);
        d->ws_status = nullptr;
        d->ws_status_cap = 0;
        CUDA_OK(cudaMalloc(&d->ws_status, st_bytes));
        d->ws_status_cap = st_bytes;
    }
    cudaStream_t* st = d->ws_stream;
    int sms = d->sm_count > 0 ? d->sm_count : 148;
    int k = 0;
    for (uint64_t t0 = 0; t0 < tiles; t0 += tiles_per_chunk, k = (k + 1) % n_buf) {
        uint64_t nt = std::min(tiles_per_chunk, tiles - t0);
        uint64_t i0 = t0 * TILE, ni = std::min<uint64_t>(n_inst - i0, nt * TILE);
        const size_t in_off = (size_t)i0 * sh.n_inputs * 32, in_len = (size_t)ni * sh.n_inputs * 32;
        if (in_len) CUDA_OK(cudaMemcpyAsync((char*)d->ws_in + in_off, (const char*)h_inputs + in_off, in_len, cudaMemcpyHostToDevice, st[k]));
        rc = launch_vm(s, d, st[k], (u32*)d->ws_vals[k], (const u32*)d->ws_in + i0 * sh.n_inputs * 8, d->ws_status + i0, ni);
        if (rc) return rc;
        g_launches++;
        CUDA_OK(vm_pack(st[k], (unsigned)sms * 8, (const u32*)d->ws_vals[k], (u32*)d->ws_compact[k], d->d_compact_off, n_slots, nt));
        CUDA_OK(cudaMemcpyAsync((char*)h_compact + t0 * ctile_bytes, d->ws_compact[k], nt * ctile_bytes, cudaMemcpyDeviceToHost, st[k]));
        CUDA_OK(cudaMemcpyAsync(h_status + i0, d->ws_status + i0, ni * 4, cudaMemcpyDeviceToHost, st[k]));
    }
    CUDA_OK(cudaStreamSynchronize(st[0]));
    CUDA_OK(cudaStreamSynchronize(st[1]));
    return 0;
}
// Host-side expansion of the compact form into full 32-byte cells (what the Rust shim does while it scatters
// cells into RecordsInner; provided here for tests and for bindings that want the plain layout).
int h2e_expand_compact(const h2e_shape* s, uint64_t n_inst, const void* h_compact, void* h_vals, int n_threads) {
    if (s->compact_off.empty()) {
        g_err = "call h2e_compact_prepare first";
        return -1;
    }
    const std::vector<uint32_t>& off = s->compact_off;
    const uint64_t n_slots = off.size() - 1, tiles = (n_inst + TILE - 1) / TILE;
    const uint64_t ctile_words = (uint64_t)off.back() * TILE;
    const uint32_t* src = (const uint32_t*)h_compact;
    uint32_t* dst = (uint32_t*)h_vals;
    if (n_threads < 1) n_threads = 1;
    const uint64_t total = tiles * n_slots;
    auto work = [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; i++) {
            const uint64_t tile = i / n_slots, sl = i % n_slots;
            const uint32_t o = off[sl], w = off[sl + 1] - o;
            const uint32_t* p = src + tile * ctile_words + (uint64_t)o * TILE;
            uint32_t* q = dst + i * TILE * 8;
            for (unsigned lane = 0; lane < (unsigned)TILE; lane++) {
                for (uint32_t k2 = 0; k2 < w; k2++) q[lane * 8 + k2] = p[lane * w + k2];
                for (uint32_t k2 = w; k2 < 8; k2++) q[lane * 8 + k2] = 0;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, total * t / n_threads, total * (t + 1) / n_threads);
    for (auto& t : th) t.join();
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

int h2e_batch_run_host(h2e_shape* s, int device, uint64_t n_inst, const void* h_inputs, void* h_vals, uint32_t* h_status) {
    if (n_inst == 0) return 0;
    DeviceState* d;
    int rc = ensure_device(s, device, &d);
    if (rc) return rc;
    const Shape& sh = s->ctx.shape;
    // chunks of whole tiles, double-buffered: chunk k+1 computes while chunk k is copied out
    const uint64_t tile_bytes = (uint64_t)sh.slot_cell.size() * TILE * 32;
    uint64_t tiles = (n_inst + TILE - 1) / TILE;
    uint64_t tiles_per_chunk = std::max<uint64_t>(1, std::min<uint64_t>(tiles, (256ull << 20) / std::max<uint64_t>(tile_bytes, 1)));
    // device workspace: grown on demand, kept in the handle (no per-call cudaMalloc/cudaFree)
    const size_t chunk_bytes = tiles_per_chunk * tile_bytes, in_bytes = h2e_inputs_bytes(s, n_inst), st_bytes = pad_tiles(n_inst) * 4;
    for (int k = 0; k < 2; k++)
        if (!d->ws_stream[k]) CUDA_OK(cudaStreamCreateWithFlags(&d->ws_stream[k], cudaStreamNonBlocking));
    rc = ensure_ws_vals(d, chunk_bytes);
    if (rc) return rc;
    const int n_buf = d->ws_vals[1] ? 2 : 1;
    if (d->ws_in_cap < std::max<size_t>(in_bytes, 32)) {
        cudaFree(d->ws_in);
        d->ws_in = nullptr;
        d->ws_in_cap = 0;
        CUDA_OK(cudaMalloc(&d->ws_in, std::max<size_t>(in_bytes, 32)));
        d->ws_in_cap = std::max<size_t>(in_bytes, 32);
    }
    if (d->ws_status_cap < st_bytes) {
        cudaFree(d->ws_status);
        d->ws_status = nullptr;
        d->ws_status_cap = 0;
        CUDA_OK(cudaMalloc(&d->ws_status, st_bytes));
        d->ws_status_cap = st_bytes;
    }
    cudaStream_t* st = d->ws_stream;
    void** d_vals = d->ws_vals;
    u32* d_status = d->ws_status;
    // inputs go up chunk by chunk on the stream that consumes them, so the first launch does not wait
    // for the whole input upload
    int k = 0;
    for (uint64_t t0 = 0; t0 < tiles; t0 += tiles_per_chunk, k ^= 1) {
        uint64_t nt = std::min(tiles_per_chunk, tiles - t0);
        uint64_t i0 = t0 * TILE, ni = std::min<uint64_t>(n_inst - i0, nt * TILE);
        const size_t in_off = (size_t)i0 * sh.n_inputs * 32, in_len = (size_t)ni * sh.n_inputs * 32;
        if (in_len) CUDA_OK(cudaMemcpyAsync((char*)d->ws_in + in_off, (const char*)h_inputs + in_off, in_len, cudaMemcpyHostToDevice, st[k]));
        rc = launch_vm(s, d, st[k], (u32*)d_vals[k], (const u32*)d->ws_in + i0 * sh.n_inputs * 8, d_status + i0, ni);
        if (rc) return rc;
        if (s->export_format == H2E_EXPORT_MONTGOMERY) {
            rc = launch_montgomery(d, st[k], (u32*)d_vals[k], nt * tile_bytes / 32);
            if (rc) return rc;
        }
        CUDA_OK(cudaMemcpyAsync((char*)h_vals + t0 * tile_bytes, d_vals[k], nt * tile_bytes, cudaMemcpyDeviceToHost, st[k]));
        CUDA_OK(cudaMemcpyAsync(h_status + i0, d_status + i0, ni * 4, cudaMemcpyDeviceToHost, st[k]));
    }
    CUDA_OK(cudaStreamSynchronize(st[0]));
    CUDA_OK(cudaStreamSynchronize(st[1]));
    return 0;
}

}  // extern "C"
