// C-ABI implementation (include/h2ecc_b200.h) + the CUDA witness-VM kernel for sm_100a.
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/h2ecc_b200.h"
#include "circuits.h"
#include "schedule.h"
#include "script_builder.h"
#include "vm_ops.cuh"

using namespace h2e;

namespace h2e {
__constant__ DeviceConsts g_consts;
}

// One thread = one circuit instance; a warp = 32 consecutive instances = one value tile, so every
// cell store of a warp is one contiguous 1 KiB run. The program is uniform across the grid.
#ifndef H2E_BLOCK
#define H2E_BLOCK 128
#endif
#ifndef H2E_TEAM_WARPS
#define H2E_TEAM_WARPS 8
#endif

// The witness VM kernel. lane = instance within a 32-instance tile, so every cell store of a warp
// is one contiguous 1 KiB run (one 256-bit store per lane).
//
//  * thread mode (team == 0): many instances, short program (e.g. 2^20 int_mul blocks). One warp owns
//    one tile and walks the whole program; level_start = {0, n_instr}.
//  * team mode (team == 1): few instances, long program (a pairing check is ~175k macro-ops and only
//    a few hundred instances fit in HBM). The program is levelised on the host (schedule.h); a
//    thread-block cluster owns one tile, every warp of the cluster takes the instructions of the
//    current dependency level round-robin, and a cluster barrier (release/acquire at cluster scope)
//    separates levels.
__global__ void __launch_bounds__(H2E_TEAM_WARPS * 32, 1)
    h2e_vm_kernel(const Instr* __restrict__ prog, const uint32_t* __restrict__ level_start, const uint32_t* __restrict__ level_mid,
                  uint32_t n_levels, u32* __restrict__ vals,
                  const u32* __restrict__ inputs, const u32* __restrict__ cpool, const u32* __restrict__ tables, u32* __restrict__ status,
                  uint64_t n_slots, uint32_t n_in_cells, uint64_t n_tiles, uint64_t n_inst, int team) {
    const unsigned lane = threadIdx.x % TILE, warp = threadIdx.x / TILE;
    unsigned C = 1, tw = 0, TW = 1;
    uint64_t tile;
    if (team) {
        unsigned rank;
        asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(C));
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
        tile = blockIdx.x / C;
        tw = warp * C + rank;  // consecutive (heaviest-first) instructions of a level go to different SMs
        TW = C * (blockDim.x / TILE);
    } else {
        tile = (uint64_t)blockIdx.x * (blockDim.x / TILE) + warp;
        if (tile >= n_tiles) return;
    }
    uint64_t inst = tile * TILE + lane;
    uint64_t in_inst = inst < n_inst ? inst : (n_inst - 1);  // padding lanes recompute the last instance
    LaneCtx ln;
    ln.vals = vals + (tile * n_slots * TILE + lane) * 8;
    ln.inputs = inputs + in_inst * (uint64_t)n_in_cells * 8;
    ln.cpool = cpool;
    ln.tables = tables;
    ln.status = 0;
    auto fetch = [&](Instr& dst_in, uint32_t pc) {
        const uint4* src = reinterpret_cast<const uint4*>(prog + pc);
        uint4* dst = reinterpret_cast<uint4*>(&dst_in);
        dst[0] = __ldg(src + 0);
        dst[1] = __ldg(src + 1);
        dst[2] = __ldg(src + 2);
        dst[3] = __ldg(src + 3);
    };
    if (!team) {
        uint32_t end = __ldg(level_start + 1);
        for (uint32_t pc = __ldg(level_start); pc < end; pc++) {
            Instr in;
            fetch(in, pc);
            exec_instr(ln, in);
        }
        status[inst] = ln.status;
        return;
    }
    // team mode. Per level: critical ops -> barrier.arrive (release: their cells become visible to the
    // cluster) -> deferred ops (the TAIL halves of int_mul; nothing reads their cells, so they overlap
    // the barrier and drain their stores during the following levels) -> barrier.wait (acquire).
    // Critical ops are dealt to team warps 0,1,2,.. and deferred ops to TW-1,TW-2,.. so that a warp
    // rarely has both. The next level's first instruction is fetched before the wait.
    uint32_t begin = __ldg(level_start);
    Instr nxt;
    bool have_nxt = false;
    for (uint32_t l = 0; l < n_levels; l++) {
        const uint32_t mid = __ldg(level_mid + l), end = __ldg(level_start + l + 1);
        for (uint32_t pc = begin + tw; pc < mid; pc += TW) {
            Instr in;
            if (have_nxt && pc == begin + tw)
                in = nxt;
            else
                fetch(in, pc);
            if (team != 3) exec_instr(ln, in);
            else ln.status |= (in.op == 0xffff);
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        for (uint32_t pc = mid + (TW - 1 - tw); pc < end; pc += TW) {
            Instr in;
            fetch(in, pc);
            if (team != 3) exec_instr(ln, in);
            else ln.status |= (in.op == 0xffff);
        }
        have_nxt = false;
        if (l + 1 < n_levels) {
            uint32_t nmid = __ldg(level_mid + l + 1);
            if (end + tw < nmid) {
                fetch(nxt, end + tw);
                have_nxt = true;
            }
        }
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        begin = end;
    }
    (void)C;
    if (ln.status) atomicOr(&status[inst], ln.status);
}

// -----------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches(0);

struct DeviceState {
    Instr* d_prog = nullptr;
    u32* d_cpool = nullptr;
    u32* d_tables = nullptr;
    Instr* d_sched_prog = nullptr;
    uint32_t* d_level_start = nullptr;
    uint32_t* d_level_mid = nullptr;
    uint32_t* d_flat_levels = nullptr;  // {0, n_instr}: thread mode
    uint32_t n_levels = 0;
    int sm_count = 0;
    bool consts_uploaded = false;
};

struct h2e_shape {
    Context ctx;
    Schedule sched;
    bool sched_ready = false;
    int force_mode = 0;  // 0 auto, 1 thread-per-instance, 2 team
    int force_cluster = 0;
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
        CUDA_OK(cudaMemcpyToSymbol(g_consts, &host_consts(), sizeof(DeviceConsts)));
        CUDA_OK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device));
        uint32_t flat[2] = {0, (uint32_t)sh.program.size()};
        CUDA_OK(cudaMalloc(&d.d_flat_levels, 8));
        CUDA_OK(cudaMemcpy(d.d_flat_levels, flat, 8, cudaMemcpyHostToDevice));
    }
    *out = &d;
    return 0;
}

static uint64_t pad_tiles(uint64_t n) { return (n + TILE - 1) / TILE * TILE; }

// Upload the levelised program on first use.
static int ensure_schedule(h2e_shape* s, DeviceState* d) {
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->sched_ready) {
        s->sched = levelise(s->ctx.shape);
        s->sched_ready = true;
    }
    if (!d->d_sched_prog) {
        const Schedule& sc = s->sched;
        CUDA_OK(cudaMalloc(&d->d_sched_prog, std::max<size_t>(sc.program.size(), 1) * sizeof(Instr)));
        CUDA_OK(cudaMalloc(&d->d_level_start, sc.level_start.size() * 4));
        CUDA_OK(cudaMemcpy(d->d_sched_prog, sc.program.data(), sc.program.size() * sizeof(Instr), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(d->d_level_start, sc.level_start.data(), sc.level_start.size() * 4, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&d->d_level_mid, std::max<size_t>(sc.level_mid.size(), 1) * 4));
        CUDA_OK(cudaMemcpy(d->d_level_mid, sc.level_mid.data(), sc.level_mid.size() * 4, cudaMemcpyHostToDevice));
        d->n_levels = (uint32_t)sc.level_start.size() - 1;
    }
    return 0;
}

// Launch one pass of the VM over `tiles` tiles. Chooses thread-per-instance (many instances, short
// program) or team mode (few instances, long program).
static int launch_vm(h2e_shape* s, DeviceState* d, cudaStream_t stream, u32* d_vals, const u32* d_inputs, u32* d_status, uint64_t n_inst) {
    const Shape& sh = s->ctx.shape;
    uint64_t padded = pad_tiles(n_inst), tiles = padded / TILE;
    int sms = d->sm_count > 0 ? d->sm_count : 148;
    bool team = sh.program.size() >= 64 && tiles * 2 <= (uint64_t)sms * 4;
    if (s->force_mode == 1) team = false;
    if (s->force_mode >= 2) team = true;
    if (!team) {
        const int block = H2E_BLOCK;
        uint64_t grid = (padded + block - 1) / block;
        h2e_vm_kernel<<<(unsigned)grid, block, 0, stream>>>(d->d_prog, d->d_flat_levels, d->d_flat_levels, 1u, d_vals, d_inputs, d->d_cpool, d->d_tables, d_status,
                                                            sh.slot_cell.size(), sh.n_inputs, tiles, n_inst, 0);
        g_launches++;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    int rc = ensure_schedule(s, d);
    if (rc) return rc;
    // cluster size: as many CTAs per tile as keep the whole GPU busy, capped by the portable limit
    unsigned C = 1;
    while (C < 8 && tiles * (C * 2) * 5 <= (uint64_t)sms * 4) C *= 2;  // keep all clusters co-resident (<= 80% of the SMs)
    if (s->force_cluster > 0) C = (unsigned)s->force_cluster;
    CUDA_OK(cudaMemsetAsync(d_status, 0, padded * 4, stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(tiles * C), 1, 1);
    cfg.blockDim = dim3(H2E_TEAM_WARPS * 32, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_OK(cudaLaunchKernelEx(&cfg, h2e_vm_kernel, (const Instr*)d->d_sched_prog, (const uint32_t*)d->d_level_start,
                               (const uint32_t*)d->d_level_mid, d->n_levels, d_vals,
                               d_inputs, (const u32*)d->d_cpool, (const u32*)d->d_tables, d_status, (uint64_t)sh.slot_cell.size(),
                               (uint32_t)sh.n_inputs, tiles, n_inst, s->force_mode == 3 ? 3 : 1));
    g_launches++;
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
            cudaFree(kv.second.d_sched_prog);
            cudaFree(kv.second.d_level_start);
            cudaFree(kv.second.d_level_mid);
            cudaFree(kv.second.d_flat_levels);
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

int h2e_shape_set_mode(h2e_shape* s, int mode, int cluster_size) {
    s->force_mode = mode;
    s->force_cluster = cluster_size;
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
    cudaStream_t st[2];
    void* d_vals[2] = {nullptr, nullptr};
    void* d_in = nullptr;
    u32* d_status = nullptr;
    CUDA_OK(cudaStreamCreate(&st[0]));
    CUDA_OK(cudaStreamCreate(&st[1]));
    CUDA_OK(cudaMalloc(&d_vals[0], tiles_per_chunk * tile_bytes));
    CUDA_OK(cudaMalloc(&d_vals[1], tiles_per_chunk * tile_bytes));
    size_t in_bytes = h2e_inputs_bytes(s, n_inst);
    CUDA_OK(cudaMalloc(&d_in, std::max<size_t>(in_bytes, 32)));
    CUDA_OK(cudaMalloc(&d_status, pad_tiles(n_inst) * 4));
    if (in_bytes) CUDA_OK(cudaMemcpy(d_in, h_inputs, in_bytes, cudaMemcpyHostToDevice));
    int k = 0;
    for (uint64_t t0 = 0; t0 < tiles; t0 += tiles_per_chunk, k ^= 1) {
        uint64_t nt = std::min(tiles_per_chunk, tiles - t0);
        uint64_t i0 = t0 * TILE, ni = std::min<uint64_t>(n_inst - i0, nt * TILE);
        rc = launch_vm(s, d, st[k], (u32*)d_vals[k], (const u32*)d_in + i0 * sh.n_inputs * 8, d_status + i0, ni);
        if (rc) return rc;
        CUDA_OK(cudaMemcpyAsync((char*)h_vals + t0 * tile_bytes, d_vals[k], nt * tile_bytes, cudaMemcpyDeviceToHost, st[k]));
    }
    CUDA_OK(cudaStreamSynchronize(st[0]));
    CUDA_OK(cudaStreamSynchronize(st[1]));
    CUDA_OK(cudaMemcpy(h_status, d_status, n_inst * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_vals[0]);
    cudaFree(d_vals[1]);
    cudaFree(d_in);
    cudaFree(d_status);
    cudaStreamDestroy(st[0]);
    cudaStreamDestroy(st[1]);
    return 0;
}

}  // extern "C"
