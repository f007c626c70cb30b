// Interface between the host library (h2e_lib.cu) and the kernel translation units (vm_kernel.cu,
// compiled once per variant).
#pragma once
#include <cuda_runtime.h>

#include "h2e_program.h"
#include "schedule.h"

namespace h2e {

typedef uint32_t u32;

struct TeamProg {
    const Instr* crit;
    const DepRec* crit_dep;
    const uint32_t* crit_off;
    const Instr* tail;
    const DepRec* tail_dep;
    const uint32_t* tail_off;
    const uint32_t* extra;
    uint32_t n_levels;  // thread mode: number of instructions in `crit`
    uint32_t n_crit;    // critical warps per CTA (the remaining warps of the CTA are tail warps)
    uint32_t G;         // CTAs per tile
    uint32_t twc;       // critical team warps per tile
    uint32_t g_crit;    // 0: every CTA hosts n_crit critical + the rest tail warps; > 0: CTAs [0, g_crit) are all-critical, the others all-tail
};

struct VmLaunch {
    unsigned grid, block;
    cudaStream_t stream;
    TeamProg prog;
    u32* vals;
    const u32* inputs;
    const u32* cpool;
    const u32* tables;
    u32* status;
    u32* progress;
    u32* scratch;        // team mode: [tile][n_scratch][lane][16 words]
    uint32_t n_scratch;
    uint64_t tile_words;  // words of one tile's COMPACT block (probe build: unused)
    uint32_t n_in_cells;
    uint64_t n_inst, n_tiles;
    int mode;
};

// variant with 8 warps per CTA (255 registers per thread) and with 16 warps per CTA (128 registers)
cudaError_t vm_upload_consts_w8(const DeviceConsts* c);
cudaError_t vm_launch_w8(const VmLaunch& L);
cudaError_t vm_montgomery_w8(cudaStream_t stream, unsigned blocks, u32* cells, uint64_t n_cells);
cudaError_t vm_imad_probe(cudaStream_t stream, unsigned blocks, uint64_t* out, uint32_t iters);  // 8 x iters IMAD.WIDE per thread, 256 threads per block
cudaError_t vm_pack(cudaStream_t stream, unsigned blocks, const u32* rec, u32* out, const u32* sel, const u32* uoff, const u32* coff,
                    uint64_t tile_words, uint32_t i0, uint32_t n_i, uint64_t n_tiles, uint64_t out_tile_words);
cudaError_t vm_expand(cudaStream_t stream, unsigned blocks, const u32* rec, u32* out, const u32* coff, uint64_t tile_words, uint64_t s0, uint64_t n_s,
                      uint64_t n_tiles, uint64_t out_tile_words);
cudaError_t vm_scatter(cudaStream_t stream, unsigned blocks, const u32* rec, u32* out, const u32* dst, const u32* ord, const u32* coff,
                       uint64_t tile_words, uint64_t n_slots, uint64_t inst0, uint64_t n_inst, uint64_t cells_per_inst, int mont);
// width-probe build (thread mode only): cells receive their width class instead of their value
cudaError_t vm_upload_consts_wprobe(const DeviceConsts* c);
cudaError_t vm_launch_wprobe(const VmLaunch& L);
cudaError_t vm_upload_consts_w16(const DeviceConsts* c);
cudaError_t vm_launch_w16(const VmLaunch& L);
cudaError_t vm_montgomery_w16(cudaStream_t stream, unsigned blocks, u32* cells, uint64_t n_cells);

}  // namespace h2e
