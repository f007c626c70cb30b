// TEST INFRASTRUCTURE ONLY. Host emulator of the witness VM: compiles the product's portable
// macro-op code (halo2ecc-s_b200/csrc/vm_ops.cuh) as plain C++ and runs a program instance by
// instance on the CPU, so the CPU test-suite can compare macro-op output with the oracle without
// a GPU. The product library never contains or calls this.
#include <cstring>
#include <vector>

#include "../../halo2ecc-s_b200/csrc/fieldinfo.h"
#include "../../halo2ecc-s_b200/csrc/vm_ops.cuh"

using namespace h2e;

namespace h2e {
const DeviceConsts* g_host_consts = nullptr;
}

extern "C" int emu_run(const uint8_t* program, uint64_t n_instr, const uint32_t* cpool, const uint32_t* tables, uint64_t n_slots, uint32_t n_in_cells,
                       uint64_t n_inst, const uint32_t* inputs, uint32_t* vals, uint32_t* status) {
    const Instr* prog = reinterpret_cast<const Instr*>(program);
    h2e::g_host_consts = &host_consts();
    uint64_t padded = (n_inst + TILE - 1) / TILE * TILE;
    (void)padded;
    // scratch entries (team-mode split ops): one 16-word entry per OP_DIV_INV
    uint32_t n_scratch = 0;
    for (uint64_t pc = 0; pc < n_instr; pc++)
        if (prog[pc].op == OP_DIV_INV) n_scratch = std::max(n_scratch, prog[pc].a[13] + 1);
    std::vector<uint32_t> scratch((size_t)std::max<uint32_t>(n_scratch, 1) * TILE * 16);
    for (uint64_t inst = 0; inst < n_inst; inst++) {  // padding lanes are not emulated
        uint64_t tile = inst / TILE, lane = inst % TILE;
        uint64_t in_inst = inst < n_inst ? inst : n_inst - 1;
        LaneCtx ln;
        ln.vals = vals + (tile * n_slots * TILE + lane) * 8;
        ln.inputs = inputs + in_inst * (uint64_t)n_in_cells * 8;
        ln.cpool = cpool;
        ln.tables = tables;
        ln.scratch = scratch.data() + lane * 16;
        ln.status = 0;
        for (uint64_t pc = 0; pc < n_instr; pc++) exec_instr(ln, prog[pc]);
        if (inst < n_inst) status[inst] = ln.status;
    }
    return 0;
}
