// TEST INFRASTRUCTURE ONLY. Host emulator of the witness VM: compiles the product's portable
// macro-op code (halo2ecc-s_b200/csrc/vm_ops.cuh) as plain C++ and runs a program instance by
// instance on the CPU, so the CPU test-suite can compare macro-op output with the oracle without
// a GPU. The product library never contains or calls this.
//
// Default build: the VM's own record layout (COMPACT: every cell at its static width class, references
// pre-translated by layout.h exactly as the product does before it uploads a program); the result is expanded
// to plain 32-byte cells for the comparison. -DH2E_WIDTH_PROBE build: plain cells, every store records the
// width class of its call site (checks the static width table).
#include <cstring>
#include <vector>

#include "../../halo2ecc-s_b200/csrc/fieldinfo.h"
#include "../../halo2ecc-s_b200/csrc/layout.h"
#include "../../halo2ecc-s_b200/csrc/vm_ops.cuh"

using namespace h2e;

namespace h2e {
const DeviceConsts* g_host_consts = nullptr;
}

// off[n_slots + 1] / width[n_slots]: the COMPACT layout of the shape (h2e_shape_layout); unused by the probe build.
// vals: [tiles][n_slots][32 lanes][8 words] (probe build: [n_slots][8 words], one instance).
extern "C" int emu_run(const uint8_t* program, uint64_t n_instr, const uint32_t* cpool, const uint32_t* tables, uint64_t n_tables, uint64_t n_slots,
                       uint32_t n_in_cells, uint64_t n_inst, const uint32_t* inputs, const uint32_t* off, const uint8_t* width, uint32_t* vals,
                       uint32_t* status) {
    std::vector<Instr> prog(reinterpret_cast<const Instr*>(program), reinterpret_cast<const Instr*>(program) + n_instr);
    h2e::g_host_consts = &host_consts();
    // scratch entries (team-mode split ops): one 16-word entry per OP_DIV_INV
    uint32_t n_scratch = 0;
    for (uint64_t pc = 0; pc < n_instr; pc++)
        if (prog[pc].op == OP_DIV_INV) {
            const unsigned L = field_info((Field)prog[pc].field).limbs;
            for (unsigned j = 0; j < std::max(1u, prog[pc].flags & 3u); j++) n_scratch = std::max(n_scratch, prog[pc].a[j * (L + 1) + L] + 1);
        }
    std::vector<uint32_t> scratch((size_t)std::max<uint32_t>(n_scratch, 1) * TILE * 16);
#if defined(H2E_WIDTH_PROBE)
    std::vector<uint32_t> tab(tables, tables + n_tables);
    LaneCtx ln;
    ln.vals = vals;
    ln.lane = 0;
    ln.inputs = inputs;
    ln.cpool = cpool;
    ln.tables = tab.data();
    ln.scratch = scratch.data();
    ln.status = 0;
    for (uint64_t pc = 0; pc < n_instr; pc++) exec_instr(ln, prog[pc]);
    status[0] = ln.status;
#else
    translate_program(prog.data(), prog.size(), off, width);
    std::vector<uint32_t> tab(n_tables);
    for (uint64_t i = 0; i < n_tables; i++) tab[i] = tables[i] < n_slots ? slot_ref(tables[i], off, width) : 0;
    const uint64_t tile_words = (uint64_t)off[n_slots] * TILE;
    std::vector<uint32_t> rec(std::max<uint64_t>(tile_words, 1));
    for (uint64_t tile = 0; tile * TILE < n_inst; tile++) {
        std::fill(rec.begin(), rec.end(), 0u);
        const uint64_t lanes = std::min<uint64_t>(TILE, n_inst - tile * TILE);
        for (uint64_t lane = 0; lane < lanes; lane++) {  // padding lanes are not emulated
            const uint64_t inst = tile * TILE + lane;
            LaneCtx ln;
            ln.vals = rec.data();
            ln.lane = (uint32_t)lane;
            ln.inputs = inputs + inst * (uint64_t)n_in_cells * 8;
            ln.cpool = cpool;
            ln.tables = tab.data();
            ln.scratch = scratch.data() + lane * 16;
            ln.status = 0;
            for (uint64_t pc = 0; pc < n_instr; pc++) exec_instr(ln, prog[pc]);
            status[inst] = ln.status;
        }
        for (uint64_t s = 0; s < n_slots; s++)
            for (uint64_t lane = 0; lane < lanes; lane++) {
                uint32_t* q = vals + ((tile * n_slots + s) * TILE + lane) * 8;
                const uint32_t w = width[s];
                for (uint32_t k = 0; k < 8; k++) q[k] = k < w ? rec[(uint64_t)off[s] * TILE + lane * w + k] : 0;
            }
    }
#endif
    return 0;
}
