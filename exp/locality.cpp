#include <cstdio>
#include "../halo2ecc-s_b200/csrc/circuits.h"
#include "../halo2ecc-s_b200/csrc/schedule.h"
using namespace h2e;
int main(int argc, char** argv) {
    int kind = argc > 1 ? atoi(argv[1]) : 2;
    uint64_t params[1] = {argc > 2 ? (uint64_t)atoll(argv[2]) : 0};
    Context ctx;
    build_circuit(ctx, kind, params, 1);
    Schedule sc = levelise(ctx.shape);
    for (uint32_t G : {5u, 9u, 148u}) {
        uint32_t gc = std::max(1u, (uint32_t)(G * 0.56 + 0.5));
        if (gc >= G) gc = G - 1;
        TeamLayout lay{gc, 8, G - gc, 8, gc};
        TeamStreams ts = build_team_streams(sc, lay);
        printf("kind %d G %u: crit warps %u, preds %lu, same warp %.1f%%, same warp & within last 4 results %.1f%%, est %.1f Mcycles\n", kind, G, ts.twc,
               (unsigned long)ts.stat_preds, 100.0 * ts.stat_same_warp / ts.stat_preds, 100.0 * ts.stat_same_warp_recent / ts.stat_preds, ts.est_cycles / 1e6);
    }
}
