"""Sweep the critical / tail CTA split of team mode on one circuit workload (records resident, COMPACT):
python exp/split_sweep.py <configs[k]> [instances] [eighths, e.g. 0,3,4,5,6] [warps 8|16|0]
eighths = critical share of a tile's CTAs in eighths (0 = the library's own choice from its cost model)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as ge

h2e = ge.load_package()
cfg = sys.argv[1]
w = [x for x in bench.CIRCUIT_WORKLOADS if x[5] == cfg][0]
name, kind, params, gen, n_inst = w[:5]
if len(sys.argv) > 2 and int(sys.argv[2]):
    n_inst = int(sys.argv[2])
eighths = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 3, 4, 5, 6]
warps = int(sys.argv[4]) if len(sys.argv) > 4 else 0
t0 = time.time()
shape = h2e.Shape.build(kind, params)
rows = bench._circuit_inputs(gen, n_inst, seed=0)
d_in = torch.from_numpy(h2e.pack_inputs(rows)).cuda()
rec = torch.empty((shape.records_bytes(h2e.REC_COMPACT, n_inst),), dtype=torch.uint8, device="cuda")
st = torch.empty(((n_inst + 31) // 32 * 32,), dtype=torch.int32, device="cuda")
print(f"{name}: {n_inst} instances, setup {time.time() - t0:.1f} s", flush=True)
for e in eighths:
    t1 = time.time()
    shape.set_mode((e << 8) | (warps << 16), 0)
    shape.run_records(d_in, h2e.REC_COMPACT, rec, st)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for r in range(3):
        shape.run_records(d_in, h2e.REC_COMPACT, rec, st)
        ev[r + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[r].elapsed_time(ev[r + 1]) for r in range(3))
    print(f"  critical eighths {e} warps {warps or 'auto'}: {ms[0]:.2f} / {ms[1]:.2f} / {ms[2]:.2f} ms, {n_inst / ms[1] * 1e3:.0f} witnesses/s, "
          f"bad status {int((st[:n_inst] != 0).sum())}, streams built in {time.time() - t1:.1f} s", flush=True)
