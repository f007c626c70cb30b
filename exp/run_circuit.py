"""Run one circuit workload resident on the GPU (for ncu captures): python exp/run_circuit.py <configs[k]> [instances] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as ge

h2e = ge.load_package()
cfg = sys.argv[1]
w = [x for x in bench.CIRCUIT_WORKLOADS if x[5] == cfg][0]
name, kind, params, gen, n_inst = w[:5]
if len(sys.argv) > 2:
    n_inst = int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
shape = h2e.Shape.build(kind, params)
rows = bench._circuit_inputs(gen, n_inst, seed=0)
d_in = torch.from_numpy(h2e.pack_inputs(rows)).cuda()
rec, st = shape.run_records(d_in, h2e.REC_COMPACT)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for r in range(reps):
    shape.run_records(d_in, h2e.REC_COMPACT, rec, st)
    ev[r + 1].record()
torch.cuda.synchronize()
print(name, n_inst, "instances:", [round(ev[r].elapsed_time(ev[r + 1]), 2) for r in range(reps)], "ms; nonzero status:", int((st[:n_inst] != 0).sum()))
