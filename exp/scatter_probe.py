"""Time h2e_records_scatter (device-side prover hand-off) on one circuit workload: python exp/scatter_probe.py <configs[k]> [instances] [encoding 0|1] [order 1|2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as ge

h2e = ge.load_package()
cfg = sys.argv[1]
w = [x for x in bench.CIRCUIT_WORKLOADS if x[5] == cfg][0]
name, kind, params, gen, n_inst = w[:5]
n_inst = int(sys.argv[2]) if len(sys.argv) > 2 else 32
enc = int(sys.argv[3]) if len(sys.argv) > 3 else 1
order = int(sys.argv[4]) if len(sys.argv) > 4 else 1
shape = h2e.Shape.build(kind, params)
rows = bench._circuit_inputs(gen, n_inst, seed=0)
d_in = torch.from_numpy(h2e.pack_inputs(rows)).cuda()
rec, st = shape.run_records(d_in, h2e.REC_COMPACT)
dense = torch.zeros((n_inst, shape.dense_cells(), 32), dtype=torch.uint8, device="cuda")
shape.records_scatter(rec, n_inst, out=dense, order=order, encoding=enc)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
for r in range(3):
    shape.records_scatter(rec, n_inst, out=dense, order=order, encoding=enc)
    ev[r + 1].record()
torch.cuda.synchronize()
ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(3)]
gb = (shape.records_bytes(h2e.REC_COMPACT, n_inst) + n_inst * shape.n_slots * 32) / 1e9
print(f"{name}: {n_inst} instances, encoding {enc}, order {order}, CTAs/SM {os.environ.get('H2E_SCATTER_CTAS', '6')}: "
      f"{[round(m, 3) for m in ms]} ms, {gb / (min(ms) * 1e-3):.0f} GB/s read + write ({gb:.2f} GB)")
