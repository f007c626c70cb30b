"""A/B: one resident pass of the 1000-point MSM on distinct per-instance inputs (as bench.py) vs one instance replicated."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import __graft_entry__ as ge
h2e = ge.load_package()
import bench
n_pts, n = int(sys.argv[1]), int(sys.argv[2])
shape = h2e.Shape.build(0, [n_pts])
rows = bench._circuit_inputs(f"msm:{n_pts}", n, seed=0)
for name, rr in (("distinct", rows), ("replicated", [rows[0]] * n), ("distinct", rows)):
    d_in = torch.from_numpy(h2e.pack_inputs(rr)).cuda()
    tiles = (n + 31) // 32
    vals = torch.empty((tiles, shape.n_slots, 32, 32), dtype=torch.uint8, device='cuda')
    st = torch.empty((tiles * 32,), dtype=torch.int32, device='cuda')
    shape.run(d_in, vals, st); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): shape.run(d_in, vals, st)
    e1.record(); torch.cuda.synchronize()
    print(f'msm n_pts={n_pts} n={n} {name}: {e0.elapsed_time(e1) / 3:.1f} ms, status max {int(st[:n].abs().max())}', flush=True)
    del vals
