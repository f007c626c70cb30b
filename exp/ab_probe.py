# A/B probe: python exp/ab_probe.py kind n_list  (env toggles are read at schedule build, so one process per setting)
import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np
import __graft_entry__ as ge
h2e = ge.load_package()
import bench
kind = sys.argv[1]; ns = [int(x) for x in sys.argv[2].split(',')]
modes = [tuple(int(y) for y in x.split(':')) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [(0, 0)]
spec = {'bn': (2, [], 'pairing_bn256'), 'bls': (3, [], 'pairing_bls12_381'), 'msm': (0, [1000], 'msm:1000')}[kind]
shape = h2e.Shape.build(spec[0], spec[1])
for n in ns:
    rows = bench._circuit_inputs(spec[2], n, 0)
    d_in = torch.from_numpy(h2e.pack_inputs(rows)).cuda()
    tiles = (n + 31) // 32
    vals = torch.empty((tiles, shape.n_slots, 32, 32), dtype=torch.uint8, device='cuda')
    st = torch.empty((tiles * 32,), dtype=torch.int32, device='cuda')
    for mode, C in modes:
        shape.set_mode(mode, C)
        shape.run(d_in, vals, st); torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); shape.run(d_in, vals, st); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f'{kind} n={n} mode={mode}:{C} env={os.environ.get("H2E_MIXED","-")}: ms {min(ts):.1f} / {sorted(ts)[len(ts)//2]:.1f} / {max(ts):.1f}  -> {n/min(ts)*1e3:.0f} inst/s  status {int(st[:n].abs().max())}', flush=True)
    del vals
