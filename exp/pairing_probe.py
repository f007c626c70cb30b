import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np
import __graft_entry__ as ge
h2e = ge.load_package()
import circuits_util as cu
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ns = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [32, 128, 512]
t = time.time()
shape = h2e.Shape.build(kind, [])
print('shape build s', round(time.time() - t, 2), 'slots', shape.n_slots, 'instr', shape.n_instr, 'bytes/inst', shape.n_slots * 32)
base = cu.bn_check_pairing_inputs(1000003, 2000003) if kind == 2 else cu.bls_check_pairing_inputs(424242, 171717, 99999999999)
modes = [(1,0),(2,1),(2,2),(2,4),(2,8)] if len(sys.argv) <= 3 else [tuple(int(y) for y in x.split(':')) for x in sys.argv[3].split(',')]
for n in ns:
  for mode, C in modes:
    shape.set_mode(mode, C)
    packed = h2e.pack_inputs([base] * n)
    d_in = torch.from_numpy(packed).cuda()
    tiles = (n + 31) // 32
    vals = torch.empty((tiles, shape.n_slots, 32, 32), dtype=torch.uint8, device='cuda')
    st = torch.empty((tiles * 32,), dtype=torch.int32, device='cuda')
    shape.run(d_in, vals, st)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): shape.run(d_in, vals, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f'n={n} mode={mode} C={C}: {ms:.1f} ms, {n / ms * 1e3:.1f} inst/s, {n * shape.n_slots * 32 / ms / 1e6:.1f} GB/s, status max {int(st[:n].abs().max())}')
    del vals
