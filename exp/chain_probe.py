import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, random
import __graft_entry__ as ge
h2e = ge.load_package()
K = 200
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
rng = random.Random(1)
sb = h2e.ScriptBuilder()
a = sb.assign_w(0); b = sb.assign_w(1)
x = a
for i in range(K):
    x = sb.int_add(x, b)
shape = h2e.Shape.from_script(0, sb.words)
for mode, C in [(1, 0), (2, 1), (2, 4)]:
    shape.set_mode(mode, C)
    packed = h2e.pack_inputs([[rng.randrange(P), rng.randrange(1, P)] for _ in range(32)])
    d_in = torch.from_numpy(packed).cuda()
    vals, st = shape.run(d_in); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); shape.run(d_in, vals, st); e1.record(); torch.cuda.synchronize()
    print(f'int_add chain mode={mode} C={C}: {e0.elapsed_time(e1)*1e3/K:.2f} us/op, instrs {shape.n_instr}', flush=True)
