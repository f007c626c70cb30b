import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np, random
import __graft_entry__ as ge
h2e = ge.load_package()
field = int(sys.argv[1]) if len(sys.argv) > 1 else 0
K = 200
P = {0: 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47, 1: 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB}[field]
rng = random.Random(1)
def chain(kind):
    sb = h2e.ScriptBuilder()
    a = sb.assign_w(0); b = sb.assign_w(1)
    x = a
    for i in range(K):
        if kind == 'int_mul': x = sb.int_mul(x, b)
        elif kind == 'int_add': x = sb.int_add(x, b)
        elif kind == 'int_sub': x = sb.int_sub(x, b)
        elif kind == 'int_div': x = sb.int_div(x, b)[1]
        elif kind == 'add+reduce': x = sb.reduce(sb.int_add(x, b))
        elif kind == 'is_int_zero': c = sb.is_int_zero(x); x = sb.bisec_int(c, b, x); x = sb.int_mul(x, b)
    return sb
for kind in ['int_add', 'int_mul', 'add+reduce', 'int_div', 'is_int_zero']:
    sb = chain(kind)
    shape = h2e.Shape.from_script(field, sb.words)
    for n, mode in [(32, 1), (32, 2)]:
        shape.set_mode(mode, 1 if mode == 2 else 0)
        packed = h2e.pack_inputs([[rng.randrange(P), rng.randrange(1, P)] for _ in range(n)])
        d_in = torch.from_numpy(packed).cuda()
        vals, st = shape.run(d_in); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); shape.run(d_in, vals, st); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f'field {field} {kind:12s} mode={mode}: {ms*1e3/K:8.1f} us per step ({shape.n_instr/K:.1f} instrs/step, {shape.n_slots/K:.0f} cells/step), status {int(st[:n].abs().max())}')
