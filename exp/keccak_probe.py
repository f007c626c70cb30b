"""Time the keccak hash workload resident on the GPU: python exp/keccak_probe.py [instances,...] [modes e.g. 0,1,2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge

h2e = ge.load_package()
ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1184, 8192]
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
sb = h2e.ScriptBuilder()
sb.keccak_hash([sb.assign(i) for i in range(4)])
shape = h2e.Shape.from_script(0, sb.words)
rng = np.random.default_rng(1)
print("keccak hash of 4 scalars:", shape.n_instr, "macro-ops,", shape.n_slots, "cells,", shape.records_bytes(h2e.REC_COMPACT, 32) // 32, "record bytes per instance")
for n in ns:
    rows = [[int.from_bytes(rng.bytes(31), "little") for _ in range(4)] for _ in range(n)]
    d_in = torch.from_numpy(h2e.pack_inputs(rows)).cuda()
    rec = torch.empty((shape.records_bytes(h2e.REC_COMPACT, n),), dtype=torch.uint8, device="cuda")
    st = torch.empty(((n + 31) // 32 * 32,), dtype=torch.int32, device="cuda")
    for mode in modes:
        shape.set_mode(mode, 0)
        shape.run_records(d_in, h2e.REC_COMPACT, rec, st)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for r in range(3):
            shape.run_records(d_in, h2e.REC_COMPACT, rec, st)
            ev[r + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[r].elapsed_time(ev[r + 1]) for r in range(3))
        print(f"  n={n} mode={mode}: {ms[0]:.3f} / {ms[1]:.3f} / {ms[2]:.3f} ms, {n / ms[1] * 1e3:.0f} hashes/s, {rec.numel() / ms[1] / 1e6:.0f} GB/s of records, bad status {int((st[:n] != 0).sum())}")
