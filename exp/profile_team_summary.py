"""Summarise the printf output of the -DH2E_PROFILE build of the VM kernel (team mode, tile 0, last pass in the log).
usage: python exp/profile_team_summary.py LOG [clock_ghz]"""
import sys, re, collections
OPS = {}
for line in open('halo2ecc-s_b200/csrc/h2e_program.h'):
    pass
def op_names():
    import re as _re
    src = open('halo2ecc-s_b200/csrc/h2e_program.h').read()
    body = src[src.index('enum Op'):]
    body = body[body.index('{') + 1:body.index('};')]
    names, k = {}, 0
    for ln in body.splitlines():
        ln = ln.split('//')[0].strip()
        for tok in ln.split(','):
            tok = tok.strip()
            if not tok:
                continue
            if '=' in tok:
                nm, v = [x.strip() for x in tok.split('=')]
                k = int(v, 0)
            else:
                nm = tok
            names[k] = nm
            k += 1
    return names
names = op_names()
ghz = float(sys.argv[2]) if len(sys.argv) > 2 else 1.965
W, O = [], []
for line in open(sys.argv[1]):
    if line.startswith('W '):
        t = line.split()
        W.append((t[1], int(t[3]), int(t[5]), int(t[7]), int(t[9]), int(t[11]), int(t[13]), int(t[15])))
    elif line.startswith('O '):
        t = line.split()
        O.append((t[1], int(t[3]), int(t[5]), int(t[7]), int(t[9]), int(t[11])))
# keep the last pass only: warps are unique by (role, cta, warp)
lastW = {}
for w in W:
    lastW[(w[0], w[1], w[2])] = w
lastO = {}
for o in O:
    lastO[(o[0], o[1], o[2])] = o
for role in ('crit', 'tail'):
    ws = [w for w in lastW.values() if w[0] == role]
    if not ws:
        continue
    n = len(ws)
    tot = max(w[7] for w in ws)
    print(f'{role}: {n} warps, span {tot / ghz / 1e6:.1f} ms; mean per warp: instr {sum(w[3] for w in ws) / n:.0f}, '
          f'wait {sum(w[4] for w in ws) / n / tot:.2f}, exec {sum(w[5] for w in ws) / n / tot:.2f}, publish {sum(w[6] for w in ws) / n / tot:.2f} of the span')
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for o in lastO.values():
        if o[0] == role:
            a = agg[o[2]]
            a[0] += o[3]; a[1] += o[4]; a[2] += o[5]
    tot_exec = sum(a[1] for a in agg.values()) or 1
    print(f'  {"op":22s} {"count":>9s} {"exec cyc/op":>12s} {"wait cyc/op":>12s} {"share of exec":>14s}')
    for op, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'  {names.get(op, op):22s} {a[0]:9d} {a[1] / a[0]:12.0f} {a[2] / a[0]:12.0f} {a[1] / tot_exec:14.3f}')
