"""Summarise `ncu --page source --csv` (per-SASS-instruction stall sampling): python exp/ncu_source_summary.py <csv> [top]
Per kernel: samples by stall reason, by opcode of the stalled instruction, and the top instructions with two lines of context."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
csv.field_size_limit(1 << 30)
kernels, cur, hdr = [], None, None
for row in csv.reader(open(path, newline="")):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1][:40], "rows": []}
        kernels.append(cur)
        hdr = None
        continue
    if row[0] == "Address":
        hdr = row
        continue
    if cur is not None and hdr is not None:
        cur["rows"].append(row)
for k in kernels:
    col = {h: i for i, h in enumerate(hdr)}
    S = col["# Samples"]
    stall_cols = [(h, i) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h]
    rows = k["rows"]
    tot = sum(int(r[S] or 0) for r in rows)
    ex = sum(int(r[col["Instructions Executed"]] or 0) for r in rows)
    print(f"=== {k['name']}: {len(rows)} SASS instructions, {ex} warp-instructions executed, {tot} samples")
    print("  by reason: " + ", ".join(f"{h[6:]} {100 * sum(int(r[i] or 0) for r in rows) / max(tot, 1):.1f}%" for h, i in sorted(stall_cols, key=lambda x: -sum(int(r[x[1]] or 0) for r in rows))[:9]))
    by_op = collections.Counter()
    by_op_long = collections.Counter()
    L = col["stall_long_sb"]
    for r in rows:
        op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
        if op.startswith("@"):
            op = r[col["Source"]].split()[1]
        op = op.split(".")[0]
        by_op[op] += int(r[S] or 0)
        by_op_long[op] += int(r[L] or 0)
    print("  samples by stalled opcode: " + ", ".join(f"{o} {100 * n / max(tot, 1):.1f}% (long_sb {100 * by_op_long[o] / max(tot, 1):.1f}%)" for o, n in by_op.most_common(12)))
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][S] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        best = max(stall_cols, key=lambda x: int(r[x[1]] or 0))
        ctx = " | ".join(rows[j][col["Source"]].strip()[:46] for j in range(max(0, i - 2), i))
        print(f"  #{i:6d} {100 * int(r[S]) / max(tot, 1):5.2f}% {best[0][6:]:10s} {r[col['Source']].strip()[:60]:60s} <- {ctx}")
