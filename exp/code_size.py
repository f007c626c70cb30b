"""SASS size of every out-of-line device function of the VM kernel (the instruction-cache footprint of the macro-ops).
usage: python exp/code_size.py halo2ecc-s_b200/build/vm_kernel_w8.o   (needs cuobjdump, nvdisasm, c++filt)"""
import os, re, subprocess, sys, tempfile
obj = os.path.abspath(sys.argv[1])
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-c", os.path.join(d, cubin)], check=True, capture_output=True, text=True).stdout
fn, sizes, order = None, {}, []
for line in dis.splitlines():
    m = re.match(r"\s*\.type\s+(\S+),@function", line)
    if m:
        fn = m.group(1)
        sizes[fn] = 0
        order.append(fn)
        continue
    if fn and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
        sizes[fn] += 1
def pretty(sym):
    inner = sym.split("$")[-1] if sym.startswith("$") else sym
    try:
        out = subprocess.run(["c++filt", inner], capture_output=True, text=True).stdout.strip()
    except Exception:
        out = inner
    out = re.sub(r"\(h2e::LaneCtx&, h2e::Instr const&\)", "", out)
    out = re.sub(r"^void ", "", out)
    return out[:90]
total = sum(sizes.values())
print(f"| function | SASS instructions | KB |\n|---|---:|---:|")
for f, n in sorted(sizes.items(), key=lambda kv: -kv[1]):
    if n >= 200:
        print(f"| `{pretty(f)}` | {n} | {n * 16 / 1024:.1f} |")
print(f"| **total ({len(sizes)} functions)** | {total} | {total * 16 / 1024:.0f} |")
