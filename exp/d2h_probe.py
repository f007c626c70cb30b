"""Pinned-host D2H rate vs. size of the pinned working set and allocation method (the write-out roofline).
python exp/d2h_probe.py            (one GPU)   |   torchrun --nproc-per-node N exp/d2h_probe.py   (all ranks at once)
Methods: 'hostalloc' = cudaHostAlloc (torch pin_memory); 'thp+register' = anonymous mmap + MADV_HUGEPAGE, first touch,
cudaHostRegister. For each working-set size S the probe streams 256 MB pieces from one device buffer across the whole
S bytes (so every page of the set is a DMA target once per pass) and reports GB/s per GPU (max time over ranks)."""
import ctypes
import json
import mmap
import os
import sys
import time

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def allmax(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


PIECE = 256 << 20
src = torch.zeros((PIECE,), dtype=torch.uint8, device=dev)
libc = ctypes.CDLL(None)
cudart = torch.cuda.cudart()


def alloc(method, nbytes):
    if method == "hostalloc":
        return torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True), None
    mm = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    buf = (ctypes.c_char * nbytes).from_buffer(mm)
    addr = ctypes.addressof(buf)
    libc.madvise(ctypes.c_void_p(addr), ctypes.c_size_t(nbytes), 14)  # MADV_HUGEPAGE
    t = torch.frombuffer(mm, dtype=torch.uint8)
    t[:: 4096] = 1  # first touch
    rc = cudart.cudaHostRegister(addr, nbytes, 0)
    assert int(rc) == 0, f"cudaHostRegister failed: {rc}"
    return t, (mm, buf, addr)


def free(method, t, h):
    if h is not None:
        cudart.cudaHostUnregister(h[2])
    del t
    try:
        torch._C._host_emptyCache()
    except Exception:
        pass


out = []
sizes_gb = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "4", "16"])]
for method in ("hostalloc", "thp+register"):
    for gb in sizes_gb:
        n = gb * (1 << 30) // PIECE
        try:
            t, h = alloc(method, n * PIECE)
        except Exception as e:
            out.append({"method": method, "gb": gb, "error": str(e)[:100]})
            continue
        views = [t[i * PIECE:(i + 1) * PIECE] for i in range(n)]
        best = None
        for rep in range(3):
            barrier()
            w0 = time.perf_counter()
            for v in views:
                v.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            sec = allmax(time.perf_counter() - w0)
            if rep:
                best = sec if best is None else min(best, sec)
        out.append({"method": method, "working_set_gb": gb, "gbs_per_gpu": round(n * PIECE / best / 1e9, 1), "ranks": world})
        del views
        free(method, t, h)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
