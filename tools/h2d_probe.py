"""Host->device copy rate of pinned buffers, default vs write-combined allocation, one process per GPU
(torchrun-compatible: every rank copies to its own GPU at the same time; rank 0 prints the per-rank and total rate).
Usage: [torchrun --nproc-per-node N] python tools/h2d_probe.py"""
import ctypes as C
import os
import time

import numpy as np
import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
rt = C.CDLL("libcudart.so.12")
NBYTES = 256 << 20


def host_alloc(flags):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(NBYTES), C.c_uint(flags))
    assert rc == 0, rc
    arr = np.ctypeslib.as_array((C.c_uint8 * NBYTES).from_address(p.value))
    arr[:] = 7
    return torch.from_numpy(arr), p


def rate(t):
    d = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
    for _ in range(2):
        d.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        d.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    return NBYTES * n / (time.perf_counter() - t0) / 1e9


for name, flags in (("default", 0), ("write-combined", 4), ("torch pin_memory", None)):
    if flags is None:
        t = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
        t.fill_(7)
    else:
        t, keep = host_alloc(flags)
    r = rate(t)
    if world > 1:
        v = torch.tensor([r], device=dev, dtype=torch.float64)
        g = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(g, v)
        if rank == 0:
            rs = [float(x) for x in g]
            print("%-18s pinned=%s  per rank %s  total %.1f GB/s" % (name, t.is_pinned(), ["%.1f" % x for x in rs], sum(rs)), flush=True)
    else:
        print("%-18s pinned=%s  %.1f GB/s" % (name, t.is_pinned(), r), flush=True)
if world > 1:
    dist.destroy_process_group()
