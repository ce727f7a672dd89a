"""Host<->device copy ceiling of this box (pinned memory, one cudaMemcpyAsync per chunk).

The e2e leg of bench.py moves [B,T,S] = 12.6 GB device->host per step; this measures what the
box can do for that transfer alone, per GPU, with 1..N GPUs copying at the same time
(`torchrun --nproc-per-node N scripts/pcie_ceiling.py`), so that e2e can be quoted as a fraction
of it.  Prints one JSON line (rank 0).
"""
import json
import os
import sys
import time

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

GB = float(os.environ.get("PCIE_GB", "4"))
n = int(GB * 1e9) // 4
dev = torch.empty(n, dtype=torch.float32, device="cuda")
host = torch.empty(n, dtype=torch.float32).pin_memory()
res = {"world": world, "bytes": 4 * n}


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return sec


def chunked(dst, src, chunk):
    for o in range(0, n, chunk):
        dst[o:o + chunk].copy_(src[o:o + chunk], non_blocking=True)


for name, dst, src in (("d2h", host, dev), ("h2d", dev, host)):
    for chunk_mb in (0, 16, 64, 256):
        chunk = n if chunk_mb == 0 else chunk_mb * (1 << 20) // 4
        sec = timed(lambda: chunked(dst, src, chunk))
        res[f"{name}_chunk{chunk_mb or 'all'}MB_GBps_per_gpu"] = 4 * n / sec / 1e9
# both directions at once on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
host2 = torch.empty(n // 4, dtype=torch.float32).pin_memory()
dev2 = torch.empty(n // 4, dtype=torch.float32, device="cuda")


def both():
    with torch.cuda.stream(s1):
        host.copy_(dev, non_blocking=True)
    with torch.cuda.stream(s2):
        dev2.copy_(host2, non_blocking=True)


sec = timed(both)
res["d2h_with_concurrent_h2d_GBps_per_gpu"] = 4 * n / sec / 1e9
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
