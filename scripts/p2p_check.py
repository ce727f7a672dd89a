"""One-shot NVLink all-reduce (ctx_p2p_allreduce) against NCCL: correctness and latency.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/p2p_check.py
Prints one JSON line (rank 0); exits non-zero on any mismatch."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

from catalax_b200.parallel import PeerAllReduce

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {"world": world}
ok = True
for dtype in (torch.float32, torch.float64):
    CH, KR = 16384, 6
    p2p = PeerAllReduce(CH * KR * 8)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    for it in range(25):            # back-to-back calls exercise both slots and the epoch protocol
        n = CH * KR if it % 3 else 1000 + 17 * it
        x = torch.randn(n, generator=g, device=dev, dtype=dtype)
        ref = x.clone()
        dist.all_reduce(ref)
        got = p2p(x.clone())
        # identical summation order on every rank -> bit-identical across ranks; vs NCCL: rounding
        tol = 1e-5 if dtype == torch.float32 else 1e-13
        err = float((got - ref).abs().max() / ref.abs().max())
        allg = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(allg, got)
        same = all(torch.equal(allg[0], a) for a in allg)
        if err > tol or not same:
            ok = False
            print(f"rank {rank} MISMATCH dtype {dtype} it {it}: err {err:.3e} identical_across_ranks {same}", flush=True)
    # latency, eager and from a CUDA graph, next to NCCL
    x = torch.randn(CH * KR, device=dev, dtype=dtype)

    def timed(fn, n=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    key = "f32" if dtype == torch.float32 else "f64"
    res[f"{key}_p2p_us"] = timed(lambda: p2p(x))
    res[f"{key}_nccl_us"] = timed(lambda: dist.all_reduce(x))
    try:
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            for _ in range(3):
                p2p(x)
            side.synchronize(); dist.barrier()
            with torch.cuda.graph(gr, stream=side):
                p2p(x)
        torch.cuda.current_stream().wait_stream(side)
        y = torch.ones(CH * KR, device=dev, dtype=dtype) * (rank + 1)
        x.copy_(y); gr.replay(); torch.cuda.synchronize()
        if not torch.allclose(x, torch.full_like(x, world * (world + 1) / 2)):
            ok = False; print(f"rank {rank} graph replay mismatch", flush=True)
        res[f"{key}_p2p_graph_us"] = timed(lambda: gr.replay())
    except Exception as exc:
        res[f"{key}_graph_error"] = f"{type(exc).__name__}: {exc}"[:200]
    p2p.close()
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
res["ok"] = int(flag.item()) == 0
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
sys.exit(0 if res["ok"] else 1)
