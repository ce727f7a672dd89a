"""Write-only HBM bandwidth on this GPU (memset / fill of a 12.6 GB buffer) next to the copy
figure of MEASURED_PEAKS.json: the T=1000 solve writes 12.58 GB and reads 0.02 GB."""
import json
import torch

n = 12_582_912_000 // 4
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n // 2, dtype=torch.float32, device="cuda")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


out = {}
ms = timed(lambda: x.zero_()); out["memset_12.6GB"] = {"ms": ms, "GBps": 4 * n / ms / 1e6}
ms = timed(lambda: x.fill_(1.5)); out["fill_12.6GB"] = {"ms": ms, "GBps": 4 * n / ms / 1e6}
ms = timed(lambda: y.copy_(x[: n // 2])); out["copy_6.3GB_to_6.3GB"] = {"ms": ms, "GBps": 2 * 4 * (n // 2) / ms / 1e6}
print(json.dumps(out))
