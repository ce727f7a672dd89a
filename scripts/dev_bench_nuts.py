"""Config 4 alone: K2 fused log-likelihood + gradient (used for profiling and A/B runs).
usage: dev_bench_nuts.py [f32|f64] [chains] [likelihood]"""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import bench
from catalax_b200 import engine, workloads as wl
from catalax_b200.tools import codegen as cg

dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
lik = sys.argv[3] if len(sys.argv) > 3 else "softlaplace"
dt = np.float32 if dtype == "f32" else np.float64
dev = torch.device("cuda:0")
field = cg.generate_field(wl.field_spec(wl.michaelis_menten()), sens_theta=True)
model = engine.CompiledModel(field, np.dtype(dt).name)
host = bench.nuts_inputs(CH, 64, 20, dt, 2, dev)
args = [torch.as_tensor(x, device=dev) for x in host]
kw = dict(likelihood=lik, rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096)
for _ in range(5):
    out, _, st = model.loglik_grad(*args, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    out, _, st = model.loglik_grad(*args, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
order = engine.cost_order(*model.last_loglik_counts, heavy_fraction=1.0)
ref = out.clone()
for _ in range(5):
    out, _, st = model.loglik_grad(*args, order=order, **kw)
torch.cuda.synchronize()
e0.record()
for _ in range(30):
    out, _, st = model.loglik_grad(*args, order=order, **kw)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"hinted_ms": round(e0.elapsed_time(e1) / 30, 4), "same_bits": bool(torch.equal(ref, out))}))
print(json.dumps({"dtype": dtype, "chains": CH, "likelihood": lik, "ms": round(ms, 4),
                  "evals_per_s": round(CH / ms * 1e3), "finite": bool(torch.isfinite(out).all()),
                  "checksum": float(out[:, 0].double().sum())}))
