"""Config 3 timing stability: per-iteration device times of ctx_solve_w (f32), 30 iterations,
with a preallocated output, straight after start-up and again after an FMA-peak burst."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from catalax_b200 import engine
from catalax_b200.tools import codegen as cg
from tests import cases
dev = torch.device("cuda:0")
field = cg.generate_field(cases.field_spec(cases.mass_action_network(32)))
B = 262144
m = engine.CompiledModel(field, "float32")
rng = np.random.default_rng(1)
y0 = torch.as_tensor(rng.uniform(0.1, 2.0, (B, 32)).astype("float32"), device=dev)
k = torch.as_tensor(np.exp(rng.uniform(np.log(0.01), np.log(1.0), (B, 64))).astype("float32"), device=dev)
c = torch.empty((B, 0), dtype=y0.dtype, device=dev)
ts = torch.linspace(0, 10, 32, dtype=y0.dtype, device=dev)
out = torch.empty((B, 32, 32), dtype=y0.dtype, device=dev)

def series(n=30):
    times = []
    for it in range(n):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); m.solve(y0, k, c, ts, in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, out=out); e1.record()
        torch.cuda.synchronize(); times.append(round(e0.elapsed_time(e1), 3))
    return times

print(json.dumps({"cold": series()}))
engine.measure_fma_peak("float32")
print(json.dumps({"after_fma_peak": series()}))
# back-to-back without a sync between launches (what bench.py's timed() does)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(20):
    m.solve(y0, k, c, ts, in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, out=out)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"back_to_back_avg_ms": round(e0.elapsed_time(e1) / 20, 3)}))
