"""Developer ablation of ctx_solve_v2 at BASELINE configs[1] (f32, B=2^20, T=1000): full kernel,
without the global stores of full groups, and with consumers that only hand slots back (stepping +
hand-off only).  Results other than the first are NOT valid trajectories."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from catalax_b200 import engine
from catalax_b200.tools import codegen as cg
from tests import cases

dev = torch.device("cuda:0")
field = cg.generate_field(cases.field_spec(cases.michaelis_menten()))
B, T = 1 << 20, 1000
y0, th, c = (torch.as_tensor(x, device=dev) for x in cases.mm_inputs(B, 0, np.float32))
ts = torch.linspace(0, 100, T, dtype=torch.float32, device=dev)
out = torch.empty((B, T, 3), dtype=torch.float32, device=dev)
os.environ["CTX_B200_NO_DISK_CACHE"] = "1"
for name, flags in (("full", ""), ("no_store", "-DV2_DBG_NOSTORE=1"), ("no_consume", "-DV2_DBG_NOCONSUME=1")):
    os.environ["CTX_B200_NVRTC_FLAGS"] = flags
    m = engine.CompiledModel(field, "float32")
    times = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        o = m.solve(y0, th, c, ts, in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, kernel=5, out=out)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    print(json.dumps({"variant": name, "ms": round(min(times[1:]), 3),
                      "steps": int((o.n_accepted.sum() + o.n_rejected.sum()).item())}), flush=True)
