"""Config 3: synthetic mass-action network 32 states / 64 reactions, 256K trajectories."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from catalax_b200 import engine
from catalax_b200.tools import codegen as cg
from tests import cases
dev = torch.device("cuda:0")
field = cg.generate_field(cases.field_spec(cases.mass_action_network(32)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
for dtype in ("float32", "float64"):
    m = engine.CompiledModel(field, dtype)
    rng = np.random.default_rng(1)
    y0 = torch.as_tensor(rng.uniform(0.1, 2.0, (B, 32)).astype(dtype), device=dev)
    k = torch.as_tensor(np.exp(rng.uniform(np.log(0.01), np.log(1.0), (B, 64))).astype(dtype), device=dev)
    c = torch.empty((B, 0), dtype=y0.dtype, device=dev)
    ts = torch.linspace(0, 10, 32, dtype=y0.dtype, device=dev)
    for kernel in [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else ['1', '4'])]:
        times = []
        for it in range(3):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); o = m.solve(y0, k, c, ts, in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, kernel=kernel); e1.record()
            torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
        steps = int((o.n_accepted.sum() + o.n_rejected.sum()).item()); ms = min(times[1:])
        print(json.dumps(dict(workload="mass_action_32x64", dtype=dtype, kernel=kernel, B=B, ms=round(ms, 2), steps_per_traj=round(steps / B, 1),
                              gsteps_s=round(steps / ms / 1e6, 3), gflops=round(steps * (6 * field.flops_rhs + 64 * 32 + 11) / ms / 1e6, 1),
                              fails=int((o.status != 0).sum().item()))), flush=True)
