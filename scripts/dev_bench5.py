"""Config 5: MLP vector field 5-64-64-64-4, 512K trajectories; tensor-core (3) vs CUDA-core (2)."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import catalax_b200 as ctx
from catalax_b200.neural import NeuralODE, RateFlowODE
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
kernels = [int(k) for k in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 3]
for kind in ("neuralode", "rateflow"):
    net = NeuralODE(4, 64, 3, ["a", "b", "c", "d"], activation="tanh", max_time=10.0, key=3) if kind == "neuralode" \
        else RateFlowODE(4, 3, 64, 3, ["a", "b", "c", "d"], activation="celu", max_time=10.0, key=3)
    rng = np.random.default_rng(3)
    y0 = torch.as_tensor(rng.uniform(0, 2, (B, 4)).astype(np.float32), device=dev)
    ts = torch.linspace(0, 10, 16, dtype=torch.float32, device=dev)
    for kernel in kernels:
        times = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); ys = net.predict_arrays(ts, y0, kernel=kernel); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        o = net.last
        steps = int((o.n_accepted.sum() + o.n_rejected.sum()).item()); ms = min(times[1:])
        evals = steps * (7 if kernel == 3 else 6)
        print(json.dumps(dict(workload=kind + " 5-64-64-64-out", kernel=kernel, B=B, ms=round(ms, 2), first_ms=round(times[0], 1),
                              steps=steps, steps_per_traj=round(steps / B, 2), gsteps_s=round(steps / ms / 1e6, 3),
                              mlp_tflops_useful=round(steps * 6 * 17536 / ms / 1e9, 1),
                              fails=int((o.status != 0).sum().item()))), flush=True)
