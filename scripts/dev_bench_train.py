"""Developer timing of the NeuralODE training gradient (ctx_mlp_adj_forward/backward)."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from catalax_b200 import engine
from catalax_b200.neural import NeuralODE

dev = torch.device("cuda:0")
for B in (256, 4096, 65536):
    rng = np.random.default_rng(3)
    net = NeuralODE(4, 64, 3, ["a", "b", "c", "d"], activation="tanh", max_time=10.0, key=3)
    y0 = torch.as_tensor(rng.uniform(0, 2, (B, 4)).astype(np.float32), device=dev)
    ts = torch.linspace(0, 10, 16, device=dev).repeat(B, 1)
    data = y0[:, None, :] * torch.exp(-0.3 * ts)[:, :, None]
    net.value_and_grad(ts, y0, data)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 5
    for _ in range(n):
        val, g = net.value_and_grad(ts, y0, data)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    model, theta = net._model(np.float32)
    th = torch.as_tensor(theta, device=dev)
    ev = [torch.cuda.Event(True) for _ in range(3)]
    ev[0].record(); ctx = model.adjoint_forward(y0, th, ts, dt0=0.0); ev[1].record()
    gq = torch.ones_like(ctx["ys"]); model.adjoint_backward(ctx, gq); ev[2].record(); torch.cuda.synchronize()
    steps = int(ctx["n_accepted"].sum().item())
    print(json.dumps(dict(B=B, wall_ms_value_and_grad=round(wall, 2), fwd_ms=round(ev[0].elapsed_time(ev[1]), 3),
                          bwd_ms=round(ev[1].elapsed_time(ev[2]), 3), accepted_steps=steps, loss=val)), flush=True)
