"""Developer micro-benchmark of the solve kernels (not the judged bench.py)."""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from catalax_b200 import engine
from catalax_b200.tools import codegen as cg
from tests import cases

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=1 << 20)
ap.add_argument("--T", type=int, nargs="+", default=[16, 1000])
ap.add_argument("--dtype", nargs="+", default=["float32", "float64"])
ap.add_argument("--kernel", type=int, nargs="+", default=[1, 2])
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--rtol", type=float, default=1e-6)
ap.add_argument("--neural", action="store_true")
ap.add_argument("--hint", action="store_true", help="cost-ordered schedule from a previous solve")
ap.add_argument("--probe", action="store_true", help="switch the library's scheduling probe on")
ap.add_argument("--km", type=float, nargs=2, default=None, help="K_m range (default: the benchmark's 0.5 .. 50)")
a = ap.parse_args()
dev = torch.device("cuda:0")
field = cg.generate_field(cases.field_spec(cases.michaelis_menten()))
for dtype in a.dtype:
    m = engine.CompiledModel(field, dtype)
    y0, th, c = cases.mm_inputs(a.B, 0, np.dtype(dtype), **({'km_range': tuple(a.km)} if a.km else {}))
    y0, th, c = (torch.as_tensor(x, device=dev) for x in (y0, th, c))
    for T in a.T:
        ts = torch.linspace(0, 100, T, dtype=y0.dtype, device=dev)
        out_buf = torch.empty((a.B, T, 3), dtype=y0.dtype, device=dev)
        for kernel in a.kernel:
            times = []
            order = None
            if a.hint:
                order = m.solve(y0, th, c, ts, in_axes=(0, 0, 0, None), rtol=a.rtol, atol=a.rtol, kernel=kernel, out=out_buf).cost_order()
                ref_bits = out_buf.clone() if a.B * T <= (1 << 28) else None
            for it in range(a.iters + 1):
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                o = m.solve(y0, th, c, ts, in_axes=(0, 0, 0, None), rtol=a.rtol, atol=a.rtol, kernel=kernel, out=out_buf, order=order, probe=a.probe)
                e1.record(); torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            steps = int((o.n_accepted.sum() + o.n_rejected.sum()).item())
            per_row = (o.n_accepted + o.n_rejected).float()
            tail = dict(max=int(per_row.max().item()), p999=int(torch.quantile(per_row[:1 << 20], 0.999).item()),
                        over300=int((per_row > 300).sum().item()),
                        checksum=float(torch.nan_to_num(out_buf[:, -1, :].double(), posinf=0.0).sum().item()))
            ms = min(times[1:])
            byt = a.B * T * 3 * y0.element_size()
            print(json.dumps(dict(dtype=dtype, T=T, kernel=kernel, ms=round(ms, 3), first_ms=round(times[0], 1),
                                  steps=steps, steps_per_traj=round(steps / a.B, 1), rej_frac=round(float(o.n_rejected.sum().item()) / steps, 3),
                                  gsteps_s=round(steps / ms / 1e6, 2), out_GBs=round(byt / ms / 1e6, 1),
                                  fails=int((o.status != 0).sum().item()), tail=tail, hint=bool(a.hint), probe=a.probe, km=a.km,
                                  same_bits=(bool(torch.equal(ref_bits, out_buf)) if (a.hint and ref_bits is not None) else None))), flush=True)

# ---- config 5: neural vector field
if "--neural" in sys.argv:
    import catalax_b200 as ctx
    from catalax_b200.neural import NeuralODE
    for x64 in (False,):
        ctx.enable_x64(x64)
        net = NeuralODE(4, 64, 3, ["a", "b", "c", "d"], activation="tanh", max_time=10.0, key=3)
        Bn = 1 << 19
        rng = np.random.default_rng(3)
        y0 = torch.as_tensor(rng.uniform(0, 2, (Bn, 4)).astype(np.float64 if x64 else np.float32), device=dev)
        tsn = torch.linspace(0, 10, 16, dtype=y0.dtype, device=dev)
        times = []
        for it in range(3):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); ys = net.predict_arrays(tsn, y0); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        o = net.last
        steps = int((o.n_accepted.sum() + o.n_rejected.sum()).item())
        ms = min(times[1:])
        print(json.dumps(dict(workload="neuralode 5-64-64-64-4 tanh", x64=x64, B=Bn, ms=round(ms, 2), first_ms=round(times[0], 1), steps=steps,
                              gsteps_s=round(steps / ms / 1e6, 3), tflops=round(steps * 6 * 17536 / ms / 1e9, 2),
                              fails=int((o.status != 0).sum().item()))), flush=True)
