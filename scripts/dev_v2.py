"""Developer check of the warp-specialised kernel (5) against the persistent kernel (2): bitwise."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from catalax_b200 import engine
from catalax_b200.tools import codegen as cg
from tests import cases
dev = torch.device("cuda:0")
field = cg.generate_field(cases.field_spec(cases.michaelis_menten()))
dtypes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["float32"]
for dtype in dtypes:
    m = engine.CompiledModel(field, dtype)
    for B, T in ((3001, 37), (1 << 16, 1000), (1 << 20, 16), (1 << 20, 1000)):
        y0, th, c = cases.mm_inputs(B, 0, np.dtype(dtype))
        if B == 3001:
            th[:5, 0] = 1e-3
        y0, th, c = (torch.as_tensor(x, device=dev) for x in (y0, th, c))
        ts = torch.linspace(0, 100, T, dtype=y0.dtype, device=dev)
        res = {}
        for kernel in (2, 5):
            out = torch.empty((B, T, 3), dtype=y0.dtype, device=dev)
            times = []
            for it in range(3):
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                o = m.solve(y0, th, c, ts, in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, kernel=kernel,
                            out=out, max_steps=512 if B == 3001 else 4096)
                e1.record(); torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            res[kernel] = (o.ys.clone(), o.status.clone(), o.n_accepted.clone(), min(times[1:]))
        same = bool(torch.equal(res[2][0], res[5][0]) and torch.equal(res[2][1], res[5][1]) and torch.equal(res[2][2], res[5][2]))
        print(json.dumps(dict(dtype=dtype, B=B, T=T, ms_v1=round(res[2][3], 3), ms_v2=round(res[5][3], 3), bit_identical=same,
                              fails=int((res[5][1] != 0).sum().item()))), flush=True)
