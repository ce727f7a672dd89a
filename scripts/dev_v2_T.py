"""v1 vs v2 over T (developer tool)."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from catalax_b200 import engine
from catalax_b200.tools import codegen as cg
from tests import cases
dev = torch.device("cuda:0")
m = engine.CompiledModel(cg.generate_field(cases.field_spec(cases.michaelis_menten())), "float32")
B = 1 << 20
y0, th, c = (torch.as_tensor(x, device=dev) for x in cases.mm_inputs(B, 0, np.float32))
for T in (128, 250, 384, 512, 640, 768, 1000, 2000):
    ts = torch.linspace(0, 100, T, dtype=torch.float32, device=dev)
    out = torch.empty((B, T, 3), dtype=torch.float32, device=dev)
    r = {}
    for kernel in (2, 5):
        tm = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); m.solve(y0, th, c, ts, in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, kernel=kernel, out=out); e1.record()
            torch.cuda.synchronize(); tm.append(e0.elapsed_time(e1))
        r[kernel] = min(tm[1:])
    print(json.dumps(dict(T=T, v1=round(r[2], 3), v2=round(r[5], 3))), flush=True)
