import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from catalax_b200 import engine
from catalax_b200.tools import codegen as cg
from tests import cases
dev = torch.device("cuda:0")
m = engine.CompiledModel(cg.generate_field(cases.field_spec(cases.michaelis_menten())), "float32")
B, T = 1 << 20, 1000
y0, th, c = (torch.as_tensor(x, device=dev) for x in cases.mm_inputs(B, 0, np.float32))
ts = torch.linspace(0, 100, T, dtype=torch.float32, device=dev)
out = torch.empty((B, T, 3), dtype=torch.float32, device=dev)
for it in range(3):
    o = m.solve(y0, th, c, ts, in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, kernel=5, out=out)
torch.cuda.synchronize()
print("ok")
