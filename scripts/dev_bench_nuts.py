"""Config 4 alone: K2 fused log-likelihood + gradient (used for profiling)."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import argparse, numpy as np, torch
import bench
a = argparse.Namespace(dtype=sys.argv[1] if len(sys.argv) > 1 else "f32", steps=5, nuts_chains=int(sys.argv[2]) if len(sys.argv) > 2 else 16384)
print(json.dumps(bench.run_nuts_leg(a, None, torch.device("cuda:0"), 0, 1, None)))
