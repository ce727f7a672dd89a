"""Extract the reference's own simulated trajectories from its example dataset.

``/root/reference/examples/datasets/croissant_dataset.zip`` was written by the reference's
``examples/GenerateData.ipynb`` (cells 3-7): ``Model.simulate(dataset, SimulationConfig(nsteps=20,
t0=0, t1=100))`` of ``ds1/dt = -v_max*s1/(K_m+s1)`` with ``K_m=100, v_max=7`` for 120 initial
conditions (``np.random.normal`` around 300 / 200 / 80 / 50), default config (Tsit5, rtol=atol=1e-5,
dt0=0.1, f32).  The values are therefore REAL OUTPUTS of the reference's diffrax path at its
default tolerance -- the only such trajectories anywhere under /root/reference that are not
converged to rounding (``tests/fixtures/data.npy`` is), so their discretisation-error pattern is
a fingerprint of diffrax's step sequence and dense interpolant.

Run here (the GPU box has no /root/reference):  python tests/golden/make_croissant_golden.py
Writes tests/golden/croissant_menten.npz: data [120, 20] f32 (sorted by file name), times [20] f32.
"""
import json
import zipfile
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/examples/datasets/croissant_dataset.zip")
OUT = Path(__file__).parent / "croissant_menten.npz"


def main():
    z = zipfile.ZipFile(SRC)
    meta = json.loads(z.read("croissant.json"))
    inits = {}
    for rs in meta["recordSet"]:
        if rs["@id"].endswith("/inits"):
            (key, val), = rs["data"].items()
            inits[key.split("/")[0]] = val
    names = sorted(n for n in z.namelist() if n.endswith(".jsonl"))
    data, times, y0 = [], [], []
    for n in names:
        rows = [json.loads(line) for line in z.read(n).decode().splitlines() if line.strip()]
        data.append([r["s1"] for r in rows])
        times.append([r["time"] for r in rows])
        y0.append(inits[n[: -len(".jsonl")]])
    data = np.asarray(data, np.float64)
    times = np.asarray(times, np.float64)
    assert (times == times[0]).all()
    assert np.array_equal(data.astype(np.float32).astype(np.float64), data)      # f32 values
    assert np.array_equal(np.asarray(y0), data[:, 0])                # inits == first point
    np.savez_compressed(OUT, data=data.astype(np.float32), times=times[0].astype(np.float32),
                        ids=np.array([n[: -len(".jsonl")] for n in names]))
    print("wrote", OUT, data.shape, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
