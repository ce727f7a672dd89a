"""Extract the measurement arrays of the reference's ``examples/datasets/enzymeml_inactivation.json``
(the input of ``examples/InhomogeneousData.ipynb``) without pyenzyme.

Run here (the GPU box has no /root/reference):  python tests/golden/make_enzymeml_golden.py
Writes tests/golden/enzymeml_inactivation.npz: states (sorted ids of the species that carry data),
data [3, 19, 3], times [3, 19], y0 [3, 3] (``initial`` of every species), odes (the document's
``equations`` of type "ode", in ``states`` order).
"""
import json
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/examples/datasets/enzymeml_inactivation.json")
OUT = Path(__file__).parent / "enzymeml_inactivation.npz"


def main():
    doc = json.loads(SRC.read_text())
    with_data = {sd["species_id"] for m in doc["measurements"] for sd in m["species_data"] if sd["data"]}
    without = {sd["species_id"] for m in doc["measurements"] for sd in m["species_data"] if not sd["data"]}
    states = sorted(with_data - without)             # enzymeml.py:384-401 (observables)
    M = len(doc["measurements"])
    T = max(len(sd["data"]) for m in doc["measurements"] for sd in m["species_data"])
    data = np.full((M, T, len(states)), np.nan)
    times = np.full((M, T), np.nan)
    y0 = np.zeros((M, len(states)))
    for i, m in enumerate(doc["measurements"]):
        for sd in m["species_data"]:
            if sd["species_id"] in states:
                j = states.index(sd["species_id"])
                data[i, : len(sd["data"]), j] = sd["data"]
                times[i, : len(sd["time"])] = sd["time"]
                y0[i, j] = sd["initial"]
    odes = {e["species_id"]: e["equation"] for e in doc["equations"] if e["equation_type"] == "ode"}
    assert sorted(odes) == states and np.isfinite(data).all()
    np.savez_compressed(OUT, states=np.array(states), data=data, times=times, y0=y0,
                        odes=np.array([odes[s] for s in states]))
    print("wrote", OUT, data.shape, states, [odes[s] for s in states])


if __name__ == "__main__":
    main()
