"""Dataset: list of measurements + dense-array export (reference: dataset/dataset.py).

Only the members the simulate / log-density path touches are provided:
``from_model`` :672-692, ``add_initial`` :188-206, ``add_measurement`` :208-231,
``add_from_jax_array`` :233-254, ``from_jax_arrays`` :767-821, ``to_jax_arrays`` :390-434,
``to_y0_matrix`` :436-451, ``to_config`` :1363-1392, ``pad`` :256-330.
File formats (Croissant, EnzymeML, pandas) and plotting are out of scope.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional

import numpy as np

from ..config import default_dtype
from .measurement import Measurement


class Dataset:
    def __init__(self, states: List[str], id: Optional[str] = None, name: Optional[str] = None,
                 description: Optional[str] = None, measurements: Optional[List[Measurement]] = None):
        self.states = list(states)
        self.id = id
        self.name = name
        self.description = description
        self.measurements: List[Measurement] = list(measurements or [])

    def __len__(self):
        return len(self.measurements)

    @classmethod
    def from_model(cls, model) -> "Dataset":
        states = model.get_state_order(modeled=False)
        return cls(id=model.name, name=model.name, states=states)

    def add_initial(self, time=None, **kwargs) -> None:
        self.measurements.append(Measurement(time=time, initial_conditions=kwargs))

    def add_measurement(self, measurement: Measurement) -> None:
        assert not any(m.id == measurement.id for m in self.measurements), \
            f"Measurement with ID={measurement.id} already exists."
        unused = [s for s in measurement.data.keys() if s not in self.states]
        if unused:
            warnings.warn(f"States {unused} are not used in the dataset.")
        self.measurements.append(measurement)

    def add_from_jax_array(self, state_order: List[str], initial_condition: Dict[str, float],
                           data, time) -> None:
        data = np.asarray(data)
        self.add_measurement(Measurement(
            initial_conditions=initial_condition, time=time,
            data={s: data[:, i] for i, s in enumerate(state_order)}))

    @classmethod
    def from_jax_arrays(cls, state_order: List[str], data, time, y0s) -> "Dataset":
        data, time, y0s = np.asarray(data), np.asarray(time), np.asarray(y0s)
        assert data.shape[0] == time.shape[0] == y0s.shape[0], \
            "Data, time and initial conditions must have the same leading dimension"
        ds = cls(states=list(state_order))
        for i in range(data.shape[0]):
            ds.add_from_jax_array(state_order=state_order, data=data[i], time=time[i],
                                  initial_condition={s: float(y0s[i, j]) for j, s in enumerate(state_order)})
        return ds

    def get_observable_states_order(self) -> List[str]:
        """Sorted states that carry data in every measurement (dataset.py:133-172)."""
        if not self.measurements:
            return []
        common = set(self.measurements[0].data.keys())
        for meas in self.measurements[1:]:
            if set(meas.data.keys()) != common:
                common &= set(meas.data.keys())
                warnings.warn("Inconsistent observables across measurements. "
                              f"Using the common ones: {sorted(common)}")
        if not common:
            raise ValueError("No common observable states across measurements")
        return sorted(common)

    def to_jax_arrays(self, state_order: List[str], inits_to_array: bool = False):
        data, time, inits = [], [], []
        for meas in self.measurements:
            d, t, i0 = meas.to_jax_arrays(state_order=state_order, inits_to_array=inits_to_array)
            data.append(d); time.append(t); inits.append(i0)
        if inits_to_array:
            return np.stack(data, 0), np.stack(time, 0), np.stack(inits, 0)
        return np.stack(data, 0), np.stack(time, 0), inits

    def to_y0_matrix(self, state_order: List[str]) -> np.ndarray:
        if len(self.measurements) == 0:
            return np.zeros((0, len(state_order)), dtype=default_dtype())
        return np.stack([m.to_y0_array(state_order=state_order) for m in self.measurements], 0) \
            .reshape(len(self.measurements), len(state_order))

    def to_config(self, nsteps: int = 100):
        """SimulationConfig spanning each measurement's time range (dataset.py:1363-1392)."""
        from ..model.simconfig import SimulationConfig

        t0 = np.array([m.time[0] for m in self.measurements], dtype=default_dtype())
        t1 = np.array([m.time[-1] for m in self.measurements], dtype=default_dtype())
        return SimulationConfig(t0=t0, t1=t1, nsteps=nsteps)

    def augment(self, n_augmentations: int, sigma: float = 0.5, seed: int = 42, append: bool = True,
                multiplicative: bool = False) -> "Dataset":
        """Noise-augmented copies of every measurement (dataset.py:1201-1250): augmentation ``i``
        uses ``seed + i`` for ALL measurements; ``append`` keeps the originals in front."""
        out = Dataset(states=list(self.states), id=self.id, name=self.name)
        if append:
            out.measurements = list(self.measurements)
        for i in range(n_augmentations):
            for m in self.measurements:
                out.measurements.append(m.augment(sigma=sigma, seed=seed + i, multiplicative=multiplicative))
        return out

    def pad(self) -> "Dataset":
        """A copy with uniform array lengths (dataset.py:256-330): the maximum length is the longest
        ``time`` array (measurements without time are ignored for it); every measurement gets a
        ``data`` entry for every state of the dataset -- right-padded with NaN, all-NaN when the
        state has no data; ``time`` continues monotonically (last+1, last+2, ...; 0, 1, ... when it
        is empty) and stays ``None`` when it was ``None``.  The original is left untouched."""
        max_len = max((len(m.time) for m in self.measurements if m.time is not None), default=0)
        for m in self.measurements:
            missing = sorted(set(self.states) - set(m.initial_conditions))
            if missing:
                raise ValueError(f"Measurement {m.id} missing definition of initial condition "
                                 f"for states: {missing}")
        out = Dataset(states=list(self.states), id=self.id, name=self.name)
        dt = default_dtype()
        for m in self.measurements:
            time = m.time
            if time is not None and len(time) < max_len:
                start = time[-1] + 1.0 if len(time) else 0.0
                time = np.concatenate([time, start + np.arange(max_len - len(time), dtype=dt)])
            data = {k: v for k, v in m.data.items() if k not in self.states}
            for s in self.states:
                v = m.data.get(s)
                if v is None:
                    data[s] = np.full(max_len, np.nan, dtype=dt)
                else:
                    data[s] = np.concatenate([v, np.full(max(max_len - len(v), 0), np.nan, dtype=dt)])
            padded = Measurement(dict(m.initial_conditions), None, time, id=m.id, name=m.name)
            padded.data = data                       # lengths need not match a ``None`` time
            out.measurements.append(padded)
        return out
