"""Measurement: initial conditions + optional time series (reference: dataset/measurement.py:33-339)."""
from __future__ import annotations

import uuid
from typing import Dict, List, Optional

import numpy as np

from ..config import default_dtype


class Measurement:
    def __init__(self, initial_conditions: Dict[str, float], data: Optional[Dict[str, np.ndarray]] = None,
                 time=None, id: Optional[str] = None, name: Optional[str] = None):
        self.initial_conditions = {k: float(v) for k, v in dict(initial_conditions).items()}
        self.data = {k: np.asarray(v, dtype=default_dtype()).reshape(-1) for k, v in (data or {}).items()}
        self.time = None if time is None else np.asarray(time, dtype=default_dtype()).reshape(-1)
        self.id = id or str(uuid.uuid4())
        self.name = name
        for k, v in self.data.items():
            if self.time is not None and v.shape[0] != self.time.shape[0]:
                raise ValueError(f"Data of state '{k}' and time have different lengths")

    @property
    def state(self) -> List[str]:
        return list(self.data.keys())

    def add_data(self, state: str, data, initial_concentration: Optional[float] = None):
        self.data[state] = np.asarray(data, dtype=default_dtype()).reshape(-1)
        if initial_concentration is not None:
            self.initial_conditions[state] = float(initial_concentration)

    def to_jax_arrays(self, state_order: List[str], inits_to_array: bool = False):
        """(data [T,S], time [T], inits) in the given state order (measurement.py:276-328)."""
        if len(self.data) == 0:
            raise ValueError(
                f"The measurement data is empty. Please add data to the measurement {self.id}")
        missing = [s for s in state_order if s not in self.data]
        if missing:
            raise ValueError("The measurement states are inconsistent with the dataset states."
                             f"Missing {missing} in measurement {self.id}")
        data = np.stack([self.data[s] for s in state_order], axis=1)
        time = np.asarray(self.time, dtype=default_dtype())
        inits = (np.array([self.initial_conditions[s] for s in state_order], dtype=default_dtype())
                 if inits_to_array else self.initial_conditions)
        return data, time, inits

    def augment(self, sigma: float, seed: int, multiplicative: bool = False) -> "Measurement":
        """A copy with Gaussian noise on every state's data (measurement.py:759-817): the noise
        vector is ``jax.random.normal(PRNGKey(seed), x.shape)`` -- the same key for every state,
        as in the reference -- applied as ``x * (noise * sigma + 1)`` or ``x + noise * sigma``.
        Initial conditions and time are kept; the copy gets a new id."""
        from . import jaxrandom

        data = {}
        for state, x in self.data.items():
            noise = jaxrandom.normal(int(seed), x.shape[0], x.dtype)
            scale = x.dtype.type(sigma)
            data[state] = x * (noise * scale + x.dtype.type(1.0)) if multiplicative else x + noise * scale
        return Measurement(dict(self.initial_conditions), data, self.time, id=str(uuid.uuid4()))

    def to_y0_array(self, state_order: List[str]) -> np.ndarray:
        return np.array([self.initial_conditions[s] for s in state_order], dtype=default_dtype())
