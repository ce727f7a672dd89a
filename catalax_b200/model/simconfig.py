from __future__ import annotations

from dataclasses import dataclass
from typing import Any

from ..solvers import Tsit5


@dataclass
class SimulationConfig:
    """Configuration for simulation parameters (fields and defaults of simconfig.py:8-20)."""

    t1: Any
    nsteps: int | None = None
    t0: Any = 0
    dt0: float = 0.1
    solver: Any = Tsit5
    rtol: float = 1e-5
    atol: float = 1e-5
    max_steps: int = 4096
    throw: bool = True
