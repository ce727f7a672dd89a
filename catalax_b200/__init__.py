"""catalax_b200 -- B200-native engine for Catalax's batched ODE-solve hot path.

Public surface mirrors ``catalax/__init__.py``: ``Model``, ``Dataset``, ``Measurement``,
``SimulationConfig``, ``InAxes`` (+ ``INITS/PARAMETERS/TIME``), ``enable_x64``,
``set_platform``, ``set_host_count`` and the solver classes the reference takes from diffrax.
"""
from sympy import Symbol  # noqa: F401

from .config import enable_x64, set_host_count, set_platform, x64_enabled
from .dataset import Dataset, Measurement
from .model import InAxes, Model, SimulationConfig
from .solvers import Dopri5, Euler, Heun, Tsit5
from .tools.optimization import optimize

__all__ = ["SimulationConfig", "Dataset", "Measurement", "InAxes", "Model", "enable_x64",
           "set_platform", "set_host_count", "optimize", "Tsit5", "Dopri5", "Euler", "Heun"]
__version__ = "0.1.0"

PARAMETERS = InAxes.PARAMETERS
TIME = InAxes.TIME
INITS = InAxes.Y0
