from .inaxes import InAxes
from .simconfig import SimulationConfig
from .model import Model

__all__ = ["Model", "SimulationConfig", "InAxes"]
