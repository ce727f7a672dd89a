"""``catalax.model.ode`` import path (reference: catalax/model/ode.py)."""
from .model import ODE

__all__ = ["ODE"]
