"""``catalax.model.state`` import path (reference: catalax/model/state.py)."""
from .model import State

__all__ = ["State"]
