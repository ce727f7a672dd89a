"""``catalax.model.reaction`` import path (reference: catalax/model/reaction.py)."""
from .model import Reaction, ReactionElement

__all__ = ["Reaction", "ReactionElement"]
