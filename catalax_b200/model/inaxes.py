from __future__ import annotations

from enum import Enum


class InAxes(Enum):
    """Batch-axis specifications of the solve seam (inaxes.py:5-27)."""

    Y0 = (0, None, 0, None)
    PARAMETERS = (None, 0, None, None)
    CONSTANTS = (None, None, 0, None)
    TIME = (None, None, None, 0)

    def __add__(self, other: "InAxes"):
        assert isinstance(other, InAxes), "Other spec must be an instance of InAxes"
        return tuple(
            0 if (self.value[i] is not None or other.value[i] is not None) else None
            for i in range(4)
        )
