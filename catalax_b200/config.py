"""Process-wide switches mirroring catalax/__init__.py:32-70 (numpyro toggles in the reference)."""
from __future__ import annotations

import numpy as np

_X64 = False


def enable_x64(use_x64: bool = True) -> None:
    """Flip the engine between float32 (reference default) and float64 arithmetic."""
    global _X64
    _X64 = bool(use_x64)


def x64_enabled() -> bool:
    return _X64


def default_dtype() -> np.dtype:
    return np.dtype(np.float64 if _X64 else np.float32)


def set_platform(platform: str = "gpu") -> None:
    """The engine is GPU-only; 'cpu' is rejected loudly (no CPU fallback)."""
    assert platform in ["cpu", "gpu"], "platform must be one of 'cpu' or 'gpu'"
    if platform == "cpu":
        raise RuntimeError("catalax_b200 is a B200-native engine: there is no CPU path")


def set_host_count(n: int = 1) -> None:
    """Host-device count is a JAX/pmap notion; multi-GPU runs use one process per GPU."""
    if n < 1:
        raise ValueError("n must be >= 1")
