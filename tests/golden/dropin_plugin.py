"""pytest plugin for running the REFERENCE's own test files against this package (build container
only): ``catalax`` and its sub-packages resolve to ``catalax_b200``; ``jax.numpy`` resolves to
numpy (the reference's tests use it to build and compare arrays).  Used by
tests/test_reference_suite_dropin.py through ``PYTHONPATH=tests/golden python -m pytest -p dropin_plugin <file>``."""
import importlib
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np

import catalax_b200


class _Alias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """``catalax[.x.y]`` -> ``catalax_b200[.x.y]`` (the same module objects)."""

    def find_spec(self, name, path=None, target=None):
        if name == "catalax" or name.startswith("catalax."):
            return importlib.machinery.ModuleSpec(name, self)
        return None

    def create_module(self, spec):
        real = "catalax_b200" + spec.name[len("catalax"):]
        return importlib.import_module(real)

    def exec_module(self, module):
        pass


class _Clamped(np.ndarray):
    """jax arrays clamp out-of-range integer indices instead of raising (the reference's
    ``test_dataset_creation`` reads row 1 of its one-row fixture that way)."""

    def __getitem__(self, idx):
        items = idx if isinstance(idx, tuple) else (idx,)
        fixed = tuple(min(max(i, -n), n - 1) if isinstance(i, (int, np.integer)) else i
                      for i, n in zip(items, self.shape))
        return np.ndarray.__getitem__(self, fixed + tuple(items[len(fixed):]))


class _Jnp(types.ModuleType):
    def __getattr__(self, name):
        if name == "load":
            return lambda *a, **k: np.load(*a, **k).view(_Clamped)
        if hasattr(np, name):
            return getattr(np, name)
        raise AttributeError(name)


sys.meta_path.insert(0, _Alias())
sys.modules["catalax"] = catalax_b200
jnp = _Jnp("jax.numpy")
jax = types.ModuleType("jax")
jax.numpy = jnp
jax.Array = np.ndarray
sys.modules.setdefault("jax", jax)
sys.modules.setdefault("jax.numpy", jnp)
