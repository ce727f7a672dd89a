"""Golden records for the host-side Model rules, produced by the reference's REAL
``catalax.model.Model`` class (run in the build container, where /root/reference exists).

``catalax/model/model.py`` and everything it imports from the reference are imported UNMODIFIED.
The third-party packages that are not installed here (jax, equinox, diffrax, numpyro, arviz,
pyenzyme, optax, dotted_dict, matplotlib, ...) are replaced by an import hook that hands out
attribute-absorbing stub modules; only three need behaviour: ``jax.numpy`` -> numpy,
``equinox.Module`` -> an empty base class, ``dotted_dict.DottedDict`` -> a dict with attribute
access, ``bigtree``'s ``Node`` -> a 25-line tree node (name, parent, depth, pre-order descendants).  The Model rules exercised here are pure sympy / pydantic / Python: ordering of states,
parameters, constants and reactions, ``{state}_init`` handling, observable states, inlining of
(nested) assignments, the parameter vector, reactions from scheme strings.
Output: tests/golden/model_facade_reference.json
"""
import importlib
import importlib.abc
import importlib.machinery
import json
import sys
import types
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np
import pandas  # noqa: F401  (real; imported before the hook is installed)
import pydantic  # noqa: F401
import sympy as sp

OUT = Path(__file__).parent / "model_facade_reference.json"
MISSING = {"jax", "matplotlib", "equinox", "numpyro", "diffrax", "arviz", "pyenzyme", "optax", "dotted_dict",
           "xarray", "mlcroissant", "lmfit", "seaborn", "ipywidgets", "gpjax", "corner", "bigtree", "IPython"}


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] not in MISSING:
            return None
        return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class Module:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__()


class DottedDict(dict):
    """dotted_dict.DottedDict: attribute access, nested dicts (also inside lists) wrapped."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, DottedDict):
            return cls(v)
        if isinstance(v, list):
            return [cls._wrap(x) for x in v]
        return v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v


class Node:
    """bigtree.node.node.Node, the five members catalax/model/assignment.py uses: name, a settable
    parent, depth (root = 1), descendants (pre-order, children in insertion order)."""

    def __init__(self, name):
        self.node_name = name
        self.children = []
        self._parent = None

    @property
    def parent(self):
        return self._parent

    @parent.setter
    def parent(self, node):
        if self._parent is not None:
            self._parent.children.remove(self)
        self._parent = node
        if node is not None:
            node.children.append(self)

    @property
    def depth(self):
        return 1 if self._parent is None else self._parent.depth + 1

    @property
    def descendants(self):
        for c in self.children:
            yield c
            yield from c.descendants


def _mod(name, **attrs):
    m = _Stub(name)
    m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _At:
    """``x.at[idx].add(v)`` / ``.set(v)`` of jax arrays: functional scatter updates."""

    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Upd:
            def add(self, v):
                out = arr.copy()
                np.add.at(out, idx, v)
                return out

            def set(self, v):
                out = arr.copy()
                out[idx] = v
                return out

        return _Upd()


class JArray(np.ndarray):
    """numpy array with jax's ``.at`` property (what ``jnp.zeros`` hands out below)."""

    @property
    def at(self):
        return _At(self)


class _NumpyStub(_Stub):
    """jax.numpy -> numpy: every function numpy has, by name (``zeros`` returns a JArray)."""

    def __getattr__(self, name):
        if name == "zeros":
            return lambda *a, **k: np.zeros(*a, **k).view(JArray)
        if hasattr(np, name):
            return getattr(np, name)
        return super().__getattr__(name)


jnp = _NumpyStub("jax.numpy")
jnp.__path__ = []
sys.modules["jax.numpy"] = jnp
def tree_map(f, tree, *rest, is_leaf=None):
    """jax.tree_util.tree_map for the trees the reference builds (single leaves, lists, dicts)."""
    if is_leaf is not None and is_leaf(tree):
        return f(tree)
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, x, is_leaf=is_leaf) for x in tree)
    if isinstance(tree, dict):
        return {k: tree_map(f, v, is_leaf=is_leaf) for k, v in tree.items()}
    return f(tree)


import scipy.special as _sps  # noqa: E402

_mod("jax.tree_util", tree_map=tree_map)
_mod("jax.scipy", special=_sps)
_mod("jax.scipy.special", **{k: getattr(_sps, k) for k in ("erf", "erfc", "gammaln") if hasattr(_sps, k)})
_mod("jax", numpy=jnp, Array=np.ndarray, config=MagicMock(), tree_util=sys.modules["jax.tree_util"],
     scipy=sys.modules["jax.scipy"])
def tree_at(where, pytree, replace):
    """equinox.tree_at for the two shapes the reference uses: ``where`` selects ONE attribute of
    the module (``lambda tree: tree.modules``); a shallow copy with that attribute replaced."""
    import copy as _copy

    target = where(pytree)
    names = [k for k, v in vars(pytree).items() if v is target]
    assert len(names) == 1, names
    out = _copy.copy(pytree)
    object.__setattr__(out, names[0], replace)
    return out


_mod("equinox", Module=Module, field=lambda *a, **k: None, tree_at=tree_at)
_mod("dotted_dict", DottedDict=DottedDict)
_mod("bigtree")
_mod("bigtree.node")
_mod("bigtree.node.node", Node=Node)
sys.meta_path.append(StubFinder())
sys.path.insert(0, "/root/reference")
Model = importlib.import_module("catalax.model.model").Model


def record(model, values=None):
    """Everything the rules decide, as plain JSON."""
    for name, v in (values or {}).items():
        model.parameters[name].value = v
    resolved = model._replace_assignments()
    rec = {
        "state_order": model.get_state_order(),
        "state_order_unmodeled": model.get_state_order(modeled=False),
        "state_indices": model.get_state_order(as_indices=True),
        "parameter_order": model.get_parameter_order(),
        "constants_order": model.get_constants_order(),
        "observable_states": model.get_observable_state_order(),
        "observable_indices": model.get_observable_state_order(as_indices=True),
        "resolved_odes": {k: str(v.equation) for k, v in resolved.odes.items()},
    }
    if len(model.reactions):
        rec["reaction_order"] = model.get_reaction_order()
        rec["reacting_states"] = model.get_reacting_state_order()
        rec["reactions"] = {
            k: {"reactants": [[float(e.stoichiometry), e.state] for e in r.reactants],
                "products": [[float(e.stoichiometry), e.state] for e in r.products],
                "equation": str(resolved.reactions[k].equation), "reversible": bool(r.reversible)}
            for k, r in model.reactions.items()}
    if values is not None:
        rec["parameter_vector"] = [float(x) for x in model._get_parameters()]
    return rec


def _tolist(a):
    """Nested lists with NaN -> None (JSON)."""
    a = np.asarray(a, dtype=float)
    return json.loads(json.dumps(np.where(np.isnan(a), None, a.astype(object)).tolist()))


def dataset_records():
    """the reference's real Dataset / Measurement classes -- ragged measurements,
    a state without data, ``pad``, ``to_jax_arrays`` (with and without ``inits_to_array``),
    ``to_y0_matrix``, observable states."""
    Dataset = importlib.import_module("catalax.dataset.dataset").Dataset
    Measurement = importlib.import_module("catalax.dataset.measurement").Measurement
    import warnings

    ds = Dataset(id="d", name="n", states=["A", "B", "C"])
    ds.add_measurement(Measurement(id="m1", initial_conditions={"A": 1.0, "B": 0.5, "C": 0.0},
                                   time=np.linspace(0, 4, 5), data={"A": np.arange(5.0), "B": np.arange(5.0) * 2}))
    ds.add_measurement(Measurement(id="m2", initial_conditions={"A": 2.0, "B": 0.1, "C": 0.3},
                                   time=np.linspace(0, 2, 3), data={"A": np.arange(3.0) + 10, "B": np.arange(3.0) + 20}))
    ds.add_measurement(Measurement(id="m3", initial_conditions={"A": 3.0, "B": 0.2, "C": 0.4},
                                   time=np.array([0.0, 0.5, 1.5, 4.0]),
                                   data={"A": np.array([1.0, np.nan, 3.0, 4.0]), "B": np.array([5.0, 6.0, 7.0, 8.0])}))
    rec = {"observable_states": ds.get_observable_states_order()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        padded = ds.pad()
        data, time, y0 = padded.to_jax_arrays(["A", "B"], inits_to_array=True)
        rec["padded_AB"] = {"data": _tolist(data), "time": _tolist(time), "y0": _tolist(y0)}
        data, time, y0 = padded.to_jax_arrays(["B", "A", "C"], inits_to_array=True)
        rec["padded_BAC"] = {"data": _tolist(data), "time": _tolist(time), "y0": _tolist(y0)}
        _, _, y0d = padded.to_jax_arrays(["A", "B"])
        rec["inits_dicts"] = [dict(d) for d in y0d]
    # augment (dataset.py:1201-1250, measurement.py:759-817) with jax.random -> the Threefry
    # restatement of catalax_b200/dataset/jaxrandom.py (pinned on jax's known answers separately)
    sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
    from catalax_b200.dataset.jaxrandom import normal as threefry_normal

    jrandom = sys.modules.get("jax.random") or _mod("jax.random")
    jrandom.PRNGKey = lambda seed: int(seed)
    jrandom.normal = lambda key, shape=(), dtype=np.float32: threefry_normal(key, int(np.prod(shape)), dtype).reshape(shape)
    sys.modules["jax"].random = jrandom
    f32 = Dataset(id="d", name="n", states=["A", "B"])
    for mid, scale in (("a1", 1.0), ("a2", 3.0)):
        f32.add_measurement(Measurement(id=mid, initial_conditions={"A": 1.0, "B": 2.0}, time=np.linspace(0, 1, 7),
                                        data={"A": (scale * np.linspace(1, 2, 7)).astype(np.float32),
                                              "B": (scale * np.linspace(3, 5, 7)).astype(np.float32)}))
    for tag, kw in (("mult_append", dict(n_augmentations=2, sigma=0.05, seed=0, append=True, multiplicative=True)),
                    ("add_only", dict(n_augmentations=1, sigma=0.3, seed=7, append=False, multiplicative=False))):
        aug = f32.augment(**kw)
        rec[f"augment_{tag}"] = {"n": len(aug.measurements),
                                 "data": [[[float(v) for v in np.asarray(m.data[s], dtype=np.float64)] for s in ("A", "B")]
                                          for m in aug.measurements]}
    rec["y0_matrix_ABC"] = _tolist(ds.to_y0_matrix(state_order=["A", "B", "C"]))
    rec["y0_matrix_CA"] = _tolist(ds.to_y0_matrix(state_order=["C", "A"]))
    return rec


def main():
    out = {}

    m = Model(name="assignment + constant + init symbol")
    m.add_state(S="Substrate", E="Enzyme", P="Product")
    m.add_constant(I="Inhibitor")
    m.add_assignment("K_app", "K_m * (1 + I / K_i)")
    m.add_assignment("v", "k_cat * E * S / (K_app + S)")          # nested: v uses K_app
    m.add_ode("S", "-v")
    m.add_ode("P", "v * S_init / S_init")
    m.add_ode("E", "0", observable=False)
    out["assignments"] = {"build": "see make_model_facade_golden.py",
                          **record(m, {"K_m": 3.0, "k_cat": 0.5, "K_i": 2.0})}

    m = Model(name="init symbol declared as constant")
    m.add_state("s1, s2")
    m.add_constant(s1_init="initial s1", c0="a constant")
    m.add_ode("s1", "-k * s1 * c0")
    m.add_ode("s2", "k * s1_init - q * s2")
    out["init_constant"] = record(m, {"k": 0.1, "q": 2.0})

    m = Model(name="reactions")
    m.add_state(A="a", B="b", C="c", D="unused")
    m.add_reaction("A + 2 B -> C", symbol="r2", equation="k2 * A * B**2")
    m.add_reaction(symbol="r1", reactants=[(1.0, "C")], products=[(1.0, "A"), (0.5, "B")],
                   equation="k1 * C / (K + C)", reversible=True)
    out["reactions"] = record(m, {"k1": 1.5, "k2": 0.25, "K": 4.0})

    m = Model(name="mixed")
    m.add_state(X="x", Y="y", Z="z")
    m.add_assignment("flux", "vmax * X / (km + X)")
    m.add_reaction("X -> Y", symbol="conv", equation="flux")
    m.add_ode("Z", "kz * Y - dz * Z")
    out["mixed"] = record(m, {"vmax": 2.0, "km": 1.0, "kz": 0.3, "dz": 0.7})

    d = json.loads((Path(__file__).parent / "reference_fixtures" / "uode_model.json").read_text())
    m = Model.from_dict(d)
    out["uode_fixture"] = record(m, None)
    out["uode_fixture"]["parameter_vector"] = [float(x) for x in m._get_parameters()]

    out["__dataset__"] = dataset_records()
    OUT.write_text(json.dumps(out, indent=1, sort_keys=True))
    for k, v in out.items():
        if k.startswith("__"):
            continue
        print(k, {a: v[a] for a in ("state_order", "parameter_order", "constants_order", "observable_states")})
        print("   ", v["resolved_odes"])


if __name__ == "__main__":
    main()
