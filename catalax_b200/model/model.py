"""Host-side mirror of the reference ``Model`` for the simulate path only.

Same names, argument meaning and error behaviour as ``catalax/model/model.py`` for:
``add_state(s)`` :332-, ``add_constant(s)``, ``add_ode(s)`` :160-223, ``add_reaction`` :225-302,
``add_assignment(s)`` :304-330, ordering rules :882-999, ``simulate`` :511-567,
``_setup_saveat`` :658-672, ``_run_simulation`` :688-702, ``_format_simulation_results``
:704-737, ``_model_changed`` :739-745, ``_setup_system`` :766-786, ``_replace_assignments``
:788-820, ``_get_parameters`` :822-880, ``reset`` :1053-1057.

Everything that is not on the hot path (EnzymeML, HDI plots, JSON export, ...) is out of
scope; ``from_dict``/``load`` read the reference's model JSON so its fixtures can be used.
"""
from __future__ import annotations

import copy
import json
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
from sympy import Expr, Symbol, sympify

from ..config import default_dtype
from ..dataset import Dataset
from ..tools.simulation import Simulation, SimulationInput
from .inaxes import InAxes
from .simconfig import SimulationConfig

# sympy names the reference shadows so that E, S, I, N, O, oo parse as plain symbols
# (catalax/model/utils.py:13-21)
LOCALS = {n: Symbol(n) for n in ("I", "E", "oo", "OO", "O", "S", "N")}


class AttrDict(dict):
    """dict with attribute access (``model.parameters.k_cat.value``; utils.py:24-33)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


@dataclass
class HDI:
    lower: float
    upper: float
    lower_50: float
    upper_50: float
    q: float


@dataclass
class Parameter:
    name: str
    symbol: Expr
    value: Optional[float] = None
    constant: bool = False
    initial_value: Optional[float] = None
    equation: Any = None
    lower_bound: Optional[float] = None
    upper_bound: Optional[float] = None
    hdi: Optional[HDI] = None
    prior: Any = None


@dataclass
class State:
    name: str
    symbol: Expr


@dataclass
class Constant:
    name: str
    symbol: Expr


@dataclass
class ReactionElement:
    state: str
    stoichiometry: float = 1.0

    @classmethod
    def from_tuple(cls, t):
        return cls(state=t[1], stoichiometry=float(t[0]))


class _Equation:
    """Free symbols that are not states / t / *_init / constants / assignments become
    parameters of the model (equation.py:48-89)."""

    def __init__(self, equation):
        if isinstance(equation, str):
            equation = sympify(equation, locals=LOCALS)
        self.equation: Expr = sympify(equation)
        self.inits: frozenset = frozenset()
        self.parameters: AttrDict = AttrDict()
        self._model = None

    def bind(self, model: "Model"):
        self._model = model
        self.add_parameters_to_model()

    def add_parameters_to_model(self):
        m = self._model
        if m is None:
            return
        for symbol in self.equation.free_symbols:
            s = str(symbol)
            if s in m.states or s == "t":
                continue
            if s.endswith("_init"):
                state = s.split("_init")[0]
                if state not in m.states:
                    raise ValueError(
                        f"State '{state}' not found in the model for initial value statement '{symbol}'")
                self.inits = self.inits | {state}
                continue
            if s in m.constants or s in m.assignments:
                continue
            if s in m.parameters:
                self.parameters[s] = m.parameters[s]
                continue
            p = Parameter(name=s, symbol=symbol)
            self.parameters[s] = p
            m.parameters[s] = p


class ODE(_Equation):
    def __init__(self, equation, state: State, observable: bool = True):
        super().__init__(equation)
        self.state = state
        self.observable = observable

    def __repr__(self):
        return f"ODE(state={self.state.symbol}, equation={self.equation})"


class Assignment(_Equation):
    def __init__(self, symbol: str, equation):
        super().__init__(equation)
        self.symbol = symbol


class Reaction(_Equation):
    _ARROWS = [r"<===>", r"<==>", r"<=>", r"<->", r"=>", r"->", r">"]

    def __init__(self, symbol, reactants, products, equation, reversible=False):
        super().__init__(equation)
        self.symbol = symbol
        self.reactants: List[ReactionElement] = list(reactants)
        self.products: List[ReactionElement] = list(products)
        self.reversible = reversible

    @classmethod
    def from_scheme(cls, symbol, schema, equation, reversible, states=None):
        """Parse "2 A + B -> C" style schemes (reaction.py:31-181)."""
        left = right = None
        for pat in cls._ARROWS:
            if re.search(pat, schema):
                parts = re.split(pat, schema, maxsplit=1)
                if len(parts) == 2:
                    left, right = parts[0].strip(), parts[1].strip()
                    break
        if left is None:
            raise ValueError(f"No supported arrow pattern found in schema: {schema}")

        def side(text):
            out = []
            for term in [t.strip() for t in text.split("+") if t.strip()]:
                mt = re.match(r"^(\d+(?:\.\d+)?)?\s*([a-zA-Z_][a-zA-Z0-9_]*)", term)
                if not mt:
                    raise ValueError(f"Could not parse term: {term}")
                coef, sp_ = mt.groups()
                out.append(ReactionElement(state=sp_, stoichiometry=float(coef) if coef else 1.0))
            return out

        reactants, products = side(left), side(right)
        if len(reactants) == 0 or len(products) == 0:
            raise ValueError("Reaction must have at least one reactant and one product")
        if any(e.stoichiometry <= 0 for e in reactants + products):
            raise ValueError("Stoichiometries must be positive and not zero")
        if states is not None:
            undefined = [e.state for e in reactants + products if e.state not in states]
            if undefined:
                raise ValueError(f"States {undefined} not found in states")
        return cls(symbol, reactants, products, equation, reversible)


def check_symbol(symbol: str) -> None:
    if not re.match(r"^[a-zA-Z_][a-zA-Z0-9_]*$", str(symbol)):
        raise ValueError(f"Symbol '{symbol}' is not a valid identifier")


class Model:
    """Kinetic model: states, constants, ODEs / reactions / assignments, parameters."""

    def __init__(self, name: str):
        self.name = name
        self.odes: AttrDict = AttrDict()
        self.reactions: AttrDict = AttrDict()
        self.assignments: AttrDict = AttrDict()
        self.states: AttrDict = AttrDict()
        self.parameters: AttrDict = AttrDict()
        self.constants: AttrDict = AttrDict()
        self._sim_func = None
        self._in_axes = None
        self._dt0 = None
        self._stack = None

    # ------------------------------------------------------------------ construction
    @staticmethod
    def _split(text: str):
        return [s.strip() for s in text.split(",") if s.strip()]

    def add_state(self, state_string: str = "", **state_map) -> List[State]:
        if not all(isinstance(v, str) for v in state_map.values()):
            raise TypeError("State names must be of type str")
        if state_string:
            state_map.update({s: s for s in self._split(state_string)})
        added = []
        for symbol, name in state_map.items():
            check_symbol(symbol)
            self.states[symbol] = State(name=name, symbol=Symbol(symbol))
            added.append(self.states[symbol])
        return added

    def add_states(self, **states):
        for symbol, name in states.items():
            self.add_state(**{symbol: name})

    add_species = add_state

    def add_constant(self, constant_string: str = "", **constants_map):
        if constant_string:
            constants_map.update({c: c for c in self._split(constant_string)})
        for symbol, name in constants_map.items():
            if not isinstance(name, str):
                raise TypeError(f"Constants names must be of type str. Got {type(name)}")
            check_symbol(symbol)
            self.constants[symbol] = Constant(name=name, symbol=Symbol(symbol))

    def add_constants(self, **constants):
        for symbol, name in constants.items():
            self.add_constant(**{symbol: name})

    def add_ode(self, state: str, equation, observable: bool = True):
        if any(str(o.state.name) == state for o in self.odes.values()):
            raise ValueError(
                f"There already exists an ODE for state '{state}'. Please edit the existing ODE instead.")
        if state not in self.states:
            self.add_state(state)
        self.odes[state] = ODE(equation=equation, state=self.states[state], observable=observable)
        self.odes[state].bind(self)

    def add_odes(self, **odes):
        for state, equation in odes.items():
            self.add_ode(state, equation)

    def add_reaction(self, scheme: Optional[str] = None, *, symbol: str,
                     reactants: List[Tuple[float, str]] = (), products: List[Tuple[float, str]] = (),
                     equation, reversible: bool = False):
        reactants, products = list(reactants), list(products)
        if scheme and (reactants or products):
            raise ValueError("Cannot specify both scheme and reactants/products")
        if scheme is not None:
            self.reactions[symbol] = Reaction.from_scheme(
                symbol=symbol, schema=scheme, equation=equation, reversible=reversible,
                states=self.get_state_order(modeled=False))
            self.reactions[symbol].bind(self)
            return
        if len(reactants) == 0 or len(products) == 0:
            raise ValueError("Reaction must have at least one reactant and one product")
        if any(st <= 0 for st, _ in reactants + products):
            raise ValueError("Stoichiometries must be positive and not zero")
        undefined = [s for _, s in reactants + products if s not in self.states]
        if undefined:
            raise ValueError(f"States {undefined} not found in model")
        self.reactions[symbol] = Reaction(
            symbol, [ReactionElement.from_tuple(t) for t in reactants],
            [ReactionElement.from_tuple(t) for t in products], equation, reversible)
        self.reactions[symbol].bind(self)

    def add_assignment(self, symbol: str, equation, **kwargs):
        if symbol in self.assignments:
            raise ValueError(f"Assignment for symbol '{symbol}' already exists")
        self.assignments[symbol] = Assignment(symbol=symbol, equation=equation)
        self.assignments[symbol].bind(self)
        for d in (self.parameters, self.constants, self.states):
            if symbol in d:
                del d[symbol]

    def add_assignments(self, **assignments):
        for symbol, equation in assignments.items():
            self.add_assignment(symbol=symbol, equation=equation)

    # ------------------------------------------------------------------ ordering rules (a1)
    def get_parameter_order(self) -> List[str]:
        return sorted(self.parameters.keys())

    def get_reacting_state_order(self) -> List[str]:
        return sorted({e.state for r in self.reactions.values() for e in r.reactants + r.products})

    def get_state_order(self, modeled: bool = True, as_indices: bool = False):
        if not modeled:
            return list(range(len(self.states))) if as_indices else sorted(self.states.keys())
        ode_species = {str(o.state.symbol) for o in self.odes.values()}
        modeled_species = sorted(ode_species | set(self.get_reacting_state_order()))
        return list(range(len(modeled_species))) if as_indices else modeled_species

    get_species_order = get_state_order

    def get_reaction_order(self) -> List[str]:
        return sorted(self.reactions.keys())

    def get_observable_state_order(self, as_indices: bool = False):
        out = []
        reacting = self.get_reacting_state_order()
        for i, state in enumerate(self.states.keys()):
            if state not in reacting and state not in self.odes:
                continue
            if state in self.odes and not self.odes[state].observable:
                continue
            out.append(i if as_indices else state)
        return sorted(out)

    def _is_init_symbol(self, name: str) -> bool:
        return name.endswith("_init") and name[: -len("_init")] in self.states

    def get_constants_order(self) -> List[str]:
        return sorted(n for n in self.constants.keys() if not self._is_init_symbol(n))

    def n_parameters(self) -> int:
        return len(self.parameters)

    # ------------------------------------------------------------------ hot path
    @property
    def sim_input(self) -> SimulationInput:
        odes = [self.odes[s] for s in self.get_state_order() if s in self.odes]
        reactions = [self.reactions[r] for r in self.get_reaction_order()]
        return SimulationInput(odes=odes, reactions=reactions, states=self.get_state_order(),
                               parameters=self.get_parameter_order(),
                               constants=self.get_constants_order())

    # ------------------------------------------------------------------ rate function (K4)
    def _setup_rate_function(self, in_axes=None):
        """model.py:1027-1051: the vector field as a callable ``f(t, y, (parameters, constants[,
        inits]))`` -- the reference returns its stack (or ``jit(vmap(stack))``); here one ``ctx_rates``
        launch evaluates the generated field, natively batched: ``t`` scalar | [N], ``y`` [S] | [N,S],
        ``parameters`` [P] | [N,P], ``constants`` [C] | [N,C], ``inits`` [S] | [N,S] -> [S] | [N,S].
        ``in_axes`` is accepted for signature compatibility (the callable broadcasts by shape)."""
        assert in_axes is None or len(in_axes) == 4, \
            "Got invalid dimension for 'in_axes' - Needs to have four ints/Nones"
        from .. import engine
        from ..tools import codegen as cg

        dtype = default_dtype()
        field = cg.generate_field(self._replace_assignments().sim_input.field_spec())
        compiled = engine.CompiledModel(field, np.dtype(dtype).name)

        def rate_fun(t, y, args):
            import torch

            parameters, constants, *rest = args
            inits = rest[0] if rest else None
            on_device = isinstance(y, torch.Tensor) and y.is_cuda
            dev = y.device if on_device else torch.device("cuda", torch.cuda.current_device())
            tt = lambda a: (a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=dtype))).to(dev)
            y_t = tt(y)
            single = y_t.dim() == 1
            if single:
                y_t = y_t[None]
            if field.C == 0:
                c_t = torch.empty((0,), dtype=y_t.dtype, device=dev)
            else:
                c_t = tt(constants).reshape(-1, field.C) if np.ndim(constants) > 1 else tt(constants).reshape(field.C)
            out = compiled.rates(tt(t), y_t, tt(parameters), c_t, y0=None if inits is None else tt(inits))
            out = out[0] if single else out
            return out if on_device else out.cpu().numpy()

        return rate_fun

    def rates(self, t, y, constants=None):
        """model.py:605-629: the model's rates at N points, ``t`` [N], ``y`` [N,S], ``constants``
        [N,C] | None -> [N,S], with the model's current parameter values.  (The reference's body
        hands these arguments to its SOLVE function; this does what its docstring says.)"""
        y = np.asarray(y)
        y = y.reshape(-1, 1) if y.ndim == 1 else y
        t = np.asarray(t)
        assert t.shape[0] == y.shape[0], "Time and state must have the same number of rows"
        C = len(self.get_constants_order())
        if constants is None:
            constants = np.zeros((y.shape[0], 0))
        constants = np.asarray(constants)
        constants = constants.reshape(-1, 1) if constants.ndim == 1 else constants
        assert constants.shape[0] == y.shape[0] or C == 0, \
            "Constants must have the same number of rows as time and state"
        f = self._setup_rate_function()
        return f(t, y, (self._get_parameters(), constants.reshape(y.shape[0], C)))

    def predict_rates(self, dataset: Dataset):
        """model.py:631-645: rates at every measured point of a Dataset, [M*T, S]; the constants of
        a measurement apply to all of its time points."""
        data, times, _ = dataset.to_jax_arrays(self.get_state_order())
        data, times = np.asarray(data), np.asarray(times)
        constants = np.asarray(dataset.to_y0_matrix(state_order=self.get_constants_order()))
        M, T, S = data.shape
        return self.rates(times.ravel(), data.reshape(M * T, S), np.repeat(constants.reshape(M, -1), T, axis=0))

    def simulate(self, dataset: Dataset, config: SimulationConfig, saveat=None, parameters=None,
                 in_axes: InAxes = InAxes.Y0, return_array: bool = False, use_hdi=None,
                 hdi_prob: float = 0.95):
        """Simulate the model over a dataset's initial conditions (model.py:511-567)."""
        if config.nsteps is not None and saveat is not None and not isinstance(saveat, type(None)):
            raise ValueError("Cannot specify both nsteps and saveat")
        y0, constants = self._initialize_simulation_data(dataset)
        saveat = self._setup_saveat(config, saveat)
        parameters = self._get_parameters(parameters, use_hdi, hdi_prob)
        inaxes = (in_axes + InAxes.TIME) if saveat.ndim != 1 else in_axes.value
        if self._model_changed(inaxes, config.dt0) or self._sim_func is None:
            self._setup_simulation(config, inaxes)
        result = self._run_simulation(y0, parameters, constants, saveat)
        return self._format_simulation_results(result, saveat, return_array, y0, constants)

    def predict(self, dataset: Dataset, config: Optional[SimulationConfig] = None,
                n_steps: int = 100, use_times: bool = False, hdi=None) -> Dataset:
        """Convenience wrapper around ``simulate`` (model.py:569-603): the configuration comes from
        the dataset when none is given; ``use_times`` solves on the dataset's own time points."""
        if config is None:
            config = dataset.to_config(nsteps=n_steps)
        if use_times:
            config.nsteps = None
            _, times, _ = dataset.to_jax_arrays(self.get_state_order())
        else:
            times = None
        if config.nsteps != n_steps and not use_times:
            config.nsteps = n_steps
        return self.simulate(dataset, config, saveat=times, use_hdi=hdi)

    def _initialize_simulation_data(self, dataset: Dataset):
        return (dataset.to_y0_matrix(state_order=self.get_state_order()),
                dataset.to_y0_matrix(state_order=self.get_constants_order()))

    def _setup_saveat(self, config: SimulationConfig, saveat):
        if config.nsteps is not None:
            saveat = np.linspace(np.asarray(config.t0, dtype=default_dtype()),
                                 np.asarray(config.t1, dtype=default_dtype()),
                                 config.nsteps, dtype=default_dtype()).T
        if saveat is None:
            raise ValueError("No saveat or config.nsteps provided. Please provide one of the two.")
        return np.asarray(saveat, dtype=default_dtype())

    def _setup_simulation(self, config: SimulationConfig, in_axes: Tuple) -> None:
        self._setup_system(in_axes=in_axes, config=config)
        self._in_axes = in_axes
        self._dt0 = config.dt0

    def _run_simulation(self, y0, parameters, constants, saveat):
        if self._sim_func is None:
            raise RuntimeError("Simulation function not initialized")
        result = self._sim_func(y0, parameters, constants, saveat)
        if result.ndim != 3:
            result = result[None]
        return result

    def _format_simulation_results(self, result, saveat, return_array, y0, constants):
        if return_array:
            return saveat, result
        result = np.asarray(result.cpu() if hasattr(result, "cpu") else result)
        dataset = Dataset.from_model(self)
        order = self.get_state_order()
        for i in range(result.shape[0]):
            init = {s: float(y0[i, j]) for j, s in enumerate(order)}
            for j, c in enumerate(self.get_constants_order()):
                init[c] = float(constants[i, j])
            dataset.add_from_jax_array(state_order=order, data=result[i],
                                       time=saveat if saveat.ndim == 1 else saveat[i],
                                       initial_condition=init)
        return dataset

    def _model_changed(self, in_axes: Tuple, dt0: float) -> bool:
        # only in_axes and dt0 are compared -- the reference's cache key (model.py:739-745)
        return self._in_axes != in_axes or self._dt0 != dt0

    def _setup_system(self, config: SimulationConfig, in_axes: Tuple = InAxes.Y0.value):
        model = self._replace_assignments()
        sim = Simulation(sim_input=model.sim_input, config=config)
        self._sim_func, self._stack = sim._prepare_func(in_axes=in_axes)

    def _replace_assignments(self) -> "Model":
        """Inline assignment expressions into ODE / reaction equations (model.py:788-820)."""
        model = copy.copy(self)
        model.odes = AttrDict({k: copy.copy(v) for k, v in self.odes.items()})
        model.reactions = AttrDict({k: copy.copy(v) for k, v in self.reactions.items()})
        resolved = resolve_assignments({k: a.equation for k, a in self.assignments.items()})
        subs = {Symbol(k): v for k, v in resolved.items()}
        if subs:
            for eq in list(model.odes.values()) + list(model.reactions.values()):
                if eq.equation.free_symbols & set(subs):
                    eq.equation = eq.equation.subs(subs)
        return model

    def _get_parameters(self, parameters=None, use_hdi=None, hdi_prob: float = 0.95):
        if parameters is not None:
            return parameters
        missing = [p for p in self.parameters.values() if p.value is None]
        if missing:
            raise ValueError(
                f"Missing values for parameters: {', '.join(p.name for p in missing)}")
        order = self.get_parameter_order()
        if use_hdi in ("lower", "upper", "lower_50", "upper_50"):
            assert all(self.parameters[p].hdi is not None for p in order), \
                "HDI bound requested, but no HDI found for parameters"
            return np.array([getattr(self.parameters[p].hdi, use_hdi) for p in order],
                            dtype=default_dtype())
        if use_hdi is not None:
            raise ValueError("Invalid use_hdi value. Either 'lower' or 'upper' expected.")
        return np.array([self.parameters[p].value for p in order], dtype=default_dtype())

    def reset(self):
        """Drop the cached solve function so the next simulate() rebuilds it (model.py:1053)."""
        self._sim_func = None
        self._in_axes = None
        self._dt0 = None

    # ------------------------------------------------------------------ model JSON (fixtures)
    @classmethod
    def from_dict(cls, d: Dict) -> "Model":
        m = cls(name=d["name"])
        for sp_ in d.get("states", d.get("species", [])):
            m.add_state(**{sp_["symbol"]: sp_.get("name", sp_["symbol"])})
        for c in d.get("constants", []) or []:
            m.add_constant(**{c["symbol"]: c.get("name", c["symbol"])})
        for a in d.get("assignments", []) or []:
            m.add_assignment(symbol=a["symbol"], equation=a["equation"])
        for o in d.get("odes", []) or []:
            m.add_ode(o.get("state", o.get("species")), o["equation"],
                      observable=o.get("observable", True))
        for r in d.get("reactions", []) or []:
            m.add_reaction(symbol=r["symbol"],
                           reactants=[(e["stoichiometry"], e["state"]) for e in r["reactants"]],
                           products=[(e["stoichiometry"], e["state"]) for e in r["products"]],
                           equation=r["equation"], reversible=r.get("reversible", False))
        for p in d.get("parameters", []) or []:
            if p["symbol"] not in m.parameters:
                continue
            q = m.parameters[p["symbol"]]
            for k in ("value", "constant", "initial_value", "lower_bound", "upper_bound"):
                if k in p and p[k] is not None:
                    setattr(q, k, p[k])
            q.prior = p.get("prior")
        return m

    def to_dict(self) -> Dict:
        """Serializable dictionary in the reference's model-JSON schema (model.py:1265-1292;
        ``from_dict`` reads it back).  None-valued attributes are left out (``exclude_none``);
        reactions are written too (the reference's ``to_dict`` drops them)."""
        def clean(d):
            return {k: v for k, v in d.items() if v is not None}

        res = {
            "name": self.name,
            "states": [clean({"name": s.name, "symbol": str(s.symbol)}) for s in self.states.values()],
            "constants": [clean({"name": c.name, "symbol": str(c.symbol)}) for c in self.constants.values()],
            "odes": [{"equation": str(o.equation), "observable": bool(o.observable), "state": k}
                     for k, o in self.odes.items()],
            "parameters": [clean({"name": q.name, "symbol": str(q.symbol), "value": q.value,
                                  "constant": bool(q.constant), "initial_value": q.initial_value,
                                  "lower_bound": q.lower_bound, "upper_bound": q.upper_bound})
                           for q in self.parameters.values()],
            "assignments": [{"symbol": a.symbol, "equation": str(a.equation)}
                            for a in self.assignments.values()],
            "reactions": [{"symbol": r.symbol, "equation": str(r.equation),
                           "reversible": bool(r.reversible),
                           "reactants": [{"state": e.state, "stoichiometry": e.stoichiometry} for e in r.reactants],
                           "products": [{"state": e.state, "stoichiometry": e.stoichiometry} for e in r.products]}
                          for r in self.reactions.values()],
        }
        for key in ("constants", "assignments", "reactions"):
            if not res[key]:
                res.pop(key)
        return res

    def save(self, path: str, name: Optional[str] = None) -> str:
        """Write the model JSON (model.py:1214-1230)."""
        import os

        fname = os.path.join(str(path), f"{(name or self.name).replace(' ', '_')}.json")
        with open(fname, "w") as f:
            json.dump(self.to_dict(), f, indent=2, default=str)
        return fname

    @classmethod
    def load(cls, path) -> "Model":
        with open(path) as f:
            return cls.from_dict(json.load(f))


def resolve_assignments(assignments: Dict[str, Expr]) -> Dict[str, Expr]:
    """Substitute assignments into each other until none refers to another
    (assignment.py:21-; the reference uses a dependency tree, the result is the same)."""
    syms = {Symbol(k) for k in assignments}
    out = dict(assignments)
    for _ in range(len(out) + 1):
        changed = False
        for k, e in out.items():
            deps = e.free_symbols & syms
            if deps:
                if Symbol(k) in deps:
                    raise ValueError(f"Assignment '{k}' depends on itself")
                out[k] = e.subs({d: out[str(d)] for d in deps})
                changed = True
        if not changed:
            return out
    raise ValueError("Circular dependency between assignments")
