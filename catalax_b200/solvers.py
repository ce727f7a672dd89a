"""Stand-ins for the diffrax solver CLASSES the reference passes around
(``SimulationConfig.solver`` is a class, simconfig.py:16; instantiated at simulation.py:261).
Only the solvers the hot path distinguishes are provided (simulation.py:23-26)."""
from __future__ import annotations


class AbstractSolver:
    name = "abstract"
    adaptive = True

    def __repr__(self):
        return f"{type(self).__name__}()"


class Tsit5(AbstractSolver):
    name = "tsit5"


class Dopri5(AbstractSolver):
    name = "dopri5"


class Euler(AbstractSolver):
    name = "euler"
    adaptive = False


class Heun(AbstractSolver):
    name = "heun"
    adaptive = False


NON_ADAPTIVE_SOLVERS = [Euler, Heun]


def solver_name(solver) -> str:
    """Accept a class, an instance or a string."""
    if isinstance(solver, str):
        n = solver.lower()
    elif isinstance(solver, type):
        n = getattr(solver, "name", solver.__name__.lower())
    else:
        n = getattr(solver, "name", type(solver).__name__.lower())
    if n not in ("tsit5", "dopri5", "euler", "heun"):
        raise NotImplementedError(f"solver {solver!r} is not supported by the B200 engine "
                                  "(Tsit5, Dopri5, Euler, Heun)")
    return n
