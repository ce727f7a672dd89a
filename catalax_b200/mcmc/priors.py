"""Parameter priors (reference: catalax/mcmc/priors.py:22-121) as plain log-density + gradient
objects.  The reference wraps numpyro distributions; here each prior provides
``log_prob(x)`` and ``grad(x)`` on torch tensors (tiny scalar work that stays host/torch side,
SURVEY §8 a13)."""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

_SQRT2 = math.sqrt(2.0)
_LOG_SQRT_2PI = 0.5 * math.log(2 * math.pi)


class Prior:
    @property
    def type(self) -> str:
        return self.__class__.__name__

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def grad(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


@dataclass
class Normal(Prior):
    mu: float
    sigma: float

    def log_prob(self, x):
        z = (x - self.mu) / self.sigma
        return -0.5 * z * z - math.log(self.sigma) - _LOG_SQRT_2PI

    def grad(self, x):
        return -(x - self.mu) / self.sigma ** 2


@dataclass
class TruncatedNormal(Prior):
    mu: float
    sigma: float
    low: float = 1e-6
    high: float = 1e6

    def _log_z(self):
        cdf = lambda v: 0.5 * (1 + math.erf((v - self.mu) / (self.sigma * _SQRT2)))
        return math.log(max(cdf(self.high) - cdf(self.low), 1e-300))

    def log_prob(self, x):
        z = (x - self.mu) / self.sigma
        lp = -0.5 * z * z - math.log(self.sigma) - _LOG_SQRT_2PI - self._log_z()
        inside = (x >= self.low) & (x <= self.high)
        return torch.where(inside, lp, torch.full_like(lp, -math.inf))

    def grad(self, x):
        return -(x - self.mu) / self.sigma ** 2


@dataclass
class Uniform(Prior):
    low: float
    high: float

    def log_prob(self, x):
        inside = (x >= self.low) & (x <= self.high)
        return torch.where(inside, torch.full_like(x, -math.log(self.high - self.low)),
                           torch.full_like(x, -math.inf))

    def grad(self, x):
        return torch.zeros_like(x)


@dataclass
class LogUniform(Prior):
    low: float
    high: float

    def log_prob(self, x):
        inside = (x >= self.low) & (x <= self.high)
        lp = -torch.log(x) - math.log(math.log(self.high / self.low))
        return torch.where(inside, lp, torch.full_like(lp, -math.inf))

    def grad(self, x):
        return -1.0 / x


def from_dict(d) -> Prior:
    """Priors as stored in the reference's model JSON ({"type": "Uniform", "low":..,"high":..})."""
    d = dict(d)
    kind = d.pop("type")
    return {"Normal": Normal, "TruncatedNormal": TruncatedNormal, "Uniform": Uniform,
            "LogUniform": LogUniform}[kind](**d)
