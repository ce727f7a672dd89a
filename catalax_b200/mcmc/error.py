"""Priors over the concentration-space noise std ``sigma_y`` (reference: catalax/mcmc/error.py
:72-181; default HalfNormal(mean(yerrs)) per observable, :181)."""
from __future__ import annotations

import math
from typing import Optional

import torch


class ErrorModel:
    shared = False

    def _resolve(self, value, scale_hint: torch.Tensor) -> torch.Tensor:
        if value is not None:
            return torch.as_tensor(value, dtype=scale_hint.dtype, device=scale_hint.device).expand_as(scale_hint)
        return scale_hint.mean().expand_as(scale_hint) if self.shared else scale_hint

    sampled = True

    def log_prob(self, sigma: torch.Tensor, scale_hint: torch.Tensor) -> torch.Tensor:
        """sigma [CH, n_obs] -> summed log prior [CH]."""
        raise NotImplementedError

    def grad(self, sigma: torch.Tensor, scale_hint: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class HalfNormal(ErrorModel):
    def __init__(self, scale: Optional[float] = None, shared: bool = False):
        self.scale, self.shared = scale, shared

    def log_prob(self, sigma, scale_hint):
        sc = self._resolve(self.scale, scale_hint)
        lp = math.log(2.0) - torch.log(sc) - 0.5 * math.log(2 * math.pi) - 0.5 * (sigma / sc) ** 2
        lp = torch.where(sigma >= 0, lp, torch.full_like(lp, -math.inf))
        return lp.sum(-1)

    def grad(self, sigma, scale_hint):
        return -sigma / self._resolve(self.scale, scale_hint) ** 2


class LogNormal(ErrorModel):
    def __init__(self, median: Optional[float] = None, sigma: float = 0.7, shared: bool = False):
        self.median, self.sigma, self.shared = median, sigma, shared

    def log_prob(self, sigma, scale_hint):
        loc = torch.log(self._resolve(self.median, scale_hint))
        z = (torch.log(sigma) - loc) / self.sigma
        lp = -0.5 * z * z - torch.log(sigma) - math.log(self.sigma) - 0.5 * math.log(2 * math.pi)
        return lp.sum(-1)

    def grad(self, sigma, scale_hint):
        loc = torch.log(self._resolve(self.median, scale_hint))
        return -(torch.log(sigma) - loc) / (self.sigma ** 2 * sigma) - 1.0 / sigma


class Gamma(ErrorModel):
    def __init__(self, concentration: float = 2.0, rate: Optional[float] = None, shared: bool = False):
        self.concentration, self.rate, self.shared = concentration, rate, shared

    def _rate(self, scale_hint):
        if self.rate is not None:
            return torch.as_tensor(self.rate, dtype=scale_hint.dtype, device=scale_hint.device).expand_as(scale_hint)
        return self.concentration / self._resolve(None, scale_hint)

    def log_prob(self, sigma, scale_hint):
        a, r = self.concentration, self._rate(scale_hint)
        lp = a * torch.log(r) - math.lgamma(a) + (a - 1) * torch.log(sigma) - r * sigma
        return lp.sum(-1)

    def grad(self, sigma, scale_hint):
        return (self.concentration - 1) / sigma - self._rate(scale_hint)


class Fixed(ErrorModel):
    sampled = False

    def __init__(self, sigma_y):
        self.sigma_y = sigma_y

    def log_prob(self, sigma, scale_hint):
        return torch.zeros(sigma.shape[:-1], dtype=sigma.dtype, device=sigma.device)

    def grad(self, sigma, scale_hint):
        return torch.zeros_like(sigma)


DEFAULT_ERROR_MODEL: ErrorModel = HalfNormal()
