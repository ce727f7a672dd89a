from . import error, priors
from .mcmc import BayesianModel, MCMCConfig, extract_dataset_components

__all__ = ["BayesianModel", "MCMCConfig", "priors", "error", "extract_dataset_components"]
