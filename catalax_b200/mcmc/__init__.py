from . import error, priors
from .mcmc import BayesianModel, MCMCConfig, extract_dataset_components
from .sampler import HMC, HMCResults, run_mcmc

__all__ = ["BayesianModel", "MCMCConfig", "HMC", "HMCResults", "run_mcmc", "priors", "error",
           "extract_dataset_components"]
