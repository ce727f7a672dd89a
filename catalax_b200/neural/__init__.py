from .neuralbase import NeuralBase, NeuralODE, RateFlowODE, RBFLayer, UniversalODE
from . import penalties
from .trainer import AdaBelief, Adam, Modes, Penalties, Step, Strategy, train_neural_ode

__all__ = ["NeuralBase", "NeuralODE", "RateFlowODE", "UniversalODE", "RBFLayer", "Strategy", "Step",
           "Modes", "Penalties", "penalties", "AdaBelief", "Adam", "train_neural_ode"]
