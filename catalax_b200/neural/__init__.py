from .neuralbase import NeuralBase, NeuralODE, RateFlowODE, RBFLayer, UniversalODE

__all__ = ["NeuralBase", "NeuralODE", "RateFlowODE", "UniversalODE", "RBFLayer"]
