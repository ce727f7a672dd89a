from .neuralbase import NeuralBase, NeuralODE, RateFlowODE, UniversalODE

__all__ = ["NeuralBase", "NeuralODE", "RateFlowODE", "UniversalODE"]
