from .dataset import Dataset
from .measurement import Measurement

__all__ = ["Dataset", "Measurement"]
