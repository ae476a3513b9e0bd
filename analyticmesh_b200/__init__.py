"""B200-native Analytic Marching engine (drop-in for the AnalyticMesh hot path)."""
from .model import MLP
from .onnx_io import save_model, load_model

__all__ = ["MLP", "save_model", "load_model"]
