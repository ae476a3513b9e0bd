"""B200-native Analytic Marching engine (drop-in for the AnalyticMesh hot path)."""
from .model import MLP
from .onnx_io import save_model, load_model

__all__ = ["MLP", "save_model", "load_model"]
from .utils import get_boundary, estimate_am_time, simplify  # noqa: E402
from .polymesh import PolyMesh, poly2tri, get_faces_num, load_ply_header  # noqa: E402


def AnalyticMarching(*args, **kwargs):
    """Lazy wrapper: importing the package must not require the CUDA library (see main.AnalyticMarching)."""
    from .main import AnalyticMarching as _am
    return _am(*args, **kwargs)


__all__ += ["AnalyticMarching", "get_boundary", "estimate_am_time", "simplify", "PolyMesh", "poly2tri", "get_faces_num",
            "load_ply_header"]
