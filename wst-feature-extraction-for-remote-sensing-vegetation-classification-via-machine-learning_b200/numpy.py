"""`from wst_b200.numpy import Scattering2D` replaces `from kymatio.numpy import Scattering2D`
(train_and_save_model.py:46, visualize_features.py:30)."""
from ._api import ScatteringNumPy2D as Scattering2D  # noqa: F401

__all__ = ["Scattering2D"]
