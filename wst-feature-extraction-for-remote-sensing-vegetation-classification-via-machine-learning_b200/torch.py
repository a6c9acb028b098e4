"""`from wst_b200.torch import Scattering2D` replaces `from kymatio.torch import Scattering2D`
(inference.py:39)."""
from ._api import ScatteringTorch2D as Scattering2D  # noqa: F401

__all__ = ["Scattering2D"]
