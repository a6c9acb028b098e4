"""wst_b200 — B200-native 2-D wavelet scattering feature extraction (see wst_b200/__init__.py alias)."""
from ._api import *  # noqa: F401,F403
from ._api import __all__  # noqa: F401
from . import numpy, torch  # noqa: E402,F401  kymatio-style frontends: wst_b200.numpy.Scattering2D, wst_b200.torch.Scattering2D
from . import _ops  # noqa: E402,F401  registers torch.ops.wst.scattering2d_features / scattering2d_maps
