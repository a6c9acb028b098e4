"""Importable alias for the package directory
`wst-feature-extraction-for-remote-sensing-vegetation-classification-via-machine-learning_b200/`
(its name is fixed by the project layout but is not a valid Python identifier).
`import wst_b200` resolves every submodule from that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "wst-feature-extraction-for-remote-sensing-vegetation-classification-via-machine-learning_b200")
if not _os.path.isdir(_real):
    raise ImportError("wst_b200: package directory not found: " + _real)
__path__.insert(0, _real)

from ._api import *  # noqa: E402,F401,F403
from ._api import __all__  # noqa: E402,F401
from . import numpy, torch  # noqa: E402,F401  kymatio-style frontends: wst_b200.numpy.Scattering2D, wst_b200.torch.Scattering2D
from . import _ops  # noqa: E402,F401  registers torch.ops.wst.scattering2d_features / scattering2d_maps
