"""wst_b200 — B200-native 2-D wavelet scattering feature extraction (see wst_b200/__init__.py alias)."""
from ._api import *  # noqa: F401,F403
from ._api import __all__  # noqa: F401
