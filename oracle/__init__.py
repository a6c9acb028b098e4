"""CPU oracle for the WST hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product package (wst_b200) never does.

PARITY UNPINNED: the arithmetic of the reference's hot path lives in the third-party
package kymatio==0.3.0 (reference requirements.txt:18), which is absent from
/root/reference, not installed in this image and not installable (no network).  The
reference itself holds no tests, golden vectors or fixtures for this path.  This
oracle is therefore a restatement of kymatio 0.3.0's published NumPy algorithm,
anchored on the reference's call sites (see kymatio_scattering2d.py), and checked
against closed-form known answers instead of reference-held vectors.
"""
from .kymatio_scattering2d import (  # noqa: F401
    Scattering2D, compute_padding, filter_bank, scattering2d, num_coefficients,
    extract_wst_features_training, extract_wst_features_inference,
    extract_wst_features_visualization, compute_scattering_coefficients,
    pooled_features,
)
