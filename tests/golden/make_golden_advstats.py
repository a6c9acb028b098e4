"""Generate tests/golden/advstats.npz from the REAL reference:  python tests/golden/make_golden_advstats.py

/root/reference/src/training/train_and_save_model.py imports cleanly here (kymatio is optional there), so
`extract_advanced_features` (train…:58-112) is executed as shipped on deterministic inputs: the reference's
seven synthetic patterns (tests/patterns.py) and seeded uint8-grid noise, at 128x128 (the reference's patch
size), 64x64 and 32x32.  This pins the advanced-statistics path to the reference itself."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src/training")

from tests import patterns  # noqa: E402

ref = importlib.import_module("train_and_save_model")
out = {}
for M in (128, 64, 32):
    pats = patterns.all_patterns(M).astype(np.float32)                        # [7, M, M]
    rng = np.random.default_rng(123)
    noise = (rng.integers(0, 256, (5, M, M)) / 255.0).astype(np.float32)
    x = np.concatenate([pats, noise])                                         # 12 single-channel images
    rgb = np.stack([x[0:3], x[3:6], x[6:9], x[9:12]])                          # 4 RGB images [4, 3, M, M]
    out["x%d" % M] = rgb
    out["f%d" % M] = np.stack([ref.extract_advanced_features(im) for im in rgb])   # [4, 54] float64
np.savez_compressed(os.path.join(HERE, "advstats.npz"), **out)
print({k: v.shape for k, v in out.items()})
