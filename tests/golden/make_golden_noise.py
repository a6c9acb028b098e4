"""Generate tests/golden/noise.npz from the REAL reference:  python tests/golden/make_golden_noise.py

/root/reference/src/preprocessing/add_noise.py imports cleanly here (numpy + PIL), so its five noise functions
(add_noise.py:14-72) are executed as shipped, seeded like its main() (np.random.seed(args.seed), :147-149), on a
small deterministic uint8 image at the intensities of the robustness study (gaussian 30/50,
poisson 40/60, salt_and_pepper 5/15/25, speckle 15/35/55, uniform 10/25/40).  This pins the noise path to the
reference itself."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src/preprocessing")
ref = importlib.import_module("add_noise")

CASES = [("gaussian", 30), ("gaussian", 50), ("poisson", 40), ("poisson", 60), ("salt_and_pepper", 5),
         ("salt_and_pepper", 15), ("salt_and_pepper", 25), ("speckle", 15), ("speckle", 35), ("speckle", 55),
         ("uniform", 10), ("uniform", 25), ("uniform", 40)]
FUNCS = {"gaussian": ref.add_gaussian_noise, "salt_and_pepper": ref.add_salt_and_pepper_noise,
         "speckle": ref.add_speckle_noise, "poisson": ref.add_poisson_noise, "uniform": ref.add_uniform_noise}

rng = np.random.default_rng(2024)
img = rng.integers(0, 256, (3, 24, 20, 3), dtype=np.uint8)          # three images, noised one after the other
img[0, :4] = 0
img[0, 4:8] = 255                                                   # saturated rows exercise the clip
out = {"img": img}
for kind, intensity in CASES:
    np.random.seed(42)
    out["%s_%d" % (kind, intensity)] = np.stack([FUNCS[kind](im, intensity) for im in img])
np.savez_compressed(os.path.join(HERE, "noise.npz"), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()})
