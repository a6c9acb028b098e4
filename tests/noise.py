"""The reference's five noise models (src/preprocessing/add_noise.py:14-72), restated for test inputs of the
BASELINE noise-robustness configuration (configs[3]).  uint8 HWC in, uint8 HWC out, numpy global RNG like the
reference (seeded by the caller, add_noise.py:147-149)."""
import numpy as np


def gaussian(img, intensity):                       # add_noise.py:14-21
    sigma = intensity * 255 / 100
    return np.clip(img + np.random.normal(0, sigma, img.shape), 0, 255).astype(np.uint8)


def salt_and_pepper(img, intensity):                # add_noise.py:23-43 (incl. the i-1 upper bound and size*0.5 quirks)
    out = np.copy(img)
    amount = intensity / 100
    n = int(np.ceil(amount * img.size * 0.5))
    coords = [np.random.randint(0, i - 1, n) for i in img.shape]
    out[coords[0], coords[1], :] = 255
    coords = [np.random.randint(0, i - 1, n) for i in img.shape]
    out[coords[0], coords[1], :] = 0
    return out


def speckle(img, intensity):                        # add_noise.py:45-54
    g = np.random.randn(*img.shape)
    return np.clip(img + img * g * (intensity / 100), 0, 255).astype(np.uint8)


def poisson(img, intensity):                        # add_noise.py:56-65 (higher intensity -> larger scale -> less noise)
    scale = 10 + (intensity / 100) * 90
    return np.clip(np.random.poisson(img * scale / 255.0) * 255.0 / scale, 0, 255).astype(np.uint8)


def uniform(img, intensity):                        # add_noise.py:67-72
    r = intensity * 255 / 100
    return np.clip(img + np.random.uniform(-r / 2, r / 2, img.shape), 0, 255).astype(np.uint8)


MODELS = {"gaussian": gaussian, "salt_and_pepper": salt_and_pepper, "speckle": speckle, "poisson": poisson,
          "uniform": uniform}
