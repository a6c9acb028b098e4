"""Oracle for the five noise models of the robustness sweep.  TEST INFRASTRUCTURE ONLY.

Restates src/preprocessing/add_noise.py:14-72.  PINNED: the reference module itself imports in the build
container (numpy + PIL only); tests/golden/make_golden_noise.py runs it on seeded inputs and
tests/test_noise.py holds this restatement to those outputs bit for bit (same numpy global RNG stream, seeded
like add_noise.py:147-149).

Each model is split in two so that the arithmetic can be checked separately from the random stream:
`draw_*` consumes numpy's global RNG exactly as the reference does and returns the draws; `apply_*` is the
deterministic part (what the CUDA kernel must reproduce bit for bit when it is handed the same draws).
"""
import numpy as np

KINDS = ["gaussian", "salt_and_pepper", "speckle", "poisson", "uniform"]      # add_noise.py:123


def poisson_scale(intensity):
    return 10 + (intensity / 100) * 90                        # add_noise.py:60


def salt_pepper_count(shape, intensity):
    """add_noise.py:33,38: ceil(amount * image.size * 0.5) -- image.size counts the channels too."""
    amount = intensity / 100
    return int(np.ceil(amount * int(np.prod(shape)) * 0.5)), int(np.ceil(amount * int(np.prod(shape)) * (1. - 0.5)))


def draw(kind, img, intensity):
    """The random draws of one call, taken from numpy's global RNG in the reference's order."""
    row, col, ch = img.shape
    if kind == "gaussian":
        return np.random.normal(0, intensity * 255 / 100, (row, col, ch))                       # :18-19
    if kind == "speckle":
        return np.random.randn(row, col, ch)                                                    # :48
    if kind == "uniform":
        r = intensity * 255 / 100
        return np.random.uniform(-r / 2, r / 2, (row, col, ch))                                 # :70-71
    if kind == "poisson":
        return np.random.poisson(img * poisson_scale(intensity) / 255.0)                        # :61-64
    if kind == "salt_and_pepper":
        ns, npep = salt_pepper_count(img.shape, intensity)
        salt = [np.random.randint(0, i - 1, ns) for i in img.shape]                             # :34 (3 arrays drawn)
        pep = [np.random.randint(0, i - 1, npep) for i in img.shape]                            # :39
        return np.stack([salt[0], salt[1]]), np.stack([pep[0], pep[1]])                         # channel coords unused
    raise ValueError("Unknown noise type: %s" % kind)                                           # :92


def apply(kind, img, intensity, d):
    """Deterministic part: uint8 [H, W, C] + draws -> uint8 [H, W, C]."""
    if kind in ("gaussian", "uniform"):
        return np.clip(img + d, 0, 255).astype(np.uint8)                                        # :20-21, :72-73
    if kind == "speckle":
        return np.clip(img + img * d * (intensity / 100), 0, 255).astype(np.uint8)              # :52-54
    if kind == "poisson":
        return np.clip(d * 255.0 / poisson_scale(intensity), 0, 255).astype(np.uint8)           # :64-65
    if kind == "salt_and_pepper":
        out = np.copy(img)
        salt, pep = d
        out[salt[0], salt[1], :] = 255                                                          # :35
        out[pep[0], pep[1], :] = 0                                                              # :40
        return out
    raise ValueError("Unknown noise type: %s" % kind)


def add_noise(kind, img, intensity):
    """process_image's dispatch (add_noise.py:82-92) for an in-memory uint8 HWC image."""
    return apply(kind, img, intensity, draw(kind, img, intensity))
