"""The seven deterministic synthetic images of the reference (src/visualization/visualize_features.py:48-120),
restated (vectorised) so tests and golden fixtures have reference-defined inputs at any size.
tests/golden/make_golden.py checks them against the reference's own functions when /root/reference exists."""
import numpy as np


def gradient_horizontal(size=128):          # visualize_features.py:48-51
    return np.tile(np.linspace(0, 1, size)[None, :], (size, 1))


def gradient_vertical(size=128):            # :53-56
    return np.tile(np.linspace(0, 1, size)[:, None], (1, size))


def checkerboard(size=128, squares=8):      # :58-67
    q = size // squares
    img = np.zeros((size, size))
    i, j = np.meshgrid(np.arange(squares * q) // q, np.arange(squares * q) // q, indexing="ij")
    img[: squares * q, : squares * q] = ((i + j) % 2 == 0).astype(np.float64)
    return img


def circles(size=128, num_circles=5):       # :69-81
    i, j = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    dist = np.sqrt((i - size / 2) ** 2 + (j - size / 2) ** 2)
    return np.sin(dist / (size / 2) * num_circles * np.pi) * 0.5 + 0.5


def texture(size=128, seed=42):             # :83-87
    np.random.seed(seed)
    return np.random.rand(size, size)


def vertical_texture(size=128, seed=42, frequency=8):   # :89-114
    np.random.seed(seed)
    x = np.linspace(0, frequency * 2 * np.pi, size)
    pattern = (np.tile(np.sin(x)[None, :], (size, 1)) + 1) / 2
    noise = np.random.rand(size, size) * 0.3
    return np.clip(pattern * 0.7 + noise, 0, 1)


def edge(size=128, border_width=20):        # :116-120
    img = np.zeros((size, size))
    img[border_width:size - border_width, border_width:size - border_width] = 1.0
    return img


GENERATORS = {
    "gradient_horizontal": gradient_horizontal, "gradient_vertical": gradient_vertical,
    "checkerboard": checkerboard, "circles": circles, "texture": texture,
    "vertical_texture": vertical_texture, "edge": edge,
}
REFERENCE_NAMES = {   # name in visualize_features.py
    "gradient_horizontal": "generate_gradient_horizontal", "gradient_vertical": "generate_gradient_vertical",
    "checkerboard": "generate_checkerboard", "circles": "generate_circles", "texture": "generate_texture",
    "vertical_texture": "generate_vertical_texture", "edge": "generate_edge",
}


def all_patterns(size):
    """[7, size, size] float64, in GENERATORS order; the edge border scales with size like 20/128."""
    out = []
    for name, fn in GENERATORS.items():
        if name == "edge":
            out.append(fn(size, max(1, (20 * size) // 128)))
        else:
            out.append(fn(size))
    return np.stack(out)
