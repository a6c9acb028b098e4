"""Oracle vs closed-form known answers and vs the committed golden vectors.
kymatio is absent, so these are the pins the oracle has (PARITY UNPINNED — oracle/__init__.py)."""
import os

import numpy as np
import pytest

from tests import patterns
from oracle import (Scattering2D, compute_padding, num_coefficients, filter_bank,
                    extract_wst_features_training, extract_wst_features_inference,
                    extract_wst_features_visualization, compute_scattering_coefficients)

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("M,J,L,mo,K,Np,h", [
    (32, 2, 8, 2, 81, 40, 8), (64, 3, 8, 2, 217, 80, 8), (128, 4, 8, 2, 417, 160, 8),
    (128, 2, 8, 2, 81, 136, 32), (32, 3, 6, 2, 127, 48, 4), (512, 5, 8, 2, 681, 576, 16),
    (32, 2, 8, 1, 17, 40, 8),
])
def test_geometry(M, J, L, mo, K, Np, h):
    assert num_coefficients(J, L, mo) == K
    assert compute_padding(M, M, J) == (Np, Np)
    assert Np // 2 ** J - 2 == h


def test_shapes_and_dtype():
    S = Scattering2D(J=2, shape=(32, 32), cache_filters=True)
    x = np.random.default_rng(0).random((3, 2, 32, 32), dtype=np.float32)
    y = S(x)
    assert y.shape == (3, 2, 81, 8, 8) and y.dtype == np.float32
    S1 = Scattering2D(J=2, shape=(32, 32), max_order=1, cache_filters=True)
    assert S1(x[0, 0]).shape == (17, 8, 8)
    np.testing.assert_array_equal(S1(x[0, 0]), y[0, 0, :17])            # order<=1 prefix is identical
    np.testing.assert_array_equal(S(x[1:2, 1])[0], y[1, 1])              # batch equivalence (A.4 item 7)


def test_errors():
    with pytest.raises(RuntimeError, match="smallest dimension"):
        Scattering2D(J=6, shape=(32, 32))
    S = Scattering2D(J=2, shape=(32, 32), cache_filters=True)
    with pytest.raises(TypeError):
        S([[0.0]])
    with pytest.raises(RuntimeError, match="spatial size"):
        S(np.zeros((16, 16), np.float32))


def test_constant_image():
    """A.4 item 2: S0 == c * pi / 3.1415 everywhere, higher orders ~ 0."""
    S = Scattering2D(J=2, shape=(32, 32), cache_filters=True)
    c = 0.37
    y = S(np.full((32, 32), c, np.float32))
    np.testing.assert_allclose(y[0], c * np.pi / 3.1415, rtol=2e-6)
    assert np.abs(y[1:]).max() < 1e-6 * c


def test_filter_bank_facts():
    fb = filter_bank(40, 40, 2, 8)
    assert len(fb["psi"]) == 16 and len(fb["phi"]["levels"]) == 2
    assert fb["phi"]["levels"][0].dtype == np.float32 and fb["phi"]["levels"][1].shape == (20, 20)
    assert abs(fb["phi"]["levels"][0][0, 0] - np.pi / 3.1415) < 1e-6
    for p in fb["psi"]:
        assert abs(p["levels"][0][0, 0]) < 1e-6                          # zero-mean wavelets
        assert len(p["levels"]) == min(p["j"] + 1, 1)
    # Littlewood-Paley sum is bounded (energy is not amplified)
    lp = fb["phi"]["levels"][0].astype(np.float64) ** 2 + 0.5 * sum(
        p["levels"][0].astype(np.float64) ** 2 + np.roll(p["levels"][0][::-1, ::-1], 1, (0, 1)).astype(np.float64) ** 2
        for p in fb["psi"])
    assert lp.max() < 2.1


def test_homogeneity_and_ordering():
    S = Scattering2D(J=3, shape=(64, 64), cache_filters=True)
    x = np.random.default_rng(1).random((64, 64), dtype=np.float32)
    y1, y2 = S(x), S(2.0 * x)
    np.testing.assert_allclose(y2, 2.0 * y1, rtol=1e-5, atol=1e-7)       # positive homogeneity
    # orientation selectivity: a plane wave at (scale j, angle t) peaks at order-1 index 1 + j*L + t
    L = 8
    r, c = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
    for j in (0, 1):
        xi = 3.0 / 4.0 * np.pi / 2 ** j
        for t in range(L):
            theta = (int(L - L / 2 - 1) - t) * np.pi / L
            g = np.cos(xi * (r * np.cos(theta) + c * np.sin(theta))).astype(np.float32)
            e = S(g)[1 + j * L:1 + (j + 1) * L].mean(axis=(-2, -1))
            assert int(np.argmax(e)) == t, (j, t, e)


def test_reference_wrappers_layouts():
    rgb = np.random.default_rng(2).random((3, 32, 32), dtype=np.float32)
    tr = extract_wst_features_training(rgb, cache_filters=True)
    inf = extract_wst_features_inference(rgb, cache_filters=True)
    assert tr.shape == (486,) and tr.dtype == np.float32
    assert inf.shape == (486,) and inf.dtype == np.float64
    K = 81
    for c in range(3):                                                   # A.4 item 6
        np.testing.assert_allclose(inf[c * 2 * K:(c + 1) * 2 * K:2], tr[c * 2 * K:c * 2 * K + K], rtol=1e-6)
        np.testing.assert_allclose(inf[c * 2 * K + 1:(c + 1) * 2 * K:2], tr[c * 2 * K + K:(c + 1) * 2 * K], rtol=1e-5)
    f, maps = extract_wst_features_visualization(patterns.checkerboard(32), cache_filters=True)
    assert f.shape == (162,) and maps.shape == (81, 8, 8) and maps.dtype == np.float64
    neg = compute_scattering_coefficients(rgb[0], cache_filters=True)
    assert neg.shape == (127, 4, 4) and (neg[0] < 0).all()


@pytest.mark.parametrize("tag,J,L", [("cfg1_32_J2", 2, 8), ("cfg2_64_J3", 3, 8), ("compare_32_J3_L6", 3, 6)])
def test_golden_vectors(tag, J, L):
    g = np.load(os.path.join(GOLD, tag + ".npz"))
    x = g["x"]
    M = x.shape[-1]
    c64 = Scattering2D(J=J, shape=(M, M), L=L, precision="double", cache_filters=True)(x)
    np.testing.assert_allclose(c64.mean(axis=(-2, -1)), g["mean64"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(c64.std(axis=(-2, -1)), g["std64"], rtol=1e-9, atol=1e-12)
    c32 = Scattering2D(J=J, shape=(M, M), L=L, cache_filters=True)(x)
    np.testing.assert_allclose(c32.mean(axis=(-2, -1)), g["mean32"], rtol=1e-5, atol=1e-7)
    if "maps64" in g.files:
        np.testing.assert_allclose(c64, g["maps64"], rtol=1e-10, atol=1e-12)


def test_scaler_soft_pin_envelope():
    """A.4 item 8: magnitudes of natural-like (1/f) patches at 128x128 J=2 fall in the envelope the
    reference's shipped scalers record (S0 ~0.2-0.6, order 1 ~1e-3..3e-2, order 2 an order below)."""
    rng = np.random.default_rng(3)
    f = np.fft.fftfreq(128)
    amp = 1.0 / np.maximum(np.hypot(*np.meshgrid(f, f, indexing="ij")), 1.0 / 128)
    img = np.real(np.fft.ifft2(amp * np.exp(2j * np.pi * rng.random((128, 128)))))
    img = ((img - img.min()) / (img.max() - img.min())).astype(np.float32)
    y = Scattering2D(J=2, shape=(128, 128), cache_filters=True)(img).mean(axis=(-2, -1))
    assert 0.2 < y[0] < 0.8
    assert 1e-3 < y[1:17].mean() < 5e-2
    assert y[17:].mean() < y[1:17].mean() / 3
