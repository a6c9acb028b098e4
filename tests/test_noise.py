"""Noise models of the robustness sweep: the oracle against fixtures produced by the REAL
reference module (tests/golden/make_golden_noise.py imports src/preprocessing/add_noise.py), the CUDA kernels
against both (bit for bit when handed the reference's draws), and the device generator statistically."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "noise.npz")
CASES = [("gaussian", 30), ("gaussian", 50), ("poisson", 40), ("poisson", 60), ("salt_and_pepper", 5),
         ("salt_and_pepper", 15), ("salt_and_pepper", 25), ("speckle", 15), ("speckle", 35), ("speckle", 55),
         ("uniform", 10), ("uniform", 25), ("uniform", 40)]


def _seeded_draws(kind, intensity, imgs):
    """The reference's draws for a batch noised image after image from np.random.seed(42) (add_noise.py:147-149)."""
    from oracle import add_noise as ora
    np.random.seed(42)
    return [ora.draw(kind, im, intensity) for im in imgs]


@pytest.mark.parametrize("kind,intensity", CASES)
def test_oracle_is_the_reference_bit_for_bit(kind, intensity):
    from oracle import add_noise as ora
    g = np.load(GOLD)
    np.random.seed(42)
    got = np.stack([ora.add_noise(kind, im, intensity) for im in g["img"]])
    np.testing.assert_array_equal(got, g["%s_%d" % (kind, intensity)])
    # the draw / apply split is the same function
    d = _seeded_draws(kind, intensity, g["img"])
    got2 = np.stack([ora.apply(kind, im, intensity, di) for im, di in zip(g["img"], d)])
    np.testing.assert_array_equal(got2, got)


def test_tests_noise_helper_is_the_oracle():
    """tests/noise.py (used by the config-4 sweep) and the pinned oracle are the same models."""
    from oracle import add_noise as ora
    from tests import noise
    g = np.load(GOLD)
    for kind, intensity in CASES:
        np.random.seed(7)
        a = noise.MODELS[kind](g["img"][1], intensity)
        np.random.seed(7)
        b = ora.add_noise(kind, g["img"][1], intensity)
        np.testing.assert_array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,intensity", CASES)
def test_kernel_with_reference_draws_is_bit_exact(kind, intensity):
    import wst_b200
    g = np.load(GOLD)
    imgs = g["img"]
    d = _seeded_draws(kind, intensity, imgs)
    if kind == "salt_and_pepper":
        draws = np.stack([np.stack([salt, pep]) for salt, pep in d]).astype(np.int64)        # [B, 2, 2, n]
    else:
        draws = np.stack(d).astype(np.int64 if kind == "poisson" else np.float64)
    got = wst_b200.add_noise(torch.from_numpy(imgs).cuda(), kind, intensity,
                             draws=torch.from_numpy(np.ascontiguousarray(draws)).cuda())
    np.testing.assert_array_equal(got.cpu().numpy(), g["%s_%d" % (kind, intensity)])


@pytest.mark.gpu
def test_device_generator_statistics():
    """Same distributions as the reference's numpy calls (the streams cannot match): checked on a mid-grey batch
    where nothing clips, against the closed-form moments of each model after the truncating uint8 cast."""
    import wst_b200
    B, H, W, C = 8, 128, 128, 3
    n = B * H * W * C
    base = torch.full((B, H, W, C), 128, dtype=torch.uint8, device="cuda")
    tol = 6.0 / np.sqrt(n)                                             # six standard errors of a unit-variance mean

    y = wst_b200.add_noise(base, "gaussian", 10, seed=1).double().cpu().numpy()     # sigma = 25.5, floor() shifts by -0.5
    assert abs(y.mean() - 127.5) <= 25.5 * tol + 0.02 and abs(y.std() - 25.5) <= 0.15
    z = (y - 127.5) / 25.5
    assert abs((z ** 3).mean()) <= 0.03 and abs((z ** 4).mean() - 3.0) <= 0.08

    y = wst_b200.add_noise(base, "uniform", 40, seed=2).double().cpu().numpy()      # range 102 -> +-51
    assert abs(y.mean() - 127.5) <= 30 * tol + 0.02 and abs(y.std() - 102 / np.sqrt(12)) <= 0.1
    assert y.min() >= 128 - 51 - 1 and y.max() <= 128 + 51

    y = wst_b200.add_noise(base, "speckle", 15, seed=3).double().cpu().numpy()      # 128 * (1 + 0.15 g)
    assert abs(y.mean() - 127.5) <= 19.2 * tol + 0.02 and abs(y.std() - 19.2) <= 0.12

    for intensity, level in ((40, 128), (100, 128), (0, 60)):           # lambda = level * scale / 255: both generator regimes
        scale = 10 + intensity / 100 * 90
        lam = level * scale / 255.0
        src = torch.full((B, H, W, C), level, dtype=torch.uint8, device="cuda")
        y = wst_b200.add_noise(src, "poisson", intensity, seed=4 + intensity).double().cpu().numpy()
        k = np.rint(y * scale / 255.0 + 0.49)                           # invert the truncating cast: k*255/scale -> k
        assert abs(k.mean() - lam) <= np.sqrt(lam) * tol + 0.01, (intensity, k.mean(), lam)
        assert abs(k.var() - lam) <= 0.02 * lam, (intensity, k.var(), lam)
    dark = torch.full((2, 64, 64, 3), 3, dtype=torch.uint8, device="cuda")          # lambda = 0.54: the small-lambda branch
    y = wst_b200.add_noise(dark, "poisson", 40, seed=9).double().cpu().numpy()
    k = np.rint(y * 46 / 255.0 + 0.49)
    assert abs(k.mean() - 3 * 46 / 255.0) <= 0.02

    # different seeds differ, same seed repeats, batch elements differ
    a = wst_b200.add_noise(base, "gaussian", 10, seed=5)
    assert torch.equal(a, wst_b200.add_noise(base, "gaussian", 10, seed=5))
    assert not torch.equal(a, wst_b200.add_noise(base, "gaussian", 10, seed=6))
    assert not torch.equal(a[0], a[1])


@pytest.mark.gpu
def test_device_salt_and_pepper_conventions():
    """add_noise.py:23-43: ceil(amount * H*W*C * 0.5) coordinates per colour, rows in [0, H-1), columns in [0, W-1),
    every channel of a hit pixel set, pepper over salt."""
    import wst_b200
    B, H, W, C = 4, 64, 48, 3
    base = torch.full((B, H, W, C), 100, dtype=torch.uint8, device="cuda")
    y = wst_b200.add_noise(base, "salt_and_pepper", 15, seed=11).cpu().numpy()
    assert set(np.unique(y)) <= {0, 100, 255}
    assert (y[:, H - 1] == 100).all() and (y[:, :, W - 1] == 100).all()                  # never touched
    assert (y[..., 0] == y[..., 1]).all() and (y[..., 0] == y[..., 2]).all()           # all channels together
    ncoord = int(np.ceil(0.15 * H * W * C * 0.5))
    cells = (H - 1) * (W - 1)
    p_pepper = 1 - (1 - 1 / cells) ** ncoord                                             # pixel hit by >= 1 pepper draw
    p_salt = (1 - p_pepper) * p_pepper                                                   # salt hit and no pepper after
    got_pepper = (y[:, :H - 1, :W - 1, 0] == 0).mean()
    got_salt = (y[:, :H - 1, :W - 1, 0] == 255).mean()
    assert abs(got_pepper - p_pepper) <= 0.02 and abs(got_salt - p_salt) <= 0.02
    with pytest.raises(ValueError):
        wst_b200.add_noise(torch.zeros(1, 8, 8, 1, dtype=torch.uint8, device="cuda"), "salt_and_pepper", 5)
    with pytest.raises(ValueError):
        wst_b200.add_noise(base, "pink", 5)
    with pytest.raises(ValueError):
        wst_b200.add_noise(base, "gaussian", 101)


@pytest.mark.gpu
def test_noise_feeds_the_uint8_ingest_and_dropins():
    """Noise on the device -> uint8 ingest of the WST plan (BASELINE configs[3] end to end on the GPU) equals noising,
    downloading and re-uploading; the per-image drop-ins return uint8 HWC numpy arrays like the reference's."""
    import wst_b200
    rng = np.random.default_rng(5)
    u8 = torch.from_numpy(rng.integers(0, 256, (6, 64, 64, 3), dtype=np.uint8)).cuda()
    noisy = wst_b200.add_noise(u8, "speckle", 35, seed=42)
    plan = wst_b200.get_plan(64, 64, 3, 8)
    f1, _ = plan.forward(noisy)
    f2, _ = plan.forward(torch.from_numpy(noisy.cpu().numpy()).cuda())
    assert torch.equal(f1, f2) and tuple(f1.shape)[0] == 6
    np.random.seed(42)
    img = u8[0].cpu().numpy()
    for fn in (wst_b200.add_gaussian_noise, wst_b200.add_salt_and_pepper_noise, wst_b200.add_speckle_noise,
               wst_b200.add_poisson_noise, wst_b200.add_uniform_noise):
        out = fn(img, 25)
        assert out.dtype == np.uint8 and out.shape == img.shape and not np.array_equal(out, img)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 2, 2, 2), (3, 5, 7, 3), (2, 33, 17, 4), (0, 8, 8, 3)])
@pytest.mark.parametrize("kind", ["gaussian", "salt_and_pepper", "speckle", "poisson", "uniform"])
def test_kernel_ragged_shapes_with_reference_draws(kind, shape):
    """Odd sizes, more than three channels, the smallest image the reference's salt-and-pepper accepts, an empty batch."""
    import wst_b200
    from oracle import add_noise as ora
    rng = np.random.default_rng(sum(shape))
    imgs = rng.integers(0, 256, shape, dtype=np.uint8)
    if shape[0] == 0:
        out = wst_b200.add_noise(torch.from_numpy(imgs).cuda(), kind, 25, seed=1)
        assert tuple(out.shape) == shape
        return
    np.random.seed(42)
    d = [ora.draw(kind, im, 25) for im in imgs]
    ref = np.stack([ora.apply(kind, im, 25, di) for im, di in zip(imgs, d)])
    if kind == "salt_and_pepper":
        draws = np.stack([np.stack([s, p]) for s, p in d]).astype(np.int64)
    else:
        draws = np.stack(d).astype(np.int64 if kind == "poisson" else np.float64)
    got = wst_b200.add_noise(torch.from_numpy(imgs).cuda(), kind, 25, draws=torch.from_numpy(np.ascontiguousarray(draws)).cuda())
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    rnd = wst_b200.add_noise(torch.from_numpy(imgs).cuda(), kind, 25, seed=3)          # device generator: shape and range only
    assert tuple(rnd.shape) == shape and rnd.dtype == torch.uint8
