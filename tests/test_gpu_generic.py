"""The shape-generic DFT-matrix engine (csrc/wst_generic.cu) on the GPU: rectangular shapes and padded sizes without
a compiled cascade (the reference builds Scattering2D from the image's own shape, train_and_save_model.py:355-359),
L > 8, both arithmetic engines (fp32 SIMT, 3xTF32 tensor cores), against the float64 oracle per order, and against
the fused FFT cascade on a shape both can run."""
import numpy as np
import pytest
import torch

from tests.parity import assert_parity

pytestmark = pytest.mark.gpu


def oracle64(H, W, J, L, mo=2):
    from oracle import Scattering2D
    return Scattering2D(J=J, shape=(H, W), L=L, max_order=mo, precision="double", cache_filters=True)


@pytest.fixture(scope="module")
def wst():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import wst_b200
    return wst_b200


SHAPES = [(100, 100, 2, 8), (120, 128, 2, 8), (96, 160, 3, 8), (50, 70, 1, 8), (33, 47, 2, 6), (64, 64, 2, 12),
          (212, 60, 2, 4)]      # 212 -> padded 220 = 4 * 5 * 11; 33 -> 40, 47 -> 56


@pytest.mark.parametrize("engine", ["gemm", "gemm_tf32x3"])
@pytest.mark.parametrize("H,W,J,L", SHAPES)
def test_generic_engine_vs_oracle(wst, H, W, J, L, engine):
    rng = np.random.default_rng(H * 1000 + W)
    x = (rng.integers(0, 256, (2, 3, H, W)) / 255.0).astype(np.float32)
    plan = wst.get_plan(H, W, J, L, engine=engine)
    assert plan.engine == engine
    Hp, Wp = wst.compute_padding(H, W, J)
    assert (plan.Hp, plan.Wp, plan.h, plan.w) == (Hp, Wp, Hp // 2 ** J - 2, Wp // 2 ** J - 2)
    feats, maps = plan.forward(torch.from_numpy(x).cuda(), True, True)
    torch.cuda.synchronize()
    ref = oracle64(H, W, J, L)(x)
    K = ref.shape[-3]
    assert tuple(maps.shape) == ref.shape
    rep = assert_parity(maps.cpu().numpy().reshape(6, K, -1), ref.reshape(6, K, -1), J, L, what="%s maps" % engine)
    f = feats.cpu().numpy().reshape(6, 2, K)
    assert_parity(f[:, 0], ref.mean(axis=(-2, -1)).reshape(6, K), J, L, what="%s mean" % engine)
    assert_parity(f[:, 1], ref.std(axis=(-2, -1)).reshape(6, K), J, L, what="%s std" % engine)
    print("generic %s %dx%d J=%d L=%d per order: %s" % (engine, H, W, J, L, rep))


def test_auto_engine_picks_generic_for_uncompiled_shapes(wst):
    assert wst.get_plan(100, 100, 2).engine == "gemm"
    assert wst.get_plan(120, 128, 2).engine == "gemm"
    assert wst.get_plan(128, 128, 2).engine == "fft"
    with pytest.raises(NotImplementedError):
        wst.Plan(100, 100, 2, engine="fft")


@pytest.mark.parametrize("engine", ["gemm", "gemm_tf32x3"])
def test_generic_engine_equals_fused_cascade(wst, engine):
    """Same shape through both engines (64x64 J=3, BASELINE configs[1]): the two agree far inside the tolerance."""
    rng = np.random.default_rng(9)
    x = torch.from_numpy((rng.integers(0, 256, (4, 3, 64, 64)) / 255.0).astype(np.float32)).cuda()
    a = wst.get_plan(64, 64, 3, 8, engine="fft").forward(x, True, True)
    b = wst.get_plan(64, 64, 3, 8, engine=engine).forward(x, True, True)
    K = a[1].shape[2]
    assert_parity(b[1].cpu().numpy().reshape(12, K, -1), a[1].cpu().numpy().reshape(12, K, -1), 3, 8, tol=2e-5)
    assert_parity(b[0].cpu().numpy()[:, :, 0].reshape(12, K), a[0].cpu().numpy()[:, :, 0].reshape(12, K), 3, 8, tol=2e-5)


def test_generic_engine_entry_points(wst):
    """uint8 HWC ingest, max_order=1, chunking over several workspace chunks, the host path and the drop-in extractor
    all work on a rectangular shape."""
    import os
    H, W, J, L = 72, 100, 2, 8
    rng = np.random.default_rng(1)
    u8 = rng.integers(0, 256, (5, H, W, 3), dtype=np.uint8)
    chw = np.ascontiguousarray(np.transpose(u8.astype(np.float32) / 255.0, (0, 3, 1, 2)))
    plan = wst.get_plan(H, W, J, L)
    f_u8 = plan.forward(torch.from_numpy(u8).cuda())[0]
    f_f32 = plan.forward(torch.from_numpy(chw).cuda())[0]
    assert torch.equal(f_u8, f_f32)
    os.environ["WST_GENERIC_WORKSPACE_MB"] = "64"           # forces several chunks for 15 signals
    try:
        f_chunked = plan.forward(torch.from_numpy(chw).cuda())[0]
    finally:
        del os.environ["WST_GENERIC_WORKSPACE_MB"]
    assert torch.equal(f_chunked, f_f32)
    host = plan.forward_host(chw)
    assert np.array_equal(host, f_f32.cpu().numpy())
    one = wst.extract_wst_features(chw[0])
    assert np.array_equal(one, host[0].reshape(-1))
    from oracle import extract_wst_features_training
    ref = extract_wst_features_training(chw[0], J=J, L=L, precision="double", cache_filters=True)
    K = plan.K
    assert_parity(one.reshape(3, 2, K)[:, 0], ref.reshape(3, 2, K)[:, 0], J, L)
    # first order only
    p1 = wst.get_plan(H, W, J, L, max_order=1)
    m1 = p1.forward(torch.from_numpy(chw).cuda(), False, True)[1].cpu().numpy()
    ref1 = oracle64(H, W, J, L, 1)(chw)
    assert m1.shape == ref1.shape
    assert_parity(m1.reshape(15, p1.K, -1), ref1.reshape(15, p1.K, -1), J, L, max_order=1)
    assert plan.launch_count(5, 3) > 1


def test_frontends_and_extractors_on_a_rectangular_image(wst):
    """The reference's call pattern on an image whose shape has no compiled cascade: Scattering2D(J, L, shape=(H, W))
    built from the image (train_and_save_model.py:355-359), numpy and torch frontends, training and inference layouts."""
    import wst_b200.numpy, wst_b200.torch
    from oracle import extract_wst_features_training, extract_wst_features_inference
    H, W = 90, 120
    rng = np.random.default_rng(4)
    img = (rng.integers(0, 256, (3, H, W)) / 255.0).astype(np.float32)
    S = wst_b200.numpy.Scattering2D(J=2, L=8, shape=(H, W))
    out = S(img[0])
    ref = oracle64(H, W, 2, 8)(img[0])
    assert out.shape == ref.shape == (81, (wst.compute_padding(H, W, 2)[0] >> 2) - 2, (wst.compute_padding(H, W, 2)[1] >> 2) - 2)
    assert out.dtype == np.float32
    assert_parity(out[None].reshape(1, 81, -1), ref[None].reshape(1, 81, -1), 2, 8)
    St = wst_b200.torch.Scattering2D(J=2, shape=(H, W), L=8)
    with torch.no_grad():
        t = St(torch.from_numpy(img[1]).unsqueeze(0).unsqueeze(0).contiguous())        # inference.py:250-254
    assert tuple(t.shape) == (1, 1) + ref.shape
    assert_parity(t[0].numpy().reshape(1, 81, -1), oracle64(H, W, 2, 8)(img[1])[None].reshape(1, 81, -1), 2, 8)
    f_train = wst.extract_wst_features(img)
    f_inf = wst.extract_wst_features_interleaved(img)
    r_train = extract_wst_features_training(img, precision="double", cache_filters=True)
    r_inf = extract_wst_features_inference(img.astype(np.float64), cache_filters=True)
    assert f_train.shape == r_train.shape == (486,) and f_inf.shape == r_inf.shape == (486,)
    assert_parity(f_train.reshape(3, 2, 81)[:, 0], r_train.reshape(3, 2, 81)[:, 0], 2, 8)
    assert np.array_equal(np.ascontiguousarray(f_train.reshape(3, 2, 81).swapaxes(1, 2)).reshape(-1).astype(np.float64), f_inf)
