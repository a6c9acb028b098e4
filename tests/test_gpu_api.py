"""Drop-in surfaces added for the training dispatcher and the inference driver (8b B4), on the GPU
against the oracle's statement-for-statement wrappers of the reference call sites."""
import numpy as np
import pytest
import torch

from tests.parity import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wst():
    assert torch.cuda.is_available()
    import wst_b200
    return wst_b200


@pytest.fixture(scope="module")
def img():
    rng = np.random.default_rng(21)
    return (rng.integers(0, 256, (3, 128, 128)) / 255.0).astype(np.float32)


def ref_basic(rgb):
    f = np.zeros(6)                                   # inference.py:170-179
    for i in range(3):
        f[2 * i] = np.mean(rgb[i]); f[2 * i + 1] = np.std(rgb[i])
    return f


def test_training_dispatcher(wst, img):
    """extract_features(img, method), train_and_save_model.py:389-398."""
    from oracle import extract_wst_features_training
    from oracle.advanced_stats import extract_advanced_features
    adv = wst.extract_features(img, "advanced_stats")
    w = wst.extract_features(img, "wst")
    hyb = wst.extract_features(img, "hybrid")
    assert adv.shape == (54,) and adv.dtype == np.float64
    assert w.shape == (486,) and w.dtype == np.float32
    assert hyb.shape == (540,) and hyb.dtype == np.float64
    assert np.array_equal(hyb[:54], adv) and np.array_equal(hyb[54:], w.astype(np.float64))
    ref = extract_wst_features_training(img, precision="double", cache_filters=True).reshape(3, 2, 81)
    assert_parity(w.reshape(3, 2, 81)[:, 0], ref[:, 0], 2, 8, what="wst mean")
    assert_parity(w.reshape(3, 2, 81)[:, 1], ref[:, 1], 2, 8, what="wst std")
    ra = extract_advanced_features(img)
    ed = np.arange(54) % 18 == 17                                       # edge density: a fraction of H*W pixels
    assert np.abs(adv[~ed] - ra[~ed]).max() <= 2e-5 * np.maximum(np.abs(ra[~ed]), 1.0).max()
    assert np.abs(adv[ed] - ra[ed]).max() <= 1.5 / (128 * 128)
    assert len(wst.get_feature_names("hybrid")) == hyb.size


def test_inference_arms(wst, img):
    """ModelInference.extract_features, inference.py:272-287: wst -> basic(6) + interleaved WST = 492,
    hybrid -> advanced(54) + interleaved WST = 540, float64."""
    from oracle import extract_wst_features_inference
    ref_w = extract_wst_features_inference(img, cache_filters=True)
    m = wst.ModelInferenceFeatures()
    m.feature_method = "wst"
    f = m.extract_features(img)
    assert f.shape == (492,) and f.dtype == np.float64
    assert np.abs(f[:6] - ref_basic(img)).max() <= 2e-6
    got = f[6:].reshape(3, 81, 2)
    assert_parity(got[:, :, 0], ref_w.reshape(3, 81, 2)[:, :, 0], 2, 8, what="interleaved mean")
    assert_parity(got[:, :, 1], ref_w.reshape(3, 81, 2)[:, :, 1], 2, 8, what="interleaved std")
    m.feature_method = "hybrid"
    g = m.extract_features(img)
    assert g.shape == (540,) and np.array_equal(g[54:], f[6:]) and np.array_equal(g[:54], m.extract_advanced_features(img))
    m.feature_method = "advanced_stats"
    assert np.array_equal(m.extract_features(img), g[:54])
    assert np.array_equal(m.extract_wst_features(img, J=2, L=8), f[6:])
    assert np.array_equal(m.extract_basic_features(img), f[:6])
    # the block and interleaved layouts hold the same numbers
    blk = wst.extract_wst_features(img).reshape(3, 2, 81)
    assert np.array_equal(np.ascontiguousarray(blk.swapaxes(1, 2)).reshape(-1).astype(np.float64), f[6:])


def test_non_finite_and_short_inputs_are_refused(wst):
    bad = np.zeros((3, 32, 32), np.float32); bad[1, 3, 4] = np.nan
    with pytest.raises(ValueError, match="non-finite"):
        wst.extract_advanced_features(bad)
    with pytest.raises(IndexError):
        wst.extract_advanced_features(np.zeros((2, 32, 32), np.float32))
    four = np.random.default_rng(0).random((4, 32, 32)).astype(np.float32)
    assert wst.extract_advanced_features(four).shape == (54,)           # first three channels, like the reference


def test_forward_host_validates_out(wst):
    plan = wst.get_plan(32, 32, 2)
    x = np.zeros((2, 3, 32, 32), np.float32)
    for bad in (np.empty((2, 3, 2, 81), np.float64), np.empty((2, 3, 2, 80), np.float32),
                np.empty((2, 3, 2, 162), np.float32)[..., ::2], torch.empty((2, 3, 2, 81), device="cuda")):
        with pytest.raises(RuntimeError, match="out"):
            plan.forward_host(x, out=bad)
    out = np.empty((2, 3, 2, 81), np.float32)
    assert plan.forward_host(x, out=out) is out
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        plan.forward(torch.zeros((1, 32, 32, 3), dtype=torch.uint8))     # uint8 path validates where the pixels live
    if torch.cuda.device_count() > 1:
        with pytest.raises(RuntimeError, match="cuda:0"):
            plan.forward(torch.zeros((1, 32, 32, 3), dtype=torch.uint8, device="cuda:1"))


def test_torch_library_op(wst):
    rng = np.random.default_rng(2)
    x = torch.from_numpy((rng.integers(0, 256, (3, 3, 64, 64)) / 255.0).astype(np.float32)).cuda()
    plan = wst.get_plan(64, 64, 3)
    feats, maps = plan.forward(x, True, True)
    a = torch.ops.wst.scattering2d_features(x, 3, 8, 2, 0, False)
    assert torch.equal(a, feats.reshape(3, -1))
    b = torch.ops.wst.scattering2d_features(x, 3, 8, 2, 1, False)
    assert torch.equal(b, wst.to_interleaved(feats))
    assert torch.equal(torch.ops.wst.scattering2d_features(x, 3, 8, 2, 0, True), maps)
    assert torch.equal(torch.ops.wst.scattering2d_maps(x, 3, 8, 2), maps)
    torch.library.opcheck(torch.ops.wst.scattering2d_features.default, (x, 3, 8, 2, 0, False),
                          test_utils=("test_schema", "test_faketensor"))
    compiled = torch.compile(lambda t: torch.ops.wst.scattering2d_features(t, 3, 8, 2, 0, False) * 2.0, fullgraph=True,
                             backend="aot_eager")          # Dynamo traces through the op (fake kernel), no graph break
    assert torch.allclose(compiled(x), a * 2.0)


@pytest.mark.parametrize("M,J", [(32, 2), (64, 3), (128, 2)])
def test_ticket_scheduler_equals_fixed_stride(wst, M, J):
    """Many waves: the signals after each CTA's first one come from a device-wide ticket counter.  Every signal is
    processed exactly once and its numbers do not depend on which CTA ran it: bit-equal to the fixed-stride assignment
    (WST_STATIC_SCHED=1), call after call (the counter is re-zeroed per launch)."""
    import os
    plan = wst.get_plan(M, M, J)
    nsig = 20 * plan.grid + 3 if M <= 64 else 17 * plan.grid + 100
    g = torch.Generator(device="cuda").manual_seed(M + J)
    x = torch.randint(0, 256, (nsig, 1, M, M), device="cuda", generator=g, dtype=torch.int32).float().div_(255.0)
    f1 = plan.forward(x)[0].clone()
    f2 = plan.forward(x)[0].clone()
    os.environ["WST_STATIC_SCHED"] = "1"
    try:
        f_ref = plan.forward(x)[0].clone()
    finally:
        del os.environ["WST_STATIC_SCHED"]
    torch.cuda.synchronize()
    assert torch.equal(f1, f_ref) and torch.equal(f2, f_ref)


@pytest.mark.parametrize("M,J,tail", [(32, 2, 1), (64, 3, 7), (128, 4, 4), (128, 2, 70)])
def test_ragged_last_wave_runs_split(wst, M, J, tail):
    """nsig = 2 * grid + tail: the tail signals run as a second, split launch instead of leaving grid - tail CTAs idle
    for a whole signal time.  Same features and maps as one launch over everything, bit for bit — float32 planes, uint8
    HWC pixels (signal offset inside a patch) and the host-buffer path."""
    import os
    plan = wst.get_plan(M, M, J)
    nsig = 2 * plan.grid + tail            # (the split tail is kept for batches of up to 16 waves)
    rng = np.random.default_rng(M + J + tail)
    x = torch.from_numpy((rng.integers(0, 256, (nsig, 1, M, M)) / 255.0).astype(np.float32)).cuda()
    two = tail * 2 <= plan.grid
    assert plan.launch_count(nsig, 1) == (2 if two else 1)
    f_tail, m_tail = plan.forward(x, True, True)
    f_only = plan.forward(x)[0]
    f_host = plan.forward_host(x.cpu())
    os.environ["WST_NO_TAIL_SPLIT"] = "1"
    try:
        assert plan.launch_count(nsig, 1) == 1
        f_ref, m_ref = plan.forward(x, True, True)
    finally:
        del os.environ["WST_NO_TAIL_SPLIT"]
    torch.cuda.synchronize()
    assert torch.equal(m_tail, m_ref)
    assert torch.equal(f_tail, f_ref) and torch.equal(f_only, f_ref)
    assert np.array_equal(np.asarray(f_host).reshape(nsig, -1), f_ref.reshape(nsig, -1).cpu().numpy())
    if M <= 64:                                     # uint8 HWC: 3 signals per patch, the tail starts inside the batch
        B = (nsig + 2) // 3
        u8 = torch.from_numpy(rng.integers(0, 256, (B, M, M, 3), dtype=np.uint8)).cuda()
        f_u8 = plan.forward(u8)[0]
        os.environ["WST_NO_TAIL_SPLIT"] = "1"
        try:
            f_u8_ref = plan.forward(u8)[0]
        finally:
            del os.environ["WST_NO_TAIL_SPLIT"]
        assert torch.equal(f_u8, f_u8_ref)


@pytest.mark.parametrize("M,J,B", [(128, 2, 1), (128, 4, 2), (64, 3, 5), (32, 2, 1)])
def test_small_batches_split_signals_across_ctas(wst, M, J, B):
    """The reference calls the extractor one image at a time (train_and_save_model.py:486-488): with fewer signals than
    SMs the first-order groups of a signal are shared among several CTAs and the last one pools.  Same numbers as the
    one-CTA-per-signal path, bit for bit (the arithmetic of every group is unchanged)."""
    import os
    rng = np.random.default_rng(M + J + B)
    x = torch.from_numpy((rng.integers(0, 256, (B, 3, M, M)) / 255.0).astype(np.float32)).cuda()
    plan = wst.get_plan(M, M, J)
    f_split, m_split = plan.forward(x, True, True)
    f_only = plan.forward(x)[0]                       # maps in the per-signal scratch
    os.environ["WST_NO_SPLIT"] = "1"
    try:
        f_ref, m_ref = plan.forward(x, True, True)
    finally:
        del os.environ["WST_NO_SPLIT"]
    torch.cuda.synchronize()
    assert torch.equal(m_split, m_ref)
    assert torch.equal(f_split, f_ref) and torch.equal(f_only, f_ref)
    one = wst.extract_wst_features(x[0].cpu().numpy(), J=J)                 # host path, one image
    assert np.array_equal(one, f_ref[0].reshape(-1).cpu().numpy())
