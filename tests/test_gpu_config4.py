"""BASELINE configs[3]: noise-robustness sweep at 64x64 J=3 feeding the reference's Random-Forest trainer.

Synthetic three-class texture patches are noised with the reference's models (tests/noise.py), features are
extracted on the GPU from the uint8 pixels (wst2d_forward_u8 = load_rgb_image + extract_wst_features) and on the CPU
by the oracle, and both feature matrices go through the reference's pipeline — select_features_kbest and
train_final_model of train_and_save_model.py:147-198, imported UNCHANGED where /root/reference exists (the build
container) and otherwise through the restatement that tests/test_rf_pipeline.py holds to those functions bit for bit
and to tests/golden/rf_pipeline.npz (the reference functions' own outputs on oracle features).  The consumer must not
be able to tell the two feature sources apart."""
import os

import numpy as np
import pytest
import torch

from tests import noise, rf_pipeline

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "rf_pipeline.npz"))


def gpu_features(u8, J, L):
    import wst_b200
    plan = wst_b200.get_plan(u8.shape[1], u8.shape[2], J, L)
    feats, _ = plan.forward(torch.from_numpy(u8).cuda())                 # uint8 HWC ingest
    return wst_b200.to_block(feats).cpu().numpy()


@pytest.mark.parametrize("tag,model,intensity", [("clean", None, 0), ("gaussian30", "gaussian", 30)])
def test_against_reference_trainer_golden(tag, model, intensity):
    """GPU features -> pipeline == the record the reference's own functions produced from oracle features."""
    M, J, L = 64, 3, 8
    u8, y = rf_pipeline.make_patches(10, M, np.random.default_rng(7))
    if model:
        np.random.seed(42)                                              # add_noise.py:147-149
        u8 = np.stack([noise.MODELS[model](im, intensity) for im in u8])
    px = u8.astype(np.int64)
    assert int(px.sum()) == int(GOLD[tag + "_u8_sum"][0]), "synthetic patches drifted from the golden record's"
    Xg, Xo = gpu_features(u8, J, L), GOLD[tag + "_X"]
    assert Xg.shape == Xo.shape == (30, 3 * 2 * 217)
    tau = 1e-3 * np.abs(Xo).max(axis=1, keepdims=True)
    assert float((np.abs(Xg - Xo) / np.maximum(np.abs(Xo), tau)).max()) <= 1e-4
    names = ["f%d" % i for i in range(Xg.shape[1])]
    r = rf_pipeline.run_pipeline(Xg, y, names)
    assert len(set(r["indices"]) & set(GOLD[tag + "_indices"])) >= 18   # k-NN MI estimates: a tie may swap one or two
    both = np.intersect1d(r["indices"], GOLD[tag + "_indices"])
    gs = dict(zip(GOLD[tag + "_indices"], GOLD[tag + "_scores"])); rs = dict(zip(r["indices"], r["scores"]))
    assert max(abs(gs[i] - rs[i]) for i in both) <= 0.02                # mutual-information scores of the shared picks
    assert abs(r["cv_scores"].mean() - GOLD[tag + "_cv_scores"].mean()) <= 0.07
    assert abs(r["test_accuracy"] - float(GOLD[tag + "_test_accuracy"])) <= 0.17   # 6 test samples


@pytest.mark.parametrize("model,intensity", [("salt_and_pepper", 15), ("speckle", 35), ("poisson", 40), ("uniform", 25)])
def test_noise_sweep_through_rf_trainer(model, intensity):
    from oracle import extract_wst_features_training
    M, J, L = 64, 3, 8
    u8, y = rf_pipeline.make_patches(10, M, np.random.default_rng(7))
    np.random.seed(42)
    u8 = np.stack([noise.MODELS[model](im, intensity) for im in u8])
    Xg = gpu_features(u8, J, L)
    Xo = np.stack([extract_wst_features_training(im, J=J, L=L, cache_filters=True) for im in rf_pipeline.load_rgb(u8)])
    tau = 1e-3 * np.abs(Xo).max(axis=1, keepdims=True)
    assert float((np.abs(Xg - Xo) / np.maximum(np.abs(Xo), tau)).max()) <= 1e-4
    names = ["f%d" % i for i in range(Xg.shape[1])]
    a, b = rf_pipeline.run_pipeline(Xg, y, names), rf_pipeline.run_pipeline(Xo, y, names)
    assert len(set(a["indices"]) & set(b["indices"])) >= 18
    assert abs(a["cv_scores"].mean() - b["cv_scores"].mean()) <= 0.07
