"""BASELINE configs[3]: noise-robustness sweep at 64x64 J=3 feeding the reference's Random-Forest trainer.

Synthetic three-class texture patches are noised with the reference's models (tests/noise.py), features are
extracted on the GPU from the uint8 pixels (wst2d_forward_u8 = load_rgb_image + extract_wst_features) and on
the CPU by the oracle, and both feature matrices go through the reference's pipeline
(train_and_save_model.py:147-198: StandardScaler -> SelectKBest(mutual_info_classif, k) ->
RandomForest(max_features='sqrt', min_samples_split=5, min_samples_leaf=2), 5-fold stratified CV),
restated here with the same hyper-parameters.  The consumer must not be able to tell the two apart."""
import numpy as np
import pytest
import torch

from tests import noise

pytestmark = pytest.mark.gpu


def make_patches(n_per_class, M, rng):
    """Three vegetation-like classes: isotropic 1/f, oriented stripes + noise, blobby low-frequency."""
    f = np.fft.fftfreq(M)
    fr = np.maximum(np.hypot(*np.meshgrid(f, f, indexing="ij")), 1.0 / M)
    out, y = [], []
    for cls in range(3):
        for _ in range(n_per_class):
            ch = []
            for c in range(3):
                ph = np.exp(2j * np.pi * rng.random((M, M)))
                if cls == 0:
                    img = np.real(np.fft.ifft2(ph / fr))
                elif cls == 1:
                    xx = np.arange(M)[None, :] + 0.3 * np.arange(M)[:, None]
                    img = np.sin(2 * np.pi * xx / (6 + c)) + 0.5 * np.real(np.fft.ifft2(ph / fr)) / 0.1
                else:
                    img = np.real(np.fft.ifft2(ph / fr ** 2))
                img = (img - img.min()) / (img.max() - img.min())
                ch.append(img)
            out.append(np.stack(ch, -1))
            y.append(cls)
    return (np.stack(out) * 255).astype(np.uint8), np.array(y)        # [n, M, M, 3] uint8, HWC like PIL


def reference_pipeline(X, y, k=20):
    from sklearn.preprocessing import StandardScaler
    from sklearn.feature_selection import SelectKBest, mutual_info_classif
    from sklearn.ensemble import RandomForestClassifier
    from sklearn.model_selection import StratifiedKFold, cross_val_score
    np.random.seed(42)
    Xs = StandardScaler().fit_transform(X)
    sel = SelectKBest(lambda a, b: mutual_info_classif(a, b, random_state=42), k=k).fit(Xs, y)
    idx = sel.get_support(indices=True)
    rf = RandomForestClassifier(n_estimators=10, max_features="sqrt", min_samples_split=5, min_samples_leaf=2,
                                random_state=42)
    cv = StratifiedKFold(n_splits=5, shuffle=True, random_state=42)
    return idx, cross_val_score(rf, Xs[:, idx], y, cv=cv, scoring="accuracy")


@pytest.mark.parametrize("model,intensity", [("clean", 0), ("gaussian", 30), ("salt_and_pepper", 15),
                                             ("speckle", 35), ("poisson", 40), ("uniform", 25)])
def test_noise_sweep_through_rf_trainer(model, intensity):
    import wst_b200
    from oracle import extract_wst_features_training
    M, J, L = 64, 3, 8
    rng = np.random.default_rng(7)
    u8, y = make_patches(10, M, rng)
    if model != "clean":
        np.random.seed(42)                                              # add_noise.py:147-149
        u8 = np.stack([noise.MODELS[model](im, intensity) for im in u8])
    plan = wst_b200.get_plan(M, M, J, L)
    feats, _ = plan.forward(torch.from_numpy(u8).cuda())                 # uint8 HWC ingest
    Xg = wst_b200.to_block(feats).cpu().numpy()
    chw = np.ascontiguousarray(np.transpose(u8.astype(np.float32) / 255.0, (0, 3, 1, 2)))   # load_rgb_image
    Xo = np.stack([extract_wst_features_training(im, J=J, L=L, cache_filters=True) for im in chw])
    assert Xg.shape == Xo.shape == (30, 3 * 2 * 217)
    tau = 1e-3 * np.abs(Xo).max(axis=1, keepdims=True)
    assert float((np.abs(Xg - Xo) / np.maximum(np.abs(Xo), tau)).max()) <= 1e-4
    idx_g, cv_g = reference_pipeline(Xg, y)
    idx_o, cv_o = reference_pipeline(Xo, y)
    assert len(set(idx_g) & set(idx_o)) >= 18                           # same selected features (ties may swap one or two)
    assert abs(cv_g.mean() - cv_o.mean()) <= 0.07                       # same accuracy within one sample of 30
