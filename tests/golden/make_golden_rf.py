"""Golden record of the reference's Random-Forest pipeline on oracle features (BASELINE configs[3] shape).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_rf.py
Synthetic three-class 64x64 RGB patches (clean and gaussian sigma 30, add_noise.py:14-21 with seed 42) -> oracle WST
features (J=3, L=8, training layout) -> the reference's own select_features_kbest + train_final_model
(train_and_save_model.py:147-198, imported unchanged) -> rf_pipeline.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import extract_wst_features_training  # noqa: E402
from tests import noise, rf_pipeline  # noqa: E402


def main():
    sel, train, kind = rf_pipeline.reference_trainer()
    assert kind == "reference", "run where /root/reference exists"
    M, J, L = 64, 3, 8
    u8, y = rf_pipeline.make_patches(10, M, np.random.default_rng(7))
    out = {"y": y}
    for tag, model, intensity in (("clean", None, 0), ("gaussian30", "gaussian", 30)):
        px = u8
        if model:
            np.random.seed(42)
            px = np.stack([noise.MODELS[model](im, intensity) for im in u8])
        X = np.stack([extract_wst_features_training(im, J=J, L=L, cache_filters=True) for im in rf_pipeline.load_rgb(px)])
        names = ["f%d" % i for i in range(X.shape[1])]
        r = rf_pipeline.run_pipeline(X, y, names, fns=(sel, train))
        out[tag + "_u8_sum"] = np.array([int(px.astype(np.int64).sum()), int((px.astype(np.int64) * np.arange(px.size).reshape(px.shape) % 65521).sum())])
        out[tag + "_X"] = X.astype(np.float32)
        for k in ("indices", "scores", "cv_scores", "test_accuracy", "confusion_matrix"):
            out[tag + "_" + k] = np.asarray(r[k])
        print(tag, X.shape, "selected", r["indices"][:8], "cv", r["cv_scores"], "test", r["test_accuracy"])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rf_pipeline.npz"), **out)


if __name__ == "__main__":
    main()
