"""The Random-Forest consumer (train_and_save_model.py:147-198) on CPU: the restatement kept for the GPU box equals
the reference's own functions (imported unchanged where /root/reference exists), and both reproduce the committed
golden record that the reference's functions produced from oracle features."""
import os

import numpy as np
import pytest

from tests import rf_pipeline

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "rf_pipeline.npz"))


@pytest.mark.parametrize("tag", ["clean", "gaussian30"])
def test_restated_pipeline_reproduces_reference_golden(tag):
    X, y = GOLD[tag + "_X"], GOLD["y"]
    names = ["f%d" % i for i in range(X.shape[1])]
    r = rf_pipeline.run_pipeline(X, y, names, fns=(rf_pipeline.select_features_kbest, rf_pipeline.train_final_model))
    assert np.array_equal(r["indices"], GOLD[tag + "_indices"])
    np.testing.assert_allclose(r["scores"], GOLD[tag + "_scores"], rtol=0, atol=1e-12)
    assert np.array_equal(r["cv_scores"], GOLD[tag + "_cv_scores"])
    assert r["test_accuracy"] == float(GOLD[tag + "_test_accuracy"])


@pytest.mark.skipif(not os.path.exists(rf_pipeline.REF_TRAIN), reason="/root/reference exists in the build container only")
def test_restatement_equals_reference_functions():
    sel, train, kind = rf_pipeline.reference_trainer()
    assert kind == "reference"
    X, y = GOLD["gaussian30_X"], GOLD["y"]
    names = ["f%d" % i for i in range(100)]                    # fewer names than columns: the padding branch (:158-160)
    a = rf_pipeline.run_pipeline(X, y, names, fns=(sel, train))
    b = rf_pipeline.run_pipeline(X, y, names, fns=(rf_pipeline.select_features_kbest, rf_pipeline.train_final_model))
    assert np.array_equal(a["indices"], b["indices"]) and a["names"] == b["names"]
    assert np.array_equal(a["scores"], b["scores"]) and np.array_equal(a["cv_scores"], b["cv_scores"])
    assert a["test_accuracy"] == b["test_accuracy"] and np.array_equal(a["confusion_matrix"], b["confusion_matrix"])
