"""The consumer of the features: the reference's Random-Forest pipeline (train_and_save_model.py:147-198).

`reference_trainer()` returns the reference's OWN select_features_kbest / train_final_model, imported unchanged from
/root/reference when that tree is present (the build container).  /root/reference does not travel to the GPU box, so
a statement-for-statement restatement is kept next to it; tests/test_rf_pipeline.py holds the restatement to the
imported functions bit for bit (build container) and to tests/golden/rf_pipeline.npz, which was produced by the
imported functions (tests/golden/make_golden_rf.py).  Test infrastructure only."""
import importlib.util
import os

import numpy as np

REF_TRAIN = "/root/reference/src/training/train_and_save_model.py"


def make_patches(n_per_class, M, rng):
    """Three vegetation-like classes: isotropic 1/f, oriented stripes + noise, blobby low-frequency.
    Returns uint8 [n, M, M, 3] (HWC like PIL) and labels."""
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
    return (np.stack(out) * 255).astype(np.uint8), np.array(y)


def load_rgb(u8):
    """load_rgb_image (train_and_save_model.py:51-56) on in-memory pixels: uint8 HWC -> float32 / 255 CHW."""
    return np.ascontiguousarray(np.transpose(u8.astype(np.float32) / 255.0, (0, 3, 1, 2)))


# ---- restatement of train_and_save_model.py:147-198 (same calls, same hyper-parameters, same return values)
def select_features_kbest(X, y, feature_names, k=5):
    from sklearn.feature_selection import SelectKBest, mutual_info_classif
    from sklearn.preprocessing import StandardScaler
    scaler = StandardScaler()
    X_scaled = scaler.fit_transform(X)
    selector = SelectKBest(mutual_info_classif, k=k)
    X_selected = selector.fit_transform(X_scaled, y)
    selected_indices = selector.get_support(indices=True)
    if len(feature_names) < X.shape[1]:
        feature_names = feature_names + [f"feature_{i}" for i in range(len(feature_names), X.shape[1])]
    selected_features = [feature_names[i] for i in selected_indices]
    feature_scores = selector.scores_[selected_indices]
    return X_selected, selected_features, feature_scores, scaler, selector


def train_final_model(X, y, test_size=0.2, random_state=42, n_estimators=50, cv_folds=5):
    from sklearn.ensemble import RandomForestClassifier
    from sklearn.metrics import accuracy_score, classification_report, confusion_matrix
    from sklearn.model_selection import StratifiedKFold, cross_val_score, train_test_split
    X_train, X_test, y_train, y_test = train_test_split(X, y, test_size=test_size, random_state=random_state, stratify=y)
    rf = RandomForestClassifier(n_estimators=n_estimators, max_features='sqrt', min_samples_split=5,
                                min_samples_leaf=2, random_state=random_state)
    rf.fit(X_train, y_train)
    y_pred = rf.predict(X_test)
    test_accuracy = accuracy_score(y_test, y_pred)
    cv = StratifiedKFold(n_splits=cv_folds, shuffle=True, random_state=random_state)
    cv_scores = cross_val_score(rf, X, y, cv=cv, scoring='accuracy')
    return rf, {
        'test_accuracy': test_accuracy,
        'cv_mean_accuracy': float(np.mean(cv_scores)),
        'cv_std_accuracy': float(np.std(cv_scores)),
        'cv_scores': cv_scores.tolist(),
        'classification_report': classification_report(y_test, y_pred, output_dict=True),
        'confusion_matrix': confusion_matrix(y_test, y_pred).tolist()
    }


def reference_trainer():
    """(select_features_kbest, train_final_model, 'reference' | 'restatement')."""
    if os.path.exists(REF_TRAIN):
        spec = importlib.util.spec_from_file_location("ref_train_and_save_model", REF_TRAIN)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.select_features_kbest, mod.train_final_model, "reference"
    return select_features_kbest, train_final_model, "restatement"


def run_pipeline(X, y, names, k=20, n_estimators=10, fns=None):
    """The reference main()'s two calls (train_and_save_model.py:509-523) under np.random.seed(42) (mutual_info_classif
    draws its tie-breaking noise from numpy's global stream)."""
    sel, train = fns or reference_trainer()[:2]
    np.random.seed(42)
    X_selected, selected_features, feature_scores, scaler, selector = sel(X, y, names, k=k)
    model, perf = train(X_selected, y, n_estimators=n_estimators)
    return {"indices": selector.get_support(indices=True), "names": selected_features,
            "scores": np.asarray(feature_scores), "cv_scores": np.asarray(perf["cv_scores"]),
            "test_accuracy": float(perf["test_accuracy"]), "confusion_matrix": np.asarray(perf["confusion_matrix"])}
