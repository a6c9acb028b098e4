"""Oracle for the 'advanced statistics' extractor.  TEST INFRASTRUCTURE ONLY.

Restates src/training/train_and_save_model.py:58-112 (duplicated at src/inference/inference.py:181-235):
18 statistics per channel, channels concatenated.  Unlike the WST path this one is PINNED: the reference
function itself runs in the build container (only numpy/scipy), tests/golden/make_golden_advstats.py imports
it to generate the committed fixtures, and tests/test_advstats.py holds this restatement to them bit for bit.
"""
import numpy as np
from scipy import stats
from scipy.ndimage import sobel, laplace

FEATURES = ['mean', 'std', 'var', 'min', 'max', 'range', 'skew', 'kurt', 'cv', 'p10', 'p25', 'p50', 'p75', 'p90',
            'iqr', 'mad', 'grad_mean', 'edge_density']        # train_and_save_model.py:402-405


def extract_advanced_features(rgb_image):
    """train_and_save_model.py:58-112, statement for statement (float32 in -> float64[C*18] out)."""
    features_per_channel = 18
    C = rgb_image.shape[0]
    features = np.zeros(C * features_per_channel)
    for i in range(C):
        channel = rgb_image[i]
        ch_flat = channel.ravel()
        ch_clean = ch_flat[np.isfinite(ch_flat)]
        if len(ch_clean) == 0:
            continue
        base = i * features_per_channel
        features[base + 0] = np.mean(ch_clean)
        features[base + 1] = np.std(ch_clean)
        features[base + 2] = np.var(ch_clean)
        features[base + 3] = np.min(ch_clean)
        features[base + 4] = np.max(ch_clean)
        features[base + 5] = np.ptp(ch_clean)
        features[base + 6] = stats.skew(ch_clean)
        features[base + 7] = stats.kurtosis(ch_clean)
        mean_val = features[base + 0]
        features[base + 8] = features[base + 1] / max(mean_val, 1e-8)
        features[base + 9] = np.percentile(ch_clean, 10)
        features[base + 10] = np.percentile(ch_clean, 25)
        features[base + 11] = np.percentile(ch_clean, 50)
        features[base + 12] = np.percentile(ch_clean, 75)
        features[base + 13] = np.percentile(ch_clean, 90)
        features[base + 14] = features[base + 12] - features[base + 10]
        features[base + 15] = np.mean(np.abs(ch_clean - mean_val))
        grad_x = sobel(channel, axis=0)
        grad_y = sobel(channel, axis=1)
        grad_mag = np.sqrt(grad_x ** 2 + grad_y ** 2)
        features[base + 16] = np.mean(grad_mag.ravel())
        edges = np.abs(laplace(channel))
        edge_thr = np.percentile(edges.ravel(), 90)
        features[base + 17] = np.mean(edges.ravel() > edge_thr)
    return features


def extract_hybrid_features(rgb_image, wst_features):
    """train_and_save_model.py:380-387: concat([advanced(54) float64, wst float32]) -> float64."""
    return np.concatenate([extract_advanced_features(rgb_image), wst_features])
