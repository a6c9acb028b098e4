"""Oracle: CPU restatement of kymatio==0.3.0 `Scattering2D` (NumPy frontend) and of the
reference's four call sites around it.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED.  kymatio 0.3.0 (reference `requirements.txt:18`) is a third-party,
un-vendored dependency; its source is not under /root/reference and cannot be fetched.
The functions below restate its published algorithm:

    kymatio/scattering2d/utils.py::compute_padding
    kymatio/scattering2d/filter_bank.py::{filter_bank, periodize_filter_fft, morlet_2d, gabor_2d}
    kymatio/scattering2d/backend/numpy_backend.py::{Pad, unpad, subsample_fourier, rfft, ifft, irfft, cdgmm, modulus}
    kymatio/scattering2d/core/scattering2d.py::scattering2d
    kymatio/scattering2d/frontend/{base_frontend,numpy_frontend}.py

and the reference-side wrappers restate, with file:line,

    src/training/train_and_save_model.py:346-378        extract_wst_features (canonical, block layout)
    src/inference/inference.py:237-270                  ModelInference.extract_wst_features (interleaved)
    src/visualization/visualize_features.py:194-222     extract_wst_features (grayscale, returns maps)
    src/visualization/compare_wst_coefficients.py:35-39 compute_scattering_coefficients (L=6, J=3, negated)

Quirks kept on purpose (they change the numbers): the literal 3.1415 in the Gabor
normalisation, the float32 rotation matrix, complex64 accumulation of the 5x5
periodisation, `real(fft2(.))` for the Fourier-domain filters, full complex FFT of real
signals, mean-fold for signals vs masked sum-fold for filters, `[1:-1, 1:-1]` unpad.

`precision='double'` runs the same dataflow in float64/complex128 on the same float32
filters (what kymatio does for float64 inputs, e.g. visualize_features.py:213).
"""
from __future__ import annotations

import numpy as np
import scipy.fft as _fft


# --------------------------------------------------------------------------- geometry
def compute_padding(M, N, J):
    """kymatio/scattering2d/utils.py::compute_padding."""
    M_padded = ((M + 2 ** J) // 2 ** J + 1) * 2 ** J
    N_padded = ((N + 2 ** J) // 2 ** J + 1) * 2 ** J
    return M_padded, N_padded


def num_coefficients(J, L=8, max_order=2):
    K = 1 + L * J
    if max_order >= 2:
        K += L * L * J * (J - 1) // 2
    return K


# --------------------------------------------------------------------------- filter bank
def gabor_2d(M, N, sigma, theta, xi, slant=1.0, offset=0):
    """kymatio filter_bank.gabor_2d: 5x5-periodised Gabor, complex64 accumulation."""
    gab = np.zeros((M, N), np.complex64)
    R = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]], np.float32)
    R_inv = np.array([[np.cos(theta), np.sin(theta)], [-np.sin(theta), np.cos(theta)]], np.float32)
    D = np.array([[1, 0], [0, slant * slant]])
    curv = np.dot(R, np.dot(D, R_inv)) / (2 * sigma * sigma)

    for ex in [-2, -1, 0, 1, 2]:
        for ey in [-2, -1, 0, 1, 2]:
            [xx, yy] = np.mgrid[offset + ex * M:offset + M + ex * M,
                                offset + ey * N:offset + N + ey * N]
            arg = -(curv[0, 0] * np.multiply(xx, xx) + (curv[0, 1] + curv[1, 0]) * np.multiply(xx, yy)
                    + curv[1, 1] * np.multiply(yy, yy)) \
                + 1.j * (xx * xi * np.cos(theta) + yy * xi * np.sin(theta))
            gab += np.exp(arg)

    norm_factor = (2 * 3.1415 * sigma * sigma / slant)
    gab /= norm_factor
    return gab


def morlet_2d(M, N, sigma, theta, xi, slant=0.5, offset=0):
    """kymatio filter_bank.morlet_2d: Gabor minus its DC-cancelling envelope."""
    wv = gabor_2d(M, N, sigma, theta, xi, slant, offset)
    wv_modulus = gabor_2d(M, N, sigma, theta, 0, slant, offset)
    K = np.sum(wv) / np.sum(wv_modulus)
    mor = wv - K * wv_modulus
    return mor


def periodize_filter_fft(x, res):
    """kymatio filter_bank.periodize_filter_fft.

    kymatio zeroes the rows/cols [M/2^(res+1), M/2^(res+1) + M(1-2^-res)) and then
    sum-folds 2^res x 2^res blocks with a 4-deep Python loop.  After the mask exactly one
    alias per output sample is non-zero, so the vectorised fold below gives bit-identical
    float32 values irrespective of summation order.
    """
    M, N = x.shape
    mask = np.ones(x.shape, np.float32)
    len_x = int(M * (1 - 2 ** (-res)))
    start_x = int(M * 2 ** (-res - 1))
    len_y = int(N * (1 - 2 ** (-res)))
    start_y = int(N * 2 ** (-res - 1))
    mask[start_x:start_x + len_x, :] = 0
    mask[:, start_y:start_y + len_y] = 0
    x = np.multiply(x, mask)
    k = 2 ** res
    crop = x.reshape(k, M // k, k, N // k).sum(axis=(0, 2), dtype=x.dtype)
    return crop


def filter_bank(M, N, J, L=8):
    """kymatio filter_bank.filter_bank: real float32 Fourier-domain Morlet bank + low-pass."""
    filters = {'psi': []}
    for j in range(J):
        for theta in range(L):
            psi = {'levels': [], 'j': j, 'theta': theta}
            psi_signal = morlet_2d(M, N, 0.8 * 2 ** j,
                                   (int(L - L / 2 - 1) - theta) * np.pi / L,
                                   3.0 / 4.0 * np.pi / 2 ** j, 4.0 / L)
            psi_signal_fourier = np.real(_fft.fft2(psi_signal))
            psi_levels = []
            for res in range(min(j + 1, max(J - 1, 1))):
                psi_levels.append(periodize_filter_fft(psi_signal_fourier, res))
            psi['levels'] = psi_levels
            filters['psi'].append(psi)

    phi_signal = gabor_2d(M, N, 0.8 * 2 ** (J - 1), 0, 0)
    phi_signal_fourier = np.real(_fft.fft2(phi_signal))
    filters['phi'] = {'levels': [], 'j': J}
    for res in range(J):
        filters['phi']['levels'].append(periodize_filter_fft(phi_signal_fourier, res))
    return filters


# --------------------------------------------------------------------------- backend ops
class Pad:
    """kymatio numpy_backend.Pad: reflect padding without repeating the edge sample."""

    def __init__(self, pad_size, input_size):
        self.pad_size = pad_size
        self.input_size = input_size
        pad_size_tmp = list(pad_size)
        # kymatio: "This handles the case where the padding is equal to the image size"
        if pad_size_tmp[0] == input_size[0]:
            pad_size_tmp[0] -= 1
            pad_size_tmp[1] -= 1
        if pad_size_tmp[2] == input_size[1]:
            pad_size_tmp[2] -= 1
            pad_size_tmp[3] -= 1
        self.padding_module = ((pad_size_tmp[0], pad_size_tmp[1]), (pad_size_tmp[2], pad_size_tmp[3]))

    def __call__(self, x):
        paddings = ((0, 0),) * (x.ndim - 2) + self.padding_module
        output = np.pad(x, paddings, mode='reflect')
        if self.pad_size[0] == self.input_size[0]:
            output = np.concatenate([output[..., 1:2, :], output, output[..., -2:-1, :]], axis=-2)
        if self.pad_size[2] == self.input_size[1]:
            output = np.concatenate([output[..., :, 1:2], output, output[..., :, -2:-1]], axis=-1)
        return output


def unpad(in_):
    return in_[..., 1:-1, 1:-1]


def subsample_fourier(x, k):
    y = x.reshape(x.shape[:-2] + (k, x.shape[-2] // k, k, x.shape[-1] // k))
    return y.mean(axis=(-4, -2))


def cdgmm(A, B):
    return A * B


def modulus(x):
    return np.abs(x)


def rfft(x):
    return _fft.fft2(x)


def ifft(x):
    return _fft.ifft2(x)


def irfft(x):
    return _fft.ifft2(x).real


# --------------------------------------------------------------------------- core
def scattering2d(x, pad, unpad_fn, J, L, phi, psi, max_order):
    """kymatio core/scattering2d.py::scattering2d with out_type='array'."""
    out_S_0, out_S_1, out_S_2 = [], [], []

    U_r = pad(x)
    U_0_c = rfft(U_r)

    U_1_c = cdgmm(U_0_c, phi['levels'][0])
    U_1_c = subsample_fourier(U_1_c, k=2 ** J)
    S_0 = unpad_fn(irfft(U_1_c))
    out_S_0.append(S_0)

    for n1 in range(len(psi)):
        j1 = psi[n1]['j']
        U_1_c = cdgmm(U_0_c, psi[n1]['levels'][0])
        if j1 > 0:
            U_1_c = subsample_fourier(U_1_c, k=2 ** j1)
        U_1_c = ifft(U_1_c)
        U_1_c = modulus(U_1_c)
        U_1_c = rfft(U_1_c)

        S_1_c = cdgmm(U_1_c, phi['levels'][j1])
        S_1_c = subsample_fourier(S_1_c, k=2 ** (J - j1))
        S_1_r = unpad_fn(irfft(S_1_c))
        out_S_1.append(S_1_r)

        if max_order < 2:
            continue
        for n2 in range(len(psi)):
            j2 = psi[n2]['j']
            if j2 <= j1:
                continue
            U_2_c = cdgmm(U_1_c, psi[n2]['levels'][j1])
            U_2_c = subsample_fourier(U_2_c, k=2 ** (j2 - j1))
            U_2_c = ifft(U_2_c)
            U_2_c = modulus(U_2_c)
            U_2_c = rfft(U_2_c)

            S_2_c = cdgmm(U_2_c, phi['levels'][j2])
            S_2_c = subsample_fourier(S_2_c, k=2 ** (J - j2))
            S_2_r = unpad_fn(irfft(S_2_c))
            out_S_2.append(S_2_r)

    out_S = out_S_0 + out_S_1 + out_S_2
    return np.stack(out_S, axis=-3)


_FILTER_CACHE: dict = {}


class Scattering2D:
    """kymatio.numpy.Scattering2D (frontend/numpy_frontend.py + base_frontend.py), out_type='array'.

    `cache_filters=True` reuses a filter bank across instances with the same
    (M_padded, N_padded, J, L); the values are identical, only construction time changes
    (the reference rebuilds the bank per image: train_and_save_model.py:359).
    """

    def __init__(self, J, shape, L=8, max_order=2, pre_pad=False, backend='numpy',
                 out_type='array', frontend='numpy', precision='single', cache_filters=False):
        self.J, self.shape, self.L, self.max_order = J, tuple(shape), L, max_order
        self.pre_pad, self.out_type, self.precision = pre_pad, out_type, precision
        M, N = self.shape
        if 2 ** J > M or 2 ** J > N:
            raise RuntimeError('The smallest dimension should be larger than 2^J.')
        self._M_padded, self._N_padded = compute_padding(M, N, J)
        if not pre_pad:
            self.pad = Pad([(self._M_padded - M) // 2, (self._M_padded - M + 1) // 2,
                            (self._N_padded - N) // 2, (self._N_padded - N + 1) // 2], [M, N])
        else:
            self.pad = lambda x: x
        self.unpad = unpad
        key = (self._M_padded, self._N_padded, J, L)
        if cache_filters and key in _FILTER_CACHE:
            filters = _FILTER_CACHE[key]
        else:
            filters = filter_bank(self._M_padded, self._N_padded, J, L)
            if cache_filters:
                _FILTER_CACHE[key] = filters
        self.phi, self.psi = filters['phi'], filters['psi']

    def scattering(self, input):
        if not type(input) is np.ndarray:
            raise TypeError('The input should be a NumPy array.')
        if len(input.shape) < 2:
            raise RuntimeError('Input array must have at least two dimensions.')
        if (input.shape[-1] != self.shape[-1] or input.shape[-2] != self.shape[-2]) and not self.pre_pad:
            raise RuntimeError('NumPy array must be of spatial size (%i,%i).' % (self.shape[0], self.shape[1]))
        if (input.shape[-1] != self._N_padded or input.shape[-2] != self._M_padded) and self.pre_pad:
            raise RuntimeError('Padded array must be of spatial size (%i,%i).' % (self._M_padded, self._N_padded))
        batch_shape = input.shape[:-2]
        signal_shape = input.shape[-2:]
        input = input.reshape((-1,) + signal_shape)
        if self.precision == 'double':
            input = input.astype(np.float64)
        S = scattering2d(input, self.pad, self.unpad, self.J, self.L, self.phi, self.psi, self.max_order)
        scattering_shape = S.shape[-3:]
        return S.reshape(batch_shape + scattering_shape)

    __call__ = scattering


# --------------------------------------------------------------------------- reference call sites
def pooled_features(coeffs):
    """mean and population std over the two spatial axes (train_and_save_model.py:371-372)."""
    return np.mean(coeffs, axis=(-2, -1)), np.std(coeffs, axis=(-2, -1))


def extract_wst_features_training(rgb_image, J=2, L=8, max_order=2, precision='single', cache_filters=False):
    """src/training/train_and_save_model.py:346-378 — per channel [mean(K) || std(K)], channels concatenated.

    The reference hard-codes J=2, L=8 (:352-353); they are parameters here so the other
    BASELINE configs can be checked with the same layout.
    """
    C, H, W = rgb_image.shape
    scattering = Scattering2D(J=J, L=L, shape=(H, W), max_order=max_order, precision=precision,
                              cache_filters=cache_filters)
    all_features = []
    for c in range(C):
        channel = rgb_image[c]
        scattering_coeffs = scattering(channel)
        coeffs_mean = np.mean(scattering_coeffs, axis=(-2, -1))
        coeffs_std = np.std(scattering_coeffs, axis=(-2, -1))
        channel_features = np.concatenate([coeffs_mean, coeffs_std])
        all_features.extend(channel_features)
    return np.array(all_features)


def extract_wst_features_inference(rgb_image, J=2, L=8, cache_filters=False):
    """src/inference/inference.py:237-270 — per channel interleaved [mean0,std0,mean1,std1,...], float64 container.

    The reference runs kymatio.torch on a (1,1,H,W) float32 tensor; the torch frontend
    performs the same float32 dataflow, restated here on the NumPy path.
    """
    num_channels, height, width = rgb_image.shape
    S = Scattering2D(J=J, shape=(height, width), L=L, cache_filters=cache_filters)
    all_features = []
    for channel_idx in range(num_channels):
        channel = np.ascontiguousarray(rgb_image[channel_idx])[None, None]
        coeffs = S(channel)[0, 0]
        num_coeffs = coeffs.shape[0]
        channel_features = np.zeros(2 * num_coeffs)
        for i in range(num_coeffs):
            coeff = coeffs[i].ravel()
            channel_features[2 * i] = np.mean(coeff)
            channel_features[2 * i + 1] = np.std(coeff)
        all_features.append(channel_features)
    return np.concatenate(all_features)


def extract_wst_features_visualization(grayscale_image, J=2, L=8, cache_filters=False):
    """src/visualization/visualize_features.py:194-222 — one channel, returns (features, coefficient maps).

    Inputs there are float64 (generators :50-120), so kymatio runs in complex128 on its float32 filters.
    """
    H, W = grayscale_image.shape
    precision = 'double' if grayscale_image.dtype == np.float64 else 'single'
    scattering = Scattering2D(J=J, L=L, shape=(H, W), precision=precision, cache_filters=cache_filters)
    scattering_coeffs = scattering(grayscale_image)
    coeffs_mean = np.mean(scattering_coeffs, axis=(-2, -1))
    coeffs_std = np.std(scattering_coeffs, axis=(-2, -1))
    return np.concatenate([coeffs_mean, coeffs_std]), scattering_coeffs


def compute_scattering_coefficients(img_tensor, L=6, J=3, cache_filters=False):
    """src/visualization/compare_wst_coefficients.py:35-39 — Scattering2D(J=3, L=6, max_order=2), negated."""
    scattering = Scattering2D(J=J, shape=img_tensor.shape, L=L, max_order=2, frontend='numpy',
                              cache_filters=cache_filters)
    scat_coeffs = scattering(img_tensor)
    return -scat_coeffs
