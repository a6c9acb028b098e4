"""CPU replay of the CUDA kernels' phases (tests/emu) against the oracle: this is how the index
arithmetic of csrc/wst_cascade.h and the filter-bank formula of csrc/wst_filters.h are checked in the
GPU-less container.  The GPU itself is checked by tests/test_gpu_parity.py."""
import numpy as np
import pytest

from tests import emu
from oracle import Scattering2D, filter_bank


def floored_rel(a, b):
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    tau = 1e-3 * np.abs(b).max(axis=1, keepdims=True)
    return float((np.abs(a - b) / np.maximum(np.abs(b), tau)).max())


@pytest.mark.parametrize("M", [6, 10, 12, 18, 20, 24, 34, 40, 48, 68, 80, 136, 160])
def test_fft_passes(M):
    rng = np.random.default_rng(M)
    x = (rng.standard_normal((M, M)) + 1j * rng.standard_normal((M, M))).astype(np.complex64)
    ref = np.fft.fft2(x.astype(np.complex128))
    assert np.abs(emu.fft2(x, -1) - ref).max() <= 1e-6 * np.abs(ref).max()
    ref = np.fft.ifft2(x.astype(np.complex128)) * M * M
    assert np.abs(emu.fft2(x, +1) - ref).max() <= 1e-6 * np.abs(ref).max()


@pytest.mark.parametrize("N,J,L", [(40, 2, 8), (48, 3, 6)])
def test_filter_bank_formula(N, J, L):
    psi, phi = emu.filter_bank(N, J, L)
    fb = filter_bank(N, N, J, L)
    opsi = np.stack([p["levels"][0] for p in fb["psi"]])
    assert np.abs(psi - opsi).max() < 1e-6
    assert np.abs(phi - fb["phi"]["levels"][0]).max() < 1e-6


@pytest.mark.parametrize("M,J,L,mo", [(32, 2, 8, 2), (32, 2, 8, 1), (32, 3, 6, 2), (64, 3, 8, 2), (128, 2, 8, 2)])
def test_cascade_vs_oracle(M, J, L, mo):
    rng = np.random.default_rng(7)
    x = (rng.integers(0, 256, (2, M, M)) / 255.0).astype(np.float32)
    S = Scattering2D(J=J, shape=(M, M), L=L, max_order=mo, precision="double", cache_filters=True)
    psi = np.stack([p["levels"][0] for p in S.psi]); phi = S.phi["levels"][0]
    got, feats = emu.forward(x, J, L, mo, psi, phi, with_features=True)
    assert not np.isnan(got).any() and not np.isnan(feats).any()
    ref = S(x)
    assert floored_rel(got, ref) <= 1e-4 / 4
    assert floored_rel(feats[:, 0], ref.mean(axis=(-2, -1))) <= 1e-4 / 4       # in-kernel pooling
    tau = 1e-3 * np.abs(ref.mean(axis=(-2, -1))).max(axis=1, keepdims=True)
    assert float((np.abs(feats[:, 1] - ref.std(axis=(-2, -1))) / np.maximum(ref.std(axis=(-2, -1)), tau)).max()) <= 1e-4 / 4
