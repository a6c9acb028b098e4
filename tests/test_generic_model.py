"""The algebra the shape-generic engine (csrc/wst_generic.cu) relies on, checked against the oracle on the CPU:
"filter multiply -> Fourier fold by 2^d -> inverse FFT" is a *partial* inverse DFT matrix applied per axis, and
"phi^ multiply -> fold -> inverse FFT -> unpad" is a pair of real separable operators built from phi^'s first
column and first row.  Rectangular and non-smooth padded sizes on purpose (the reference builds its transform
from the image's own shape, train_and_save_model.py:355-359).  The CUDA engine itself is tested on the GPU
(tests/test_gpu_generic.py); this file only pins the formulation."""
import numpy as np
import pytest

from oracle import Scattering2D, compute_padding


def dft(rows, cols, period, sign, scale):
    i, k = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    return scale * np.exp(sign * 2j * np.pi * ((i * k) % period) / period)


def crop(f, mh, mw):
    Hp, Wp = f.shape
    kk = np.where(np.arange(mh) < mh // 2, np.arange(mh), Hp - mh + np.arange(mh))
    ll = np.where(np.arange(mw) < mw // 2, np.arange(mw), Wp - mw + np.arange(mw))
    return f[np.ix_(kk, ll)]


def lowpass_ops(phi0, j, J, h, w):
    Hp, Wp = phi0.shape
    rs = 1.0 / np.sqrt(phi0[0, 0])
    out = []
    for n0, nout, a_full in ((Hp, h, phi0[:, 0]), (Wp, w, phi0[0, :])):
        m = n0 >> j
        kk = np.where(np.arange(m) < m // 2, np.arange(m), n0 - m + np.arange(m))
        a = a_full[kk] * rs
        x = np.arange(m)
        g = (a[None, :] * np.cos(2 * np.pi * ((x[:, None] * np.arange(m)[None, :]) % m) / m)).sum(1) / m
        s = 1 << (J - j)
        out.append(np.stack([g[((i + 1) * s - x) % m] for i in range(nout)]))
    return out


def model(x, J, L, S):
    """DFT-matrix formulation, float64."""
    H, W = x.shape
    Hp, Wp = compute_padding(H, W, J)
    top, left = (Hp - H) // 2, (Wp - W) // 2
    z = np.pad(x.astype(np.float64), ((top, Hp - H - top), (left, Wp - W - left)), mode="reflect")
    h, w = (Hp >> J) - 2, (Wp >> J) - 2
    phi0 = S.phi["levels"][0].astype(np.float64)
    psi = [p["levels"][0].astype(np.float64) for p in S.psi]
    F = lambda n: dft(n, n, n, -1, 1.0)
    A = lambda nc, npar: dft(nc, npar, nc, +1, 1.0 / npar)
    low = lambda u, j: (lambda G: G[0] @ u @ G[1].T)(lowpass_ops(phi0, j, J, h, w))
    U0 = F(Hp) @ z @ F(Wp).T
    out = [low(z, 0)]
    o2 = []
    for n1 in range(J * L):
        j1 = n1 // L
        H1, W1 = Hp >> j1, Wp >> j1
        U1 = np.abs(A(H1, Hp) @ (U0 * psi[n1]) @ A(W1, Wp).T)
        out.append(low(U1, j1))
        if j1 < J - 1:
            U1h = F(H1) @ U1 @ F(W1).T
            for n2 in range(J * L):
                j2 = n2 // L
                if j2 <= j1:
                    continue
                H2, W2 = Hp >> j2, Wp >> j2
                U2 = np.abs(A(H2, H1) @ (U1h * crop(psi[n2], H1, W1)) @ A(W2, W1).T)
                o2.append(low(U2, j2))
    return np.stack(out + o2)


@pytest.mark.parametrize("H,W,J,L", [(20, 28, 2, 4), (25, 22, 1, 8), (36, 44, 3, 3)])
def test_dft_matrix_formulation_matches_oracle(H, W, J, L):
    rng = np.random.default_rng(3)
    x = rng.random((H, W))
    S = Scattering2D(J=J, shape=(H, W), L=L, precision="double")
    ref = S(x)
    got = model(x, J, L, S)
    assert got.shape == ref.shape
    tau = 1e-3 * np.abs(ref).max()
    assert (np.abs(got - ref) / np.maximum(np.abs(ref), tau)).max() < 2e-6
