"""CPU phase-emulation of the CUDA kernels — test infrastructure only (see wst_emu.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "wst-feature-extraction-for-remote-sensing-vegetation-classification-via-machine-learning_b200", "csrc")
LIB = os.path.join(HERE, "libwst_emu.so")
_lib = None


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, "wst_emu.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    global _lib
    if _lib is None:
        if _stale():
            subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", CSRC,
                            os.path.join(HERE, "wst_emu.cpp"), "-o", LIB], check=True)
        lib = ctypes.CDLL(LIB)
        lib.emu_last_error.restype = ctypes.c_char_p
        lib.emu_fft2.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.emu_forward.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.emu_filter_bank.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 2
        lib.emu_query.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3
        _lib = lib
    return _lib


def fft2(x, direction):
    """2-D FFT of a square complex64 array through the kernel's pass functions (unnormalised)."""
    a = np.ascontiguousarray(x, np.complex64).copy()
    rc = load().emu_fft2(a.shape[0], direction, a.ctypes.data)
    if rc != 0:
        raise RuntimeError(load().emu_last_error().decode())
    return a


def filter_bank(N, J, L):
    psi = np.empty((J * L, N, N), np.float32)
    phi = np.empty((N, N), np.float32)
    rc = load().emu_filter_bank(N, J, L, psi.ctypes.data, phi.ctypes.data)
    if rc != 0:
        raise RuntimeError(load().emu_last_error().decode())
    return psi, phi


def forward(x, J, L, max_order, psi_hat, phi_hat, with_features=False):
    """Replay the cascade kernel for signals x [nsig, H, W]; returns maps [nsig, K, h, h]
    (and the in-kernel pooled features [nsig, 2, K] when with_features)."""
    x = np.ascontiguousarray(x, np.float32)
    nsig, H, W = x.shape
    N = psi_hat.shape[-1]
    K = 1 + L * J + (L * L * J * (J - 1) // 2 if max_order >= 2 else 0)
    h = N // 2 ** J - 2
    out = np.full((nsig, K, h, h), np.nan, np.float32)
    psi_hat = np.ascontiguousarray(psi_hat, np.float32)
    phi_hat = np.ascontiguousarray(phi_hat, np.float32)
    feats = np.full((nsig, 2, K), np.nan, np.float32) if with_features else None
    rc = load().emu_forward(N, J, L, max_order, H, W, psi_hat.ctypes.data, phi_hat.ctypes.data,
                            x.ctypes.data, nsig, out.ctypes.data, feats.ctypes.data if with_features else None)
    if rc != 0:
        raise RuntimeError(load().emu_last_error().decode())
    return (out, feats) if with_features else out
