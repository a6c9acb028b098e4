"""The C-ABI library builds for sm_100a, loads without a GPU and exports exactly what include/wst2d.h
declares; argument validation and the no-GPU failure mode are checked without any compute call."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "wst2d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wst2d_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    names = declared_functions()
    assert {"wst2d_plan_create", "wst2d_forward", "wst2d_forward_host", "wst2d_forward_u8", "wst2d_query",
            "wst2d_plan_destroy", "wst2d_last_error"} <= set(names)
    for n in names:
        assert hasattr(built_lib, n), n
    from wst_b200 import _lib
    assert sorted(_lib.SYMBOLS) == names


def test_sass_is_sm100a():
    from wst_b200._build import LIB_PATH
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_argument_validation(built_lib):
    h = ctypes.c_void_p()
    assert built_lib.wst2d_plan_create(None, 0, 32, 32, 2, 8, 2) == -1
    assert built_lib.wst2d_plan_create(ctypes.byref(h), 0, 32, 32, 6, 8, 2) == -1          # J out of range
    assert built_lib.wst2d_plan_create(ctypes.byref(h), 0, 32, 32, 2, 8, 3) == -1          # max_order
    assert built_lib.wst2d_plan_create(ctypes.byref(h), 0, 16, 16, 5, 8, 2) == -1          # 2^J > M
    assert b"2^J" in built_lib.wst2d_last_error()
    FFT = 1                                                                                # WST2D_ENGINE_FFT
    assert built_lib.wst2d_plan_create_ex(ctypes.byref(h), 0, 100, 100, 2, 8, 2, FFT) == -2   # no compiled cascade
    assert b"no compiled cascade" in built_lib.wst2d_last_error()
    assert built_lib.wst2d_plan_create_ex(ctypes.byref(h), 0, 32, 64, 2, 8, 2, FFT) == -2     # non-square
    assert built_lib.wst2d_plan_create_ex(ctypes.byref(h), 0, 32, 32, 2, 16, 2, FFT) == -2    # L > 8
    assert built_lib.wst2d_plan_create_ex(ctypes.byref(h), 0, 32, 32, 2, 8, 2, 7) == -1       # unknown engine
    assert built_lib.wst2d_plan_create(ctypes.byref(h), 0, 16, 40, 4, 8, 2) == -2             # pad as wide as the image
    assert built_lib.wst2d_debug_num_phase_tags() == 168
    assert built_lib.wst2d_query(None, None, None, None, None, None) == -1
    assert built_lib.wst2d_forward(None, None, 1, 1, None, None, None) == -1
    assert built_lib.wst2d_plan_destroy(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built_lib):
    h = ctypes.c_void_p()
    assert built_lib.wst2d_plan_create(ctypes.byref(h), 0, 32, 32, 2, 8, 2) == -3
    assert not h.value
    import numpy as np
    import wst_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wst_b200.extract_wst_features(np.zeros((3, 32, 32), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wst_b200.numpy.Scattering2D(J=2, shape=(32, 32))(np.zeros((32, 32), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "wst-feature-extraction-for-remote-sensing-vegetation-classification-via-machine-learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".inc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "tests.emu" not in txt, f
