"""ctypes binding of libwst_b200.so (include/wst2d.h).  Fails loudly if the library is missing."""
import ctypes
import os

from ._build import LIB_PATH

WST2D_OK, WST2D_ERR_ARG, WST2D_ERR_UNSUPPORTED, WST2D_ERR_CUDA = 0, -1, -2, -3

# every symbol include/wst2d.h declares
SYMBOLS = ["wst2d_plan_create", "wst2d_plan_destroy", "wst2d_query", "wst2d_forward", "wst2d_forward_u8",
           "wst2d_forward_host", "wst2d_plan_filters", "wst2d_launch_count", "wst2d_last_error",
           "wst2d_version", "wst2d_profile", "wst2d_profile_read", "wst2d_fma_peak",
           "wst2d_debug_phase_cycles", "wst2d_forward_scene", "wst2d_advanced_stats",
           "wst2d_advanced_stats_last_error", "wst2d_add_noise", "wst2d_add_noise_draws", "wst2d_noise_last_error",
           "wst2d_plan_create_ex", "wst2d_plan_engine", "wst2d_debug_num_phase_tags", "wst2d_forward_host_u8", "wst2d_plan_grid"]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "wst_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.wst2d_plan_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32, i32, i32]
    lib.wst2d_plan_create_ex.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32, i32, i32, i32]
    lib.wst2d_plan_engine.argtypes = [vp]
    lib.wst2d_plan_grid.argtypes = [vp]
    lib.wst2d_debug_num_phase_tags.argtypes = []
    lib.wst2d_plan_destroy.argtypes = [vp]
    lib.wst2d_query.argtypes = [vp] + [ctypes.POINTER(i32)] * 5
    lib.wst2d_forward.argtypes = [vp, vp, i64, i32, vp, vp, vp]
    lib.wst2d_forward_u8.argtypes = [vp, vp, i64, i32, vp, vp, vp]
    lib.wst2d_forward_host.argtypes = [vp, vp, i64, i32, vp]
    lib.wst2d_forward_host_u8.argtypes = [vp, vp, i64, i32, vp]
    lib.wst2d_forward_scene.argtypes = [vp, vp, i32, i32, i32, i32, i32, i64, i64, vp, vp, vp]
    lib.wst2d_plan_filters.argtypes = [vp, vp, vp]
    lib.wst2d_launch_count.argtypes = [vp, i64, i32]
    lib.wst2d_profile.argtypes = [vp, i32]
    lib.wst2d_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i32)]
    lib.wst2d_debug_phase_cycles.argtypes = [vp, vp, i64, vp, i32]
    lib.wst2d_fma_peak.argtypes = [i32, ctypes.POINTER(ctypes.c_double)]
    lib.wst2d_advanced_stats.argtypes = [i32, vp, i32, i64, i32, i32, i32, vp, vp]
    lib.wst2d_advanced_stats_last_error.restype = ctypes.c_char_p
    f64, u64 = ctypes.c_double, ctypes.c_uint64
    lib.wst2d_add_noise.argtypes = [i32, i32, f64, vp, i64, i32, i32, i32, u64, vp, vp]
    lib.wst2d_add_noise_draws.argtypes = [i32, i32, f64, vp, i64, i32, i32, i32, vp, i64, vp, vp]
    lib.wst2d_noise_last_error.restype = ctypes.c_char_p
    lib.wst2d_last_error.restype = ctypes.c_char_p
    lib.wst2d_version.restype = ctypes.c_char_p
    for name in SYMBOLS:
        if name not in ("wst2d_last_error", "wst2d_version", "wst2d_advanced_stats_last_error", "wst2d_noise_last_error"):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def last_error():
    return load().wst2d_last_error().decode("utf-8", "replace")


def check(rc):
    """Map C-ABI return codes to the exceptions kymatio's frontends raise."""
    if rc == WST2D_OK:
        return
    msg = last_error()
    if rc == WST2D_ERR_UNSUPPORTED:
        raise NotImplementedError("wst_b200: " + msg)
    raise RuntimeError(msg if rc == WST2D_ERR_ARG else "wst_b200: " + msg)
