"""Host-side mirror of the reference's WST interface, backed by libwst_b200.so (sm_100a CUDA).

Reference surfaces mirrored:
  B3  kymatio `Scattering2D(J, shape, L=8, max_order=2, ...)`  -> Scattering2D (numpy / torch flavours)
  B1  extract_wst_features(rgb_image)            train_and_save_model.py:346-378  (block layout)
  B2  ModelInference.extract_wst_features(...)   inference.py:237-270             (interleaved layout)
      extract_wst_features(grayscale_image)      visualize_features.py:194-222    (features, maps)
      compute_scattering_coefficients(img, L, J) compare_wst_coefficients.py:35-39
  B4  batched op scattering_features / scattering_maps on CUDA tensors (what the kernels sit behind)

torch is used for device memory and streams only.  There is no CPU fallback: without the CUDA
library or a GPU every entry point raises.
"""
import ctypes
import threading

import numpy as np
import torch

from . import _lib
from ._parallel import shard_range, shard_sizes, gather_features

__all__ = [
    "Plan", "get_plan", "scattering_features", "scattering_maps", "features_from_host", "scene_features",
    "Scattering2D", "ScatteringNumPy2D", "ScatteringTorch2D",
    "extract_wst_features", "extract_wst_features_interleaved", "extract_wst_features_gray",
    "compute_scattering_coefficients", "extract_wst_features_batch", "num_coefficients",
    "ENGINES", "compute_padding", "to_interleaved", "to_block", "advanced_stats", "extract_advanced_features",
    "extract_hybrid_features", "hybrid_features", "ADVANCED_STAT_NAMES",
    "extract_features", "get_feature_names", "extract_basic_features", "basic_stats",
    "extract_features_inference", "ModelInferenceFeatures",
    "NOISE_TYPES", "add_noise", "add_gaussian_noise", "add_salt_and_pepper_noise", "add_speckle_noise",
    "add_poisson_noise", "add_uniform_noise", "shard_range", "shard_sizes", "gather_features", "fma_peak_tflops",
]


# ----------------------------------------------------------------------------- geometry
def compute_padding(M, N, J):
    """kymatio scattering2d/utils.py::compute_padding."""
    return ((M + 2 ** J) // 2 ** J + 1) * 2 ** J, ((N + 2 ** J) // 2 ** J + 1) * 2 ** J


def num_coefficients(J, L=8, max_order=2):
    K = 1 + L * J
    if max_order >= 2:
        K += L * L * J * (J - 1) // 2
    return K


# ----------------------------------------------------------------------------- plan
def _device_index(device):
    """CUDA device ordinal of `device` (None or an index-less 'cuda' = the current device)."""
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError("wst_b200 runs on CUDA devices only (got %s); there is no CPU fallback." % d)
    return torch.cuda.current_device() if d.index is None else d.index


ENGINES = {"auto": 0, "fft": 1, "gemm": 2, "gemm_tf32x3": 3}      # include/wst2d.h WST2D_ENGINE_*


class Plan:
    """Owns a wst2d_plan (filter bank + tables on one device) for Scattering2D(J, shape, L, max_order).

    engine: "auto" (the fused FFT cascade when one is compiled for the padded size, else the DFT-matrix engine, which
    takes any H x W and any L), "fft", "gemm" (fp32 SIMT) or "gemm_tf32x3" (tensor cores)."""

    def __init__(self, H, W, J, L=8, max_order=2, device=None, engine="auto"):
        if not torch.cuda.is_available():
            raise RuntimeError("wst_b200: no CUDA device available (this package has no CPU fallback).")
        lib = _lib.load()
        self.device = _device_index(device)
        self.H, self.W, self.J, self.L, self.max_order = int(H), int(W), int(J), int(L), int(max_order)
        h = ctypes.c_void_p()
        if engine not in ENGINES:
            raise ValueError("engine must be one of %s" % sorted(ENGINES))
        _lib.check(lib.wst2d_plan_create_ex(ctypes.byref(h), self.device, self.H, self.W, self.J, self.L,
                                            self.max_order, ENGINES[engine]))
        self._h = h
        self.engine = {v: k for k, v in ENGINES.items()}[lib.wst2d_plan_engine(h)]
        self.grid = lib.wst2d_plan_grid(h)          # signals in flight of the fused cascade (persistent grid)
        q = [ctypes.c_int() for _ in range(5)]
        _lib.check(lib.wst2d_query(h, *[ctypes.byref(v) for v in q]))
        self.K, self.h, self.w, self.Hp, self.Wp = [v.value for v in q]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.load().wst2d_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- device path -------------------------------------------------------------------------
    def _check_x(self, x):
        if not isinstance(x, torch.Tensor):
            raise TypeError("The input should be a PyTorch Tensor.")
        if x.dim() != 4:
            raise RuntimeError("Input tensor must be [B, C, H, W].")
        if x.shape[-2] != self.H or x.shape[-1] != self.W:
            raise RuntimeError("Tensor must be of spatial size (%i,%i)." % (self.H, self.W))
        if not x.is_cuda or x.device.index != self.device:
            raise RuntimeError("Input tensor must live on cuda:%d." % self.device)
        if not x.is_contiguous():
            raise RuntimeError("Tensor must be contiguous.")

    def forward(self, x, want_features=True, want_maps=False):
        """x: [B, C, H, W] float32 (or uint8 [B, H, W, C]) CUDA tensor -> (feats [B, C, 2, K] | None, maps | None)."""
        lib = _lib.load()
        u8 = x.dtype == torch.uint8
        if u8:
            if x.dim() != 4 or x.shape[1] != self.H or x.shape[2] != self.W:
                raise RuntimeError("uint8 input must be [B, H, W, C] of spatial size (%i,%i)." % (self.H, self.W))
            if not x.is_cuda or not x.is_contiguous():
                raise RuntimeError("uint8 input must be a contiguous CUDA tensor.")
            if x.device.index != self.device:
                raise RuntimeError("Input tensor must live on cuda:%d." % self.device)
            B, C = x.shape[0], x.shape[3]
        else:
            self._check_x(x)
            if x.dtype != torch.float32:
                raise TypeError("Input tensor must be float32 (or uint8 HWC).")
            B, C = x.shape[0], x.shape[1]
        if B == 0:
            return (torch.empty((0, C, 2, self.K), dtype=torch.float32, device=x.device) if want_features else None,
                    torch.empty((0, C, self.K, self.h, self.w), dtype=torch.float32, device=x.device) if want_maps else None)
        feats = torch.empty((B, C, 2, self.K), dtype=torch.float32, device=x.device) if want_features else None
        maps = torch.empty((B, C, self.K, self.h, self.w), dtype=torch.float32, device=x.device) if want_maps else None
        stream = torch.cuda.current_stream(x.device).cuda_stream
        fn = lib.wst2d_forward_u8 if u8 else lib.wst2d_forward
        _lib.check(fn(self._h, x.data_ptr(), B, C,
                      feats.data_ptr() if feats is not None else None,
                      maps.data_ptr() if maps is not None else None,
                      ctypes.c_void_p(stream)))
        return feats, maps

    # -- whole-scene tiling -------------------------------------------------------------------
    def tile_grid(self, Himg, Wimg, stride=None):
        """(ny, nx) tiles of the plan's H x W window over a Himg x Wimg raster with the given (sy, sx) step."""
        sy, sx = (self.H, self.W) if stride is None else ((stride, stride) if isinstance(stride, int) else stride)
        return (Himg - self.H) // sy + 1, (Wimg - self.W) // sx + 1

    def forward_scene(self, raster, stride=None, tile_range=None, want_maps=False):
        """raster: [C, Himg, Wimg] float32 CUDA tensor.  Tiles (row-major over tile_grid) tile_range=(begin, end)
        — default all — are processed straight from the raster: feats [ntiles, C, 2, K] (and maps)."""
        if not (isinstance(raster, torch.Tensor) and raster.is_cuda and raster.dim() == 3
                and raster.dtype == torch.float32 and raster.is_contiguous()):
            raise RuntimeError("raster must be a contiguous float32 CUDA tensor [C, Himg, Wimg].")
        C, Himg, Wimg = raster.shape
        sy, sx = (self.H, self.W) if stride is None else ((stride, stride) if isinstance(stride, int) else stride)
        ny, nx = self.tile_grid(Himg, Wimg, (sy, sx))
        lo, hi = (0, ny * nx) if tile_range is None else tile_range
        n = hi - lo
        feats = torch.empty((n, C, 2, self.K), dtype=torch.float32, device=raster.device)
        maps = torch.empty((n, C, self.K, self.h, self.w), dtype=torch.float32, device=raster.device) if want_maps else None
        if n > 0:
            stream = torch.cuda.current_stream(raster.device).cuda_stream
            _lib.check(_lib.load().wst2d_forward_scene(self._h, raster.data_ptr(), C, Himg, Wimg, sy, sx, lo, n,
                                                       feats.data_ptr(), maps.data_ptr() if want_maps else None,
                                                       ctypes.c_void_p(stream)))
        return feats, maps

    # -- host path ---------------------------------------------------------------------------
    def forward_host(self, x, out=None):
        """x: [B, C, H, W] float32 (or [B, H, W, C] uint8, PIL / load_rgb_image order) C-contiguous numpy array or CPU
        tensor (pinned for full overlap) -> feats [B, C, 2, K] numpy float32.  Copies are chunked and overlapped inside
        the library."""
        lib = _lib.load()
        if isinstance(x, torch.Tensor):
            if x.is_cuda:
                raise RuntimeError("forward_host expects host memory; use Plan.forward for CUDA tensors.")
            xa = x.numpy()
        else:
            xa = x
        if not isinstance(xa, np.ndarray) or xa.dtype not in (np.float32, np.uint8) or not xa.flags["C_CONTIGUOUS"] \
                or xa.ndim != 4:
            raise RuntimeError("forward_host expects a C-contiguous float32 [B, C, H, W] (or uint8 [B, H, W, C]) array.")
        u8 = xa.dtype == np.uint8
        hh, ww = (xa.shape[1], xa.shape[2]) if u8 else (xa.shape[-2], xa.shape[-1])
        if hh != self.H or ww != self.W:
            raise RuntimeError("NumPy array must be of spatial size (%i,%i)." % (self.H, self.W))
        B, C = (xa.shape[0], xa.shape[3]) if u8 else (xa.shape[0], xa.shape[1])
        if out is None:
            out = np.empty((B, C, 2, self.K), np.float32)
        if isinstance(out, torch.Tensor):
            if out.is_cuda:
                raise RuntimeError("forward_host: `out` must be a host (CPU) tensor, not a CUDA tensor.")
            oa = out.numpy()
        else:
            oa = out
        if not (isinstance(oa, np.ndarray) and oa.dtype == np.float32 and oa.flags["C_CONTIGUOUS"]
                and oa.flags["WRITEABLE"] and tuple(oa.shape) == (B, C, 2, self.K)):
            raise RuntimeError("forward_host: `out` must be a writeable C-contiguous float32 array of shape "
                               "(%d, %d, 2, %d)." % (B, C, self.K))
        fn = lib.wst2d_forward_host_u8 if u8 else lib.wst2d_forward_host
        _lib.check(fn(self._h, xa.ctypes.data, B, C, oa.ctypes.data))
        return out

    def filters(self):
        """(psi_hat [J*L, Hp, Wp], phi_hat [Hp, Wp]) float32 — the plan's full-resolution filter bank."""
        psi = np.empty((self.J * self.L, self.Hp, self.Wp), np.float32)
        phi = np.empty((self.Hp, self.Wp), np.float32)
        _lib.check(_lib.load().wst2d_plan_filters(self._h, psi.ctypes.data, phi.ctypes.data))
        return psi, phi

    def launch_count(self, B, C):
        return int(_lib.load().wst2d_launch_count(self._h, B, C))

    def phase_cycles(self, x):
        """Debug: {(phase kind, level): SM cycles of CTA 0} for one cascade launch over x [B, C, H, W]."""
        self._check_x(x)
        kinds = ["twiddle", "input", "lp1", "lp2", "rfft_row_s", "rfft_row_c", "rfft_split", "rfft_col_s",
                 "rfft_col_c", "u0_store", "prod1", "prod2", "ifft_col_c", "ifft_col_s", "ifft_row_c", "ifft_final",
                 "lp_reduce", "lp_store", "pool", "stage_load", "stage_store"]
        n = _lib.load().wst2d_debug_num_phase_tags()
        assert n == 8 * len(kinds), "phase tag table out of date"
        arr = (ctypes.c_int64 * n)()
        _lib.check(_lib.load().wst2d_debug_phase_cycles(self._h, x.data_ptr(), x.shape[0] * x.shape[1], arr, n))
        return {(kinds[i // 8], i % 8): int(arr[i]) for i in range(n) if arr[i]}

    def profile(self, enable=True):
        _lib.check(_lib.load().wst2d_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """(cascade_ms, pool_ms, cascade_launches) summed since the last read; synchronises the device."""
        c, p, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        _lib.check(_lib.load().wst2d_profile_read(self._h, ctypes.byref(c), ctypes.byref(p), ctypes.byref(n)))
        return c.value, p.value, n.value


_PLAN_CACHE = {}
_PLAN_LOCK = threading.Lock()


def get_plan(H, W, J, L=8, max_order=2, device=None, engine="auto"):
    """Plan cache keyed (device, H, W, J, L, max_order, engine): the filter bank is built once, not per image
    (the reference rebuilds it per image, train_and_save_model.py:359 — SURVEY.md F5)."""
    if not torch.cuda.is_available():
        raise RuntimeError("wst_b200: no CUDA device available (this package has no CPU fallback).")
    dev = _device_index(device)
    key = (dev, int(H), int(W), int(J), int(L), int(max_order), engine)
    with _PLAN_LOCK:
        p = _PLAN_CACHE.get(key)
        if p is None:
            p = Plan(H, W, J, L, max_order, dev, engine)
            _PLAN_CACHE[key] = p
        return p


def fma_peak_tflops(device=0):
    """Measured fp32 FMA peak of the device (TFLOP/s)."""
    t = ctypes.c_double()
    _lib.check(_lib.load().wst2d_fma_peak(int(device), ctypes.byref(t)))
    return t.value


# ----------------------------------------------------------------------------- layouts
def to_interleaved(feats_block):
    """[..., C, 2, K] (mean block, std block) -> [..., C*2*K] interleaved [mean0, std0, mean1, ...]
    (inference.py:263-266)."""
    f = feats_block
    f = f.swapaxes(-1, -2) if isinstance(f, np.ndarray) else f.transpose(-1, -2)
    return f.reshape(f.shape[:-3] + (-1,))


def to_block(feats_block):
    """[..., C, 2, K] -> [..., C*2*K] per channel [mean(K) || std(K)] (train_and_save_model.py:375)."""
    return feats_block.reshape(feats_block.shape[:-3] + (-1,))


# ----------------------------------------------------------------------------- batched device op (B4)
def scattering_features(x, J, L=8, max_order=2, layout="block"):
    """x: [B, C, H, W] float32 (or [B, H, W, C] uint8, load_rgb_image's input) CUDA tensor -> [B, C*2*K] pooled
    features on the same device.  Dispatches through the registered op torch.ops.wst.scattering2d_features."""
    if layout not in ("block", "interleaved"):
        raise ValueError("layout must be 'block' or 'interleaved'")
    _check_device_op_input(x)
    from . import _ops  # noqa: F401  (registers the wst:: ops)
    return torch.ops.wst.scattering2d_features(x, int(J), int(L), int(max_order), 0 if layout == "block" else 1, False)


def scattering_maps(x, J, L=8, max_order=2):
    """x: [B, C, H, W] float32 CUDA tensor -> coefficient maps [B, C, K, h, w] (torch.ops.wst.scattering2d_maps)."""
    _check_device_op_input(x)
    from . import _ops  # noqa: F401
    return torch.ops.wst.scattering2d_maps(x, int(J), int(L), int(max_order))


def _check_device_op_input(x):
    if not isinstance(x, torch.Tensor):
        raise TypeError("The input should be a PyTorch Tensor.")
    if not x.is_cuda:
        raise RuntimeError("wst_b200: the batched ops take CUDA tensors (this package has no CPU fallback).")


def scene_features(raster, tile, J, L=8, max_order=2, stride=None, rank=0, world_size=1, gather=False):
    """Whole-scene tiling (BASELINE configs[4]): slide a tile x tile window over raster [C, Himg, Wimg] (CUDA,
    float32), extract per-tile features, sharding the tiles contiguously over world_size ranks.
    Returns (feats [ntiles_local or ntiles, C*2*K], (ny, nx)); with gather=True every rank gets all tiles in
    row-major tile order through gather_features (NCCL over NVLink)."""
    plan = get_plan(tile, tile, J, L, max_order, raster.device)
    ny, nx = plan.tile_grid(raster.shape[1], raster.shape[2], stride)
    lo, hi = shard_range(ny * nx, rank, world_size)
    feats, _ = plan.forward_scene(raster, stride, (lo, hi))
    flat = to_block(feats)
    if gather and world_size > 1:
        flat = gather_features(flat, ny * nx)
    return flat, (ny, nx)


def features_from_host(x, J, L=8, max_order=2, layout="block", device=None):
    """x: [B, C, H, W] float32 host array -> [B, C*2*K] numpy features (host in, host out)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    plan = get_plan(x.shape[-2], x.shape[-1], J, L, max_order, device)
    feats = plan.forward_host(x)
    return to_block(feats) if layout == "block" else np.ascontiguousarray(to_interleaved(feats))


# ----------------------------------------------------------------------------- Scattering2D frontends (B3)
class _ScatteringBase2D:
    def __init__(self, J, shape, L=8, max_order=2, pre_pad=False, backend=None, out_type="array"):
        self.J, self.shape, self.L, self.max_order = J, tuple(shape), L, max_order
        self.pre_pad, self.backend, self.out_type = pre_pad, backend, out_type
        M, N = self.shape
        if 2 ** J > M or 2 ** J > N:
            raise RuntimeError("The smallest dimension should be larger than 2^J.")
        if pre_pad:
            raise NotImplementedError("wst_b200: pre_pad=True is not supported.")
        if out_type not in ("array", "list"):
            raise RuntimeError("The out_type must be one of 'array' or 'list'.")
        self._M_padded, self._N_padded = compute_padding(M, N, J)
        self._plan = None

    def _get_plan(self, device=None):
        if self._plan is None or (device is not None and self._plan.device != _device_index(device)):
            self._plan = get_plan(self.shape[0], self.shape[1], self.J, self.L, self.max_order, device)
        return self._plan

    def _meta(self):
        """kymatio's out_type='list' metadata, in output order."""
        J, L = self.J, self.L
        out = [{"j": (), "n": (), "theta": ()}]
        for n1 in range(J * L):
            out.append({"j": (n1 // L,), "n": (n1,), "theta": (n1 % L,)})
        if self.max_order >= 2:
            for n1 in range(J * L):
                for n2 in range(J * L):
                    if n2 // L > n1 // L:
                        out.append({"j": (n1 // L, n2 // L), "n": (n1, n2), "theta": (n1 % L, n2 % L)})
        return out


class ScatteringNumPy2D(_ScatteringBase2D):
    """Drop-in for `kymatio.numpy.Scattering2D`: ndarray[..., H, W] -> ndarray[..., K, h, w]."""

    def scattering(self, input):
        if not type(input) is np.ndarray:
            raise TypeError("The input should be a NumPy array.")
        if len(input.shape) < 2:
            raise RuntimeError("Input array must have at least two dimensions.")
        if input.shape[-1] != self.shape[-1] or input.shape[-2] != self.shape[-2]:
            raise RuntimeError("NumPy array must be of spatial size (%i,%i)." % (self.shape[0], self.shape[1]))
        batch_shape = input.shape[:-2]
        out_dtype = input.dtype if input.dtype in (np.float32, np.float64) else np.float32
        x = np.ascontiguousarray(input.reshape((-1, 1) + input.shape[-2:]), dtype=np.float32)
        plan = self._get_plan()
        xd = torch.from_numpy(x).to("cuda:%d" % plan.device)
        _, maps = plan.forward(xd, want_features=False, want_maps=True)
        S = maps[:, 0].cpu().numpy().astype(out_dtype, copy=False)
        S = S.reshape(batch_shape + S.shape[-3:])
        if self.out_type == "list":
            meta = self._meta()
            return [dict(coef=S[..., i, :, :], **meta[i]) for i in range(S.shape[-3])]
        return S

    __call__ = scattering


class ScatteringTorch2D(torch.nn.Module, _ScatteringBase2D):
    """Drop-in for `kymatio.torch.Scattering2D`: Tensor[..., H, W] -> Tensor[..., K, h, w].
    CPU tensors are moved to the GPU and the result is moved back (inference.py:250-257 feeds CPU tensors)."""

    def __init__(self, J, shape, L=8, max_order=2, pre_pad=False, backend=None, out_type="array"):
        torch.nn.Module.__init__(self)
        _ScatteringBase2D.__init__(self, J, shape, L, max_order, pre_pad, backend, out_type)

    def scattering(self, input):
        if not torch.is_tensor(input):
            raise TypeError("The input should be a PyTorch Tensor.")
        if len(input.shape) < 2:
            raise RuntimeError("Input tensor must have at least two dimensions.")
        if not input.is_contiguous():
            raise RuntimeError("Tensor must be contiguous.")
        if input.shape[-1] != self.shape[-1] or input.shape[-2] != self.shape[-2]:
            raise RuntimeError("Tensor must be of spatial size (%i,%i)." % (self.shape[0], self.shape[1]))
        batch_shape = input.shape[:-2]
        x = input.reshape((-1, 1) + tuple(input.shape[-2:]))
        dev = input.device if input.is_cuda else None
        plan = self._get_plan(dev)
        xd = x.to(device="cuda:%d" % plan.device, dtype=torch.float32).contiguous()
        with torch.no_grad():
            _, maps = plan.forward(xd, want_features=False, want_maps=True)
        S = maps[:, 0].to(device=input.device, dtype=input.dtype if input.dtype.is_floating_point else torch.float32)
        S = S.reshape(tuple(batch_shape) + tuple(S.shape[-3:]))
        if self.out_type == "list":
            meta = self._meta()
            return [dict(coef=S[..., i, :, :], **meta[i]) for i in range(S.shape[-3])]
        return S

    forward = scattering


def Scattering2D(J, shape, L=8, max_order=2, pre_pad=False, backend=None, out_type="array", frontend="numpy"):
    """Drop-in for `kymatio.Scattering2D(..., frontend=...)` (compare_wst_coefficients.py:37)."""
    if frontend == "numpy":
        return ScatteringNumPy2D(J, shape, L, max_order, pre_pad, backend, out_type)
    if frontend == "torch":
        return ScatteringTorch2D(J, shape, L, max_order, pre_pad, backend, out_type)
    raise RuntimeError("The frontend '%s' is not valid. Must be one of 'numpy' or 'torch'." % frontend)


# ----------------------------------------------------------------------------- advanced statistics (N3)
ADVANCED_STAT_NAMES = ["mean", "std", "var", "min", "max", "range", "skew", "kurt", "cv", "p10", "p25", "p50", "p75",
                       "p90", "iqr", "mad", "grad_mean", "edge_density"]        # train_and_save_model.py:402-405


def advanced_stats(x):
    """x: [B, C, H, W] float32 or [B, H, W, C] uint8 CUDA tensor -> [B, C, 18] float32 on the same device
    (extract_advanced_features, train_and_save_model.py:58-112, per patch and channel)."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 4 and x.is_contiguous()):
        raise RuntimeError("advanced_stats expects a contiguous 4-D CUDA tensor.")
    u8 = x.dtype == torch.uint8
    if not u8 and x.dtype != torch.float32:
        raise TypeError("advanced_stats expects float32 [B, C, H, W] or uint8 [B, H, W, C].")
    B, C, H, W = (x.shape[0], x.shape[3], x.shape[1], x.shape[2]) if u8 else tuple(x.shape)
    out = torch.empty((B, C, 18), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    rc = lib.wst2d_advanced_stats(x.device.index or 0, x.data_ptr(), 1 if u8 else 0, B, C, H, W, out.data_ptr(),
                                  ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    if rc != 0:
        msg = lib.wst2d_advanced_stats_last_error().decode()
        raise (NotImplementedError if rc == _lib.WST2D_ERR_UNSUPPORTED else RuntimeError)("wst_b200: " + msg)
    return out


def _rgb_to_device(rgb_image, channels=None):
    """Host [C, H, W] image -> float32 CUDA tensor [1, C', H, W]; rejects non-finite pixels.

    The reference drops non-finite pixels before the moments and percentiles (train_and_save_model.py:66-67) and
    lets them poison the Sobel / Laplace statistics; its inputs are uint8 PNGs / 255 (load_rgb_image, :51-56) and
    therefore always finite.  The device kernels assume finite input, so anything else is refused here instead of
    returning statistics the reference would not produce."""
    if not torch.cuda.is_available():
        raise RuntimeError("wst_b200: no CUDA device available (this package has no CPU fallback).")
    a = np.asarray(rgb_image)
    if a.ndim != 3:
        raise ValueError("rgb_image must be [C, H, W]")
    if channels is not None:
        if a.shape[0] < channels:
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (a.shape[0], a.shape[0]))
        a = a[:channels]
    a = np.ascontiguousarray(a[None], dtype=np.float32)
    if not np.isfinite(a).all():
        raise ValueError("wst_b200: non-finite pixels are not supported (the reference's inputs are uint8 / 255).")
    return torch.from_numpy(a).cuda()


def extract_advanced_features(rgb_image):
    """Drop-in for train_and_save_model.py:58-112 / inference.py:181-235: [C, H, W] float32 -> float64 [3*18].
    Like the reference, the first three channels are used whatever C is (C < 3 raises IndexError there too)."""
    return advanced_stats(_rgb_to_device(rgb_image, 3))[0].reshape(-1).cpu().numpy().astype(np.float64)


def basic_stats(x):
    """x: [B, C, H, W] float32 or [B, H, W, C] uint8 CUDA tensor -> [B, C, 2] (mean, population std) per channel:
    the batched device form of ModelInference.extract_basic_features (inference.py:170-179).  The two numbers are
    the first two of the advanced-statistics kernel's outputs (double-precision moments)."""
    return advanced_stats(x)[..., :2].contiguous()


def extract_basic_features(rgb_image):
    """Drop-in for ModelInference.extract_basic_features (inference.py:170-179): [C, H, W] float32 ->
    float64 [6] = (mean, std) of each of the first three channels."""
    return basic_stats(_rgb_to_device(rgb_image, 3))[0].reshape(-1).cpu().numpy().astype(np.float64)


def extract_hybrid_features(rgb_image, J=2, L=8):
    """Drop-in for train_and_save_model.py:380-387: concat([advanced stats (3*18), WST (C*2*K)]) -> float64."""
    return np.concatenate([extract_advanced_features(rgb_image), extract_wst_features(rgb_image, J=J, L=L)])


def extract_features(rgb_image, feature_method):
    """Drop-in for the training dispatcher, train_and_save_model.py:389-398."""
    if feature_method == "advanced_stats":
        return extract_advanced_features(rgb_image)
    elif feature_method == "wst":
        return extract_wst_features(rgb_image)
    elif feature_method == "hybrid":
        return extract_hybrid_features(rgb_image)
    else:
        raise ValueError(f"Unknown feature method: {feature_method}")


def get_feature_names(feature_method, K=81):
    """Drop-in for train_and_save_model.py:400-427: the column names the Random-Forest trainer zips with the
    feature matrix.  K = 81 is the reference's hard-coded coefficient count (J=2, L=8); other plans pass their K."""
    if feature_method == "advanced_stats":
        return [f"{c}_{stat}" for c in ["R", "G", "B"] for stat in ADVANCED_STAT_NAMES]
    elif feature_method == "wst":
        return [f"{channel}_wst_{stat}_{i}" for channel in ["R", "G", "B"] for stat in ["mean", "std"]
                for i in range(K)]
    elif feature_method == "hybrid":
        return get_feature_names("advanced_stats") + get_feature_names("wst", K)
    else:
        raise ValueError(f"Unknown feature method: {feature_method}")


def extract_features_inference(rgb_image, feature_method, J=2, L=8):
    """Drop-in for ModelInference.extract_features (inference.py:272-287): 'advanced_stats' -> 54,
    'wst' -> hstack([basic(6), interleaved WST]) = 492, 'hybrid' -> hstack([advanced(54), interleaved WST]) = 540
    (float64, the inference driver's column order — SURVEY.md F4)."""
    if feature_method == "advanced_stats":
        return extract_advanced_features(rgb_image)
    elif feature_method == "wst":
        return np.hstack([extract_basic_features(rgb_image), extract_wst_features_interleaved(rgb_image, J=J, L=L)])
    elif feature_method == "hybrid":
        return np.hstack([extract_advanced_features(rgb_image), extract_wst_features_interleaved(rgb_image, J=J, L=L)])
    else:
        raise ValueError(f"Unknown feature method: {feature_method}")


class ModelInferenceFeatures:
    """The feature-extraction methods of the reference's ModelInference (inference.py:170-287) backed by the CUDA
    library, same names and signatures, so that `class ModelInference(ModelInferenceFeatures)` (or assigning the
    methods) switches the inference driver over without touching predict_single_image (:289-320).
    `self.feature_method` is what parse_model_directory (:61-124) sets."""

    feature_method = "wst"

    def extract_basic_features(self, rgb_image):
        return extract_basic_features(rgb_image)

    def extract_advanced_features(self, rgb_image):
        return extract_advanced_features(rgb_image)

    def extract_wst_features(self, rgb_image, J=2, L=8):
        return extract_wst_features_interleaved(rgb_image, J=J, L=L)

    def extract_features(self, rgb_image):
        return extract_features_inference(rgb_image, self.feature_method)


def hybrid_features(x, J, L=8, max_order=2):
    """Batched device form: x [B, C, H, W] float32 (or [B, H, W, C] uint8) CUDA -> [B, C*18 + C*2*K] (advanced stats
    block, then WST block)."""
    adv = advanced_stats(x).reshape(x.shape[0], -1)
    return torch.cat([adv, scattering_features(x, J, L, max_order)], dim=1)


# ----------------------------------------------------------------------------- noise models (N4)
NOISE_TYPES = ["gaussian", "salt_and_pepper", "speckle", "poisson", "uniform"]       # add_noise.py:123 / wst2d.h enum


def add_noise(images, noise_type, intensity, seed=42, draws=None):
    """Batched device form of add_noise.py:14-72.  images: contiguous uint8 CUDA tensor [B, H, W, C] (or [H, W, C]);
    returns a new uint8 tensor of the same shape, ready for the uint8 ingest of the WST plan.

    draws=None: the random draws come from the library's counter-based generator (seed defaults to the
    reference's --seed 42; the stream differs from numpy's).  draws given: the reference's own numpy draws for the
    batch — a float64 tensor shaped like `images` (gaussian: already scaled by sigma; speckle; uniform), an int64
    tensor of Poisson counts, or for salt_and_pepper an int64 tensor [B, 2 (salt, pepper), 2 (row, col), n] — and
    the result is then bit-identical to the reference's."""
    if noise_type not in NOISE_TYPES:
        raise ValueError("Unknown noise type: %s" % noise_type)                      # add_noise.py:92
    if not (isinstance(images, torch.Tensor) and images.is_cuda and images.dtype == torch.uint8
            and images.dim() in (3, 4) and images.is_contiguous()):
        raise RuntimeError("add_noise expects a contiguous uint8 CUDA tensor [B, H, W, C] or [H, W, C].")
    x = images if images.dim() == 4 else images[None]
    B, H, W, C = x.shape
    out = torch.empty_like(x)
    lib = _lib.load()
    kind = NOISE_TYPES.index(noise_type)
    stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    dev = x.device.index or 0
    if draws is None:
        rc = lib.wst2d_add_noise(dev, kind, float(intensity), x.data_ptr(), B, H, W, C, int(seed) & (2 ** 64 - 1),
                                 out.data_ptr(), stream)
    else:
        want = torch.int64 if noise_type in ("poisson", "salt_and_pepper") else torch.float64
        if not (isinstance(draws, torch.Tensor) and draws.is_cuda and draws.dtype == want and draws.is_contiguous()):
            raise RuntimeError("add_noise: draws must be a contiguous %s CUDA tensor." % want)
        if noise_type == "salt_and_pepper":
            if draws.dim() != 4 or tuple(draws.shape[:3]) != (B, 2, 2):
                raise RuntimeError("add_noise: salt_and_pepper draws must be [B, 2, 2, n].")
            n = draws.shape[3]
        else:
            if draws.numel() != x.numel():
                raise RuntimeError("add_noise: draws must have one value per image element.")
            n = 0
        rc = lib.wst2d_add_noise_draws(dev, kind, float(intensity), x.data_ptr(), B, H, W, C, draws.data_ptr(), n,
                                       out.data_ptr(), stream)
    if rc != 0:
        msg = lib.wst2d_noise_last_error().decode()
        raise (ValueError if rc == _lib.WST2D_ERR_ARG else RuntimeError)("wst_b200: " + msg)
    return out if images.dim() == 4 else out[0]


def _add_noise_image(image_array, noise_type, intensity):
    if not torch.cuda.is_available():
        raise RuntimeError("wst_b200: no CUDA device available (this package has no CPU fallback).")
    a = np.ascontiguousarray(image_array)
    if a.dtype != np.uint8 or a.ndim != 3:
        raise ValueError("expected a uint8 image array [H, W, C]")
    seed = int(np.random.randint(0, 2 ** 31 - 1))       # one draw from the caller's seeded numpy stream (add_noise.py:147-149)
    return add_noise(torch.from_numpy(a).cuda(), noise_type, intensity, seed=seed).cpu().numpy()


def add_gaussian_noise(image_array, intensity):
    """Drop-in for add_noise.py:14-21 (uint8 [H, W, C] numpy in and out; seeded through numpy's global RNG)."""
    return _add_noise_image(image_array, "gaussian", intensity)


def add_salt_and_pepper_noise(image_array, intensity):
    """Drop-in for add_noise.py:23-43, including its coordinate-range and count conventions."""
    return _add_noise_image(image_array, "salt_and_pepper", intensity)


def add_speckle_noise(image_array, intensity):
    """Drop-in for add_noise.py:45-54."""
    return _add_noise_image(image_array, "speckle", intensity)


def add_poisson_noise(image_array, intensity):
    """Drop-in for add_noise.py:56-65."""
    return _add_noise_image(image_array, "poisson", intensity)


def add_uniform_noise(image_array, intensity):
    """Drop-in for add_noise.py:67-72."""
    return _add_noise_image(image_array, "uniform", intensity)


# ----------------------------------------------------------------------------- reference extractors (B1, B2)
def extract_wst_features(rgb_image, J=2, L=8, max_order=2):
    """Drop-in for train_and_save_model.py:346-378: float32 [C, H, W] -> float32 [C*2*K],
    per channel [mean(K) || std(K)], channels concatenated.  J=2, L=8 are the reference's constants."""
    rgb_image = np.asarray(rgb_image)
    if rgb_image.ndim != 3:
        raise ValueError("rgb_image must be [C, H, W]")
    x = np.ascontiguousarray(rgb_image[None], dtype=np.float32)
    plan = get_plan(x.shape[-2], x.shape[-1], J, L, max_order)
    feats = plan.forward_host(x)                       # [1, C, 2, K]
    return feats.reshape(-1)


def extract_wst_features_interleaved(rgb_image, J=2, L=8):
    """Drop-in for inference.py:237-270: float32 [C, H, W] -> float64 [C*2*K],
    per channel [mean0, std0, mean1, std1, ...]."""
    rgb_image = np.asarray(rgb_image)
    x = np.ascontiguousarray(rgb_image[None], dtype=np.float32)
    plan = get_plan(x.shape[-2], x.shape[-1], J, L, 2)
    feats = plan.forward_host(x)[0]                    # [C, 2, K]
    return np.ascontiguousarray(feats.swapaxes(-1, -2)).reshape(-1).astype(np.float64)


def extract_wst_features_gray(grayscale_image, J=2, L=8):
    """Drop-in for visualize_features.py:194-222: [H, W] -> (features [2K], coefficient maps [K, h, w]),
    dtype following the input like kymatio (float64 in -> float64 out)."""
    g = np.asarray(grayscale_image)
    out_dtype = g.dtype if g.dtype in (np.float32, np.float64) else np.float32
    H, W = g.shape
    plan = get_plan(H, W, J, L, 2)
    xd = torch.from_numpy(np.ascontiguousarray(g[None, None], dtype=np.float32)).to("cuda:%d" % plan.device)
    feats, maps = plan.forward(xd, want_features=True, want_maps=True)
    return (feats[0, 0].reshape(-1).cpu().numpy().astype(out_dtype),
            maps[0, 0].cpu().numpy().astype(out_dtype))


def compute_scattering_coefficients(img_tensor, L=6, J=3):
    """Drop-in for compare_wst_coefficients.py:35-39 (note the sign flip kept from the reference)."""
    scattering = Scattering2D(J=J, shape=img_tensor.shape, L=L, max_order=2, frontend="numpy")
    return -scattering(img_tensor)


def extract_wst_features_batch(images, J=2, L=8, max_order=2, layout="block"):
    """Batched form of extract_wst_features: [B, C, H, W] host array -> [B, C*2*K] float32,
    row b identical to extract_wst_features(images[b]) (the per-image loop at
    train_and_save_model.py:486-490 collapses into one call)."""
    return features_from_host(images, J, L, max_order, layout)
