"""torch.library registration of the batched device ops (surface B4).

    torch.ops.wst.scattering2d_features(x, J, L, max_order, layout, full_maps) -> Tensor
    torch.ops.wst.scattering2d_maps(x, J, L, max_order)                        -> Tensor

x is [B, C, H, W] float32 (or [B, H, W, C] uint8, load_rgb_image's input order) on a CUDA device.
`layout` 0 = per channel [mean(K) || std(K)] (train_and_save_model.py:375), 1 = interleaved
[mean0, std0, ...] (inference.py:263-266).  With full_maps the first op returns the coefficient maps
[B, C, K, h, w] instead of pooled features [B, C*2*K] (the form visualize_features.py:213-222 needs).

Only a CUDA kernel is registered: a CPU tensor raises NotImplementedError from the dispatcher (there is no CPU
fallback).  The fake (meta) kernels give shape / dtype propagation for FakeTensor tracing and torch.compile;
the ops are opaque to autograd (the reference runs the transform under torch.no_grad(), inference.py:253).
"""
import torch

from ._api import get_plan, num_coefficients, compute_padding, to_block, to_interleaved

__all__ = ["scattering2d_features", "scattering2d_maps"]


def _geometry(x, J, L, max_order):
    if x.dim() != 4:
        raise RuntimeError("Input tensor must be [B, C, H, W] (float32) or [B, H, W, C] (uint8).")
    if x.dtype == torch.uint8:
        B, H, W, C = x.shape
    else:
        B, C, H, W = x.shape
    if 2 ** J > H or 2 ** J > W:
        raise RuntimeError("The smallest dimension should be larger than 2^J.")
    Hp, Wp = compute_padding(H, W, J)
    return B, C, H, W, num_coefficients(J, L, max_order), Hp // 2 ** J - 2, Wp // 2 ** J - 2


@torch.library.custom_op("wst::scattering2d_features", mutates_args=(), device_types="cuda")
def scattering2d_features(x: torch.Tensor, J: int, L: int, max_order: int, layout: int, full_maps: bool) -> torch.Tensor:
    B, C, H, W, K, h, w = _geometry(x, J, L, max_order)
    plan = get_plan(H, W, J, L, max_order, x.device)
    if full_maps:
        return plan.forward(x, want_features=False, want_maps=True)[1]
    feats = plan.forward(x, want_features=True)[0]
    return (to_block(feats) if layout == 0 else to_interleaved(feats)).contiguous()


@scattering2d_features.register_fake
def _(x, J, L, max_order, layout, full_maps):
    B, C, H, W, K, h, w = _geometry(x, J, L, max_order)
    if full_maps:
        return x.new_empty((B, C, K, h, w), dtype=torch.float32)
    return x.new_empty((B, C * 2 * K), dtype=torch.float32)


@torch.library.custom_op("wst::scattering2d_maps", mutates_args=(), device_types="cuda")
def scattering2d_maps(x: torch.Tensor, J: int, L: int, max_order: int) -> torch.Tensor:
    B, C, H, W, K, h, w = _geometry(x, J, L, max_order)
    return get_plan(H, W, J, L, max_order, x.device).forward(x, want_features=False, want_maps=True)[1]


@scattering2d_maps.register_fake
def _(x, J, L, max_order):
    B, C, H, W, K, h, w = _geometry(x, J, L, max_order)
    return x.new_empty((B, C, K, h, w), dtype=torch.float32)
