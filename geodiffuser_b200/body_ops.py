"""Caller-side fused ops for the UNet body (unet_sd15.py) -- NOT part of the reference's attention-controller surface.

The body is the caller of the hot path and otherwise stock torch, exactly as under the reference.  One exception, because it hid the
path: torch's CUDA `group_norm` has no channels-last kernel, so every one of the 61 GroupNorms of a UNet evaluation cost two layout
copies plus four kernels (27 % of the device time of a gradient-free pass, profiles/r01c_phase_kernels.md).  `group_norm_act` runs
GroupNorm (+ SiLU) in two launches of csrc/body_norm.cu on the channels-last bf16 activation, forward and input-gradient; any other
input (fp32 parity runs, NCHW, CPU, C % 8 != 0, weights that require grad) goes through stock torch unchanged.
"""
import torch
import torch.nn.functional as F

from . import _lib
from ._lib import call, ptr, stream

ENABLED = True


def _eligible(x, norm):
    return (ENABLED and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0 and norm.num_groups <= 32
            and x.shape[1] <= 2560 and x.is_contiguous(memory_format=torch.channels_last)
            and norm.weight is not None and norm.bias is not None and norm.weight.dtype in (torch.bfloat16, torch.float32)
            and not (torch.is_grad_enabled() and (norm.weight.requires_grad or norm.bias.requires_grad)))


def _workspace(B, HW, C, G, dev):
    n = _lib.lib().gd_group_norm_nhwc_workspace(B, HW, C, G)
    return torch.empty(n, device=dev, dtype=torch.float32), n


class _GroupNormActNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps, silu):
        B, C, H, W = x.shape
        y = torch.empty_like(x)                      # preserves channels_last
        stats = torch.empty(B, groups, 2, device=x.device, dtype=torch.float32)
        ws, n = _workspace(B, H * W, C, groups, x.device)
        call("gd_group_norm_nhwc_fwd", ptr_cl(x), ptr(weight), ptr(bias), int(weight.dtype == torch.bfloat16), B, H * W, C, groups, float(eps),
             int(silu), ptr(ws), n, ptr(stats), ptr_cl(y), stream())
        ctx.save_for_backward(x, weight, bias, stats)
        ctx.groups, ctx.silu = groups, silu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, stats = ctx.saved_tensors
        B, C, H, W = x.shape
        dy = dy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        ws, n = _workspace(B, H * W, C, ctx.groups, x.device)
        call("gd_group_norm_nhwc_bwd", ptr_cl(x), ptr_cl(dy), ptr(weight), ptr(bias), int(weight.dtype == torch.bfloat16), ptr(stats), B, H * W, C,
             ctx.groups, int(ctx.silu), ptr(ws), n, ptr_cl(dx), stream())
        return dx, None, None, None, None, None


def ptr_cl(t):
    """device pointer of a channels-last-contiguous (B, C, H, W) tensor, i.e. of its (B, HW, C) memory"""
    import ctypes

    if not (t.is_cuda and t.is_contiguous(memory_format=torch.channels_last)):
        raise _lib.GeoDiffuserB200Error("expected a channels_last CUDA tensor")
    return ctypes.c_void_p(t.data_ptr())


def group_norm_act(norm, x, silu=False):
    """silu?(norm(x)) for an nn.GroupNorm `norm`"""
    if _eligible(x, norm):
        return _GroupNormActNHWC.apply(x, norm.weight, norm.bias, norm.num_groups, norm.eps, silu)
    y = norm(x)
    return F.silu(y) if silu else y
