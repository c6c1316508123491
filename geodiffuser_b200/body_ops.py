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
            and x.shape[1] <= 2560 and x.shape[0] <= 64 and x.is_contiguous(memory_format=torch.channels_last)
            and norm.weight is not None and norm.bias is not None and norm.weight.dtype in (torch.bfloat16, torch.float32)
            and not (torch.is_grad_enabled() and (norm.weight.requires_grad or norm.bias.requires_grad)))


_COUNTERS = {}


def _workspace(B, HW, C, G, dev):
    n = _lib.lib().gd_group_norm_nhwc_workspace(B, HW, C, G)
    return torch.empty(n, device=dev, dtype=torch.float32), n


def _counters(dev):
    """per-(device, stream) arrival counters of the reduction: zero between launches (the kernel that uses them clears them again)"""
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    c = _COUNTERS.get(key)
    if c is None:
        c = _COUNTERS[key] = torch.zeros(64, device=dev, dtype=torch.int32)
    return c


class _GroupNormActNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pre_bias, weight, bias, groups, eps, silu):
        B, C, H, W = x.shape
        y = torch.empty_like(x)                      # preserves channels_last
        stats = torch.empty(B, groups, 2, device=x.device, dtype=torch.float32)
        ws, n = _workspace(B, H * W, C, groups, x.device)
        call("gd_group_norm_nhwc_fwd", ptr_cl(x), ptr(pre_bias), ptr(weight), ptr(bias), int(weight.dtype == torch.bfloat16), B, H * W, C, groups,
             float(eps), int(silu), ptr(ws), n, ptr(_counters(x.device)), ptr(stats), ptr_cl(y), stream())
        ctx.save_for_backward(x, pre_bias, weight, bias, stats)
        ctx.groups, ctx.silu = groups, silu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pre_bias, weight, bias, stats = ctx.saved_tensors
        B, C, H, W = x.shape
        dy = dy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        ws, n = _workspace(B, H * W, C, ctx.groups, x.device)
        call("gd_group_norm_nhwc_bwd", ptr_cl(x), ptr(pre_bias), ptr_cl(dy), ptr(weight), ptr(bias), int(weight.dtype == torch.bfloat16), ptr(stats),
             B, H * W, C, ctx.groups, int(ctx.silu), ptr(ws), n, ptr(_counters(x.device)), ptr_cl(dx), stream())
        return dx, None, None, None, None, None, None


def ptr_cl(t):
    """device pointer of a channels-last-contiguous (B, C, H, W) tensor, i.e. of its (B, HW, C) memory"""
    import ctypes

    if not (t.is_cuda and t.is_contiguous(memory_format=torch.channels_last)):
        raise _lib.GeoDiffuserB200Error("expected a channels_last CUDA tensor")
    return ctypes.c_void_p(t.data_ptr())


def group_norm_act(norm, x, silu=False, pre_bias=None):
    """silu?(norm(x + pre_bias[:, :, None, None])) for an nn.GroupNorm `norm`; pre_bias (B, C) or None"""
    if _eligible(x, norm) and (pre_bias is None or (pre_bias.dtype == torch.bfloat16 and not pre_bias.requires_grad)):
        pb = None if pre_bias is None else pre_bias.contiguous()
        return _GroupNormActNHWC.apply(x, pb, norm.weight, norm.bias, norm.num_groups, norm.eps, silu)
    if pre_bias is not None:
        x = x + pre_bias[:, :, None, None]
    y = norm(x)
    return F.silu(y) if silu else y


def fast_body(x):
    """the fused body ops serve channels-last bf16 CUDA activations (the product setting); anything else takes the stock torch route"""
    return ENABLED and x.is_cuda and x.dtype == torch.bfloat16


class _GEGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proj):
        rows, F2 = proj.numel() // proj.shape[-1], proj.shape[-1]
        out = torch.empty(*proj.shape[:-1], F2 // 2, device=proj.device, dtype=proj.dtype)
        call("gd_geglu_fwd", ptr(proj), rows, F2 // 2, ptr(out), stream())
        ctx.save_for_backward(proj)
        return out

    @staticmethod
    def backward(ctx, dy):
        (proj,) = ctx.saved_tensors
        rows, F2 = proj.numel() // proj.shape[-1], proj.shape[-1]
        dproj = torch.empty_like(proj)
        call("gd_geglu_bwd", ptr(proj), ptr(dy.to(torch.bfloat16).contiguous()), rows, F2 // 2, ptr(dproj), stream())
        return dproj


def geglu(proj):
    """proj[..., :F] * gelu(proj[..., F:]) (diffusers GEGLU)"""
    if fast_body(proj) and proj.is_contiguous() and proj.shape[-1] % 16 == 0:
        return _GEGLU.apply(proj)
    a, g = proj.chunk(2, dim=-1)
    return a * F.gelu(g)


class _AddBiasResidual(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, bias):
        B, C, H, W = a.shape
        out = torch.empty_like(a)
        call("gd_add_bias_residual", ptr_cl(a), ptr_cl(b), ptr(bias), B * H * W, C, ptr_cl(out), stream())
        return out

    @staticmethod
    def backward(ctx, dy):
        return dy, dy, None


def add_bias_residual(a, b, bias):
    """a + (b + bias[:, None, None]) for channels-last (B, C, H, W) activations: a residual add with the producing convolution's bias folded in"""
    cl = torch.channels_last
    if (fast_body(a) and a.dim() == 4 and b.dtype == torch.bfloat16 and a.shape == b.shape and a.shape[1] % 8 == 0 and bias.dtype == torch.bfloat16
            and a.is_contiguous(memory_format=cl) and b.is_contiguous(memory_format=cl) and not bias.requires_grad):
        return _AddBiasResidual.apply(a, b, bias)
    return a + (b + bias[:, None, None])


def conv1x1(conv, x):
    """a 1x1 convolution of a channels-last activation is a GEMM over its (B*HW, C) memory: cuBLASLt adds the bias in the epilogue, where cuDNN's
    convolution is followed by a separate broadcast add"""
    if (fast_body(x) and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and conv.kernel_size == (1, 1)
            and conv.stride == (1, 1) and conv.padding == (0, 0)):
        B, C, H, W = x.shape
        w = conv.weight.reshape(conv.out_channels, C)      # (O, C, 1, 1): a view in either memory format
        y = F.linear(x.permute(0, 2, 3, 1).reshape(B * H * W, C), w, conv.bias)
        return y.reshape(B, H, W, conv.out_channels).permute(0, 3, 1, 2)
    return conv(x)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        C = x.shape[-1]
        rows = x.numel() // C
        y = torch.empty_like(x)
        mean = torch.empty(*x.shape[:-1], 1, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        call("gd_layer_norm_fwd", ptr(x), ptr(weight), ptr(bias), rows, C, float(eps), ptr(y), ptr(mean), ptr(rstd), stream())
        ctx.save_for_backward(x, weight, bias, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, mean, rstd = ctx.saved_tensors
        dx, _, _ = torch.ops.aten.native_layer_norm_backward(dy.contiguous(), x, [x.shape[-1]], mean, rstd, weight, bias, [True, False, False])
        return dx, None, None, None


def layer_norm(norm, x):
    """norm(x) for an nn.LayerNorm over the last dimension"""
    if (fast_body(x) and x.is_contiguous() and len(norm.normalized_shape) == 1 and x.shape[-1] % 8 == 0 and x.shape[-1] <= 1280
            and norm.weight is not None and norm.bias is not None and norm.weight.dtype == torch.bfloat16
            and not (norm.weight.requires_grad or norm.bias.requires_grad)):
        return _LayerNorm.apply(x, norm.weight, norm.bias, norm.eps)
    return norm(x)
