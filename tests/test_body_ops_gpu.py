"""Caller-side fused GroupNorm(+SiLU) on channels-last bf16 (csrc/body_norm.cu) against torch's fp32 group_norm on the same bf16 values:
forward within bf16 rounding, input gradient within the path's 2e-2, bit-identical from run to run, stock fallback for ineligible inputs."""
import pytest
import torch
import torch.nn.functional as F

from conftest import relerr

pytestmark = pytest.mark.gpu

SHAPES = [(2, 320, 64, 64), (3, 640, 32, 32), (2, 1280, 16, 16), (2, 1280, 8, 8), (3, 2560, 16, 16), (2, 1920, 32, 32), (2, 960, 64, 64),
          (2, 64, 8, 8), (1, 384, 12, 12), (2, 128, 3, 5)]


@pytest.fixture(params=[1, 0], ids=["cluster", "two-launch"])
def gn_scheme(request):
    """forward scheme of gd_group_norm_nhwc_fwd: one-launch cluster / DSMEM kernel (default) or the two-launch scheme"""
    from geodiffuser_b200 import _lib

    _lib.call("gd_group_norm_config", request.param)
    yield request.param
    _lib.call("gd_group_norm_config", 1)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("silu,shift", [(False, False), (True, False), (True, True)])
def test_group_norm_nhwc_forward_and_backward(shape, silu, shift, gn_scheme):
    from geodiffuser_b200 import body_ops

    B, C, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = (torch.randn(shape, device="cuda", generator=g) * 2.0 + 0.7).bfloat16().contiguous(memory_format=torch.channels_last)
    pb = (torch.randn(B, C, device="cuda", generator=g) * 0.8).bfloat16() if shift else None
    norm = torch.nn.GroupNorm(32, C, eps=1e-5).cuda()
    with torch.no_grad():
        norm.weight.copy_(torch.randn(C, device="cuda", generator=g) * 0.5 + 1.0)
        norm.bias.copy_(torch.randn(C, device="cuda", generator=g) * 0.3)
    norm = norm.bfloat16().requires_grad_(False)
    dy = torch.randn(shape, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)

    xr = x.float().requires_grad_(True)
    ref = F.group_norm(xr if pb is None else xr + pb.float()[:, :, None, None], 32, norm.weight.float(), norm.bias.float(), 1e-5)
    if silu:
        ref = F.silu(ref)
    (dx_ref,) = torch.autograd.grad(ref, xr, dy.float())

    outs = []
    for _ in range(2):
        xi = x.clone().requires_grad_(True)
        y = body_ops.group_norm_act(norm, xi, silu=silu, pre_bias=pb)
        assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
        (dx,) = torch.autograd.grad(y, xi, dy)
        outs.append((y.detach(), dx))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    y, dx = outs[0]
    assert relerr(y.float().cpu().numpy(), ref.detach().cpu().numpy()) <= 8e-3          # one bf16 rounding of the result
    assert relerr(dx.float().cpu().numpy(), dx_ref.cpu().numpy()) <= 2e-2
    # and against what the stock path (bf16 group_norm -> bf16 silu) returns
    stock = norm(x if pb is None else x + pb[:, :, None, None])
    stock = F.silu(stock) if silu else stock
    assert relerr(y.float().cpu().numpy(), stock.float().cpu().numpy()) <= 2e-2


def test_group_norm_falls_back_to_stock_for_ineligible_inputs():
    from geodiffuser_b200 import body_ops

    norm = torch.nn.GroupNorm(32, 64).cuda()
    x = torch.randn(2, 64, 8, 8, device="cuda")                      # fp32 (parity setting): stock torch
    assert torch.equal(body_ops.group_norm_act(norm, x, silu=True), F.silu(norm(x)))
    xb = torch.randn(2, 64, 8, 8, device="cuda").bfloat16()          # bf16 but NCHW
    nb = torch.nn.GroupNorm(32, 64).cuda().bfloat16()
    assert torch.equal(body_ops.group_norm_act(nb, xb), nb(xb))


@pytest.mark.parametrize("rows,Fd", [(2 * 4096, 1280), (3 * 1024, 2560), (2 * 64, 5120), (7, 16)])
def test_geglu_forward_and_backward(rows, Fd):
    from geodiffuser_b200 import body_ops

    g = torch.Generator(device="cuda").manual_seed(rows + Fd)
    proj = (torch.randn(rows, 2 * Fd, device="cuda", generator=g) * 1.5).bfloat16()
    dy = torch.randn(rows, Fd, device="cuda", generator=g).bfloat16()
    pr = proj.float().requires_grad_(True)
    a, gate = pr.chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    (dref,) = torch.autograd.grad(ref, pr, dy.float())
    pi = proj.clone().requires_grad_(True)
    out = body_ops.geglu(pi)
    (dp,) = torch.autograd.grad(out, pi, dy)
    assert out.dtype == torch.bfloat16 and out.shape == (rows, Fd)
    assert relerr(out.detach().float().cpu().numpy(), ref.detach().cpu().numpy()) <= 8e-3
    assert relerr(dp.float().cpu().numpy(), dref.cpu().numpy()) <= 8e-3
    a2, g2 = proj.chunk(2, dim=-1)
    assert relerr(out.detach().float().cpu().numpy(), (a2 * F.gelu(g2)).float().cpu().numpy()) <= 1.6e-2


def test_bias_residual_and_conv1x1_match_stock():
    from geodiffuser_b200 import body_ops

    g = torch.Generator(device="cuda").manual_seed(5)
    cl = torch.channels_last
    a = torch.randn(2, 320, 16, 16, device="cuda", generator=g).bfloat16().contiguous(memory_format=cl).requires_grad_(True)
    b = torch.randn(2, 320, 16, 16, device="cuda", generator=g).bfloat16().contiguous(memory_format=cl).requires_grad_(True)
    bias = torch.randn(320, device="cuda", generator=g).bfloat16()
    out = body_ops.add_bias_residual(a, b, bias)
    ref = a.float() + b.float() + bias.float()[:, None, None]
    assert out.is_contiguous(memory_format=cl)
    assert relerr(out.detach().float().cpu().numpy(), ref.detach().cpu().numpy()) <= 8e-3
    ga, gb = torch.autograd.grad(out, [a, b], torch.ones_like(out))
    assert float(ga.float().min()) == 1.0 and float(gb.float().max()) == 1.0
    conv = torch.nn.Conv2d(320, 640, 1).cuda().bfloat16().to(memory_format=cl).requires_grad_(False)
    x = a.detach()
    y = body_ops.conv1x1(conv, x)
    assert y.shape == (2, 640, 16, 16) and y.is_contiguous(memory_format=cl)
    assert relerr(y.float().cpu().numpy(), conv(x).float().cpu().numpy()) <= 1.6e-2


@pytest.mark.parametrize("shape", [(2, 4096, 320), (3, 1024, 640), (2, 256, 1280), (2, 64, 1280), (5, 7, 64), (1, 3, 136)])
def test_layer_norm_forward_and_backward(shape):
    from geodiffuser_b200 import body_ops

    C = shape[-1]
    g = torch.Generator(device="cuda").manual_seed(C)
    x = (torch.randn(shape, device="cuda", generator=g) * 3.0 + 1.5).bfloat16()
    norm = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        norm.weight.copy_(torch.randn(C, device="cuda", generator=g) * 0.5 + 1.0)
        norm.bias.copy_(torch.randn(C, device="cuda", generator=g) * 0.3)
    norm = norm.bfloat16().requires_grad_(False)
    dy = torch.randn(shape, device="cuda", generator=g).bfloat16()
    xr = x.float().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), norm.weight.float(), norm.bias.float(), norm.eps)
    (dref,) = torch.autograd.grad(ref, xr, dy.float())
    xi = x.clone().requires_grad_(True)
    y = body_ops.layer_norm(norm, xi)
    (dx,) = torch.autograd.grad(y, xi, dy)
    assert y.dtype == torch.bfloat16
    assert relerr(y.detach().float().cpu().numpy(), ref.detach().cpu().numpy()) <= 8e-3
    assert relerr(dx.float().cpu().numpy(), dref.cpu().numpy()) <= 2e-2
    assert relerr(y.detach().float().cpu().numpy(), norm(x).float().cpu().numpy()) <= 1.6e-2
