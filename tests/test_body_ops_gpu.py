"""Caller-side fused GroupNorm(+SiLU) on channels-last bf16 (csrc/body_norm.cu) against torch's fp32 group_norm on the same bf16 values:
forward within bf16 rounding, input gradient within the path's 2e-2, bit-identical from run to run, stock fallback for ineligible inputs."""
import pytest
import torch
import torch.nn.functional as F

from conftest import relerr

pytestmark = pytest.mark.gpu

SHAPES = [(2, 320, 64, 64), (3, 640, 32, 32), (2, 1280, 16, 16), (2, 1280, 8, 8), (3, 2560, 16, 16), (2, 1920, 32, 32), (2, 960, 64, 64),
          (2, 64, 8, 8), (1, 384, 12, 12), (2, 128, 3, 5)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("silu", [False, True])
def test_group_norm_nhwc_forward_and_backward(shape, silu):
    from geodiffuser_b200 import body_ops

    B, C, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = (torch.randn(shape, device="cuda", generator=g) * 2.0 + 0.7).bfloat16().contiguous(memory_format=torch.channels_last)
    norm = torch.nn.GroupNorm(32, C, eps=1e-5).cuda()
    with torch.no_grad():
        norm.weight.copy_(torch.randn(C, device="cuda", generator=g) * 0.5 + 1.0)
        norm.bias.copy_(torch.randn(C, device="cuda", generator=g) * 0.3)
    norm = norm.bfloat16().requires_grad_(False)
    dy = torch.randn(shape, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)

    xr = x.float().requires_grad_(True)
    ref = F.group_norm(xr, 32, norm.weight.float(), norm.bias.float(), 1e-5)
    if silu:
        ref = F.silu(ref)
    (dx_ref,) = torch.autograd.grad(ref, xr, dy.float())

    outs = []
    for _ in range(2):
        xi = x.clone().requires_grad_(True)
        y = body_ops.group_norm_act(norm, xi, silu=silu)
        assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
        (dx,) = torch.autograd.grad(y, xi, dy)
        outs.append((y.detach(), dx))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    y, dx = outs[0]
    assert relerr(y.float().cpu().numpy(), ref.detach().cpu().numpy()) <= 8e-3          # one bf16 rounding of the result
    assert relerr(dx.float().cpu().numpy(), dx_ref.cpu().numpy()) <= 2e-2
    # and against what the stock path (bf16 group_norm -> bf16 silu) returns
    stock = norm(x)
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
