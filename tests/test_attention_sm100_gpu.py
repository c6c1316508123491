"""tcgen05 / TMEM / TMA forward (csrc/attention_sm100.cu) against the fp32 evaluation of the same bf16 inputs and against the
mma.sync kernel, at the self-attention shapes it serves: 64^2 level (N=4096, d=40), 32^2 level (N=1024, d=80), 96^2 (768^2 images)."""
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu


def _run(entry, qs, ks, vs, scale):
    from geodiffuser_b200 import _lib
    from geodiffuser_b200._lib import call, stream

    G = len(qs)
    H, N, d = qs[0].shape
    O = torch.empty(G, H, N, d, device="cuda", dtype=torch.float32)
    L = torch.empty(G, H, N, device="cuda", dtype=torch.float32)
    call(entry, _lib.ptr_array(qs), _lib.ptr_array(ks), _lib.ptr_array(vs), _lib.ptr_array([O[g] for g in range(G)]),
         _lib.ptr_array([L[g] for g in range(G)]), G, H, N, N, d, float(scale), stream())
    torch.cuda.synchronize()
    return O, L


@pytest.mark.parametrize("N,d,H,G", [(4096, 40, 8, 3), (1024, 80, 8, 3), (128, 40, 1, 1), (256, 80, 2, 2), (9216, 40, 2, 1)])
@pytest.mark.timeout(120)
def test_sm100_forward(N, d, H, G):
    g = torch.Generator(device="cuda").manual_seed(N + d)
    mk = lambda: (torch.randn(H, N, d, device="cuda", generator=g) * 1.5).bfloat16()
    qs = [mk() for _ in range(G)]
    k, v = mk(), mk()
    k2, v2 = mk(), mk()
    ks, vs = [k] * G, [v] * G
    if G > 1:   # streams may point at different K/V (plain CFG entries) or share them (warp / edit streams)
        ks[0], vs[0] = k2, v2
    scale = d ** -0.5
    O, L = _run("gd_attn_fwd_sm100", qs, ks, vs, scale)
    O2, L2 = _run("gd_attn_fwd_generic", qs, ks, vs, scale)
    for i in range(G):
        if N <= 4096:
            s = torch.einsum("hnd,hkd->hnk", qs[i].float(), ks[i].float()) * scale
            ref = torch.softmax(s, -1) @ vs[i].float()
            assert relerr(O[i].cpu().numpy(), ref.cpu().numpy()) <= 1e-2, i
            assert relerr(L[i].cpu().numpy(), torch.logsumexp(s, -1).cpu().numpy()) <= 1e-3, i
        assert relerr(O[i].cpu().numpy(), O2[i].cpu().numpy()) <= 1e-2, i
        assert relerr(L[i].cpu().numpy(), L2[i].cpu().numpy()) <= 1e-3, i


@pytest.mark.timeout(120)
def test_sm100_large_logits_lazy_rescale():
    """rows whose running max keeps growing (sorted keys) exercise the in-TMEM O correction"""
    H, N, d = 2, 1024, 40
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(H, N, d, device="cuda", generator=g).bfloat16()
    k = (torch.randn(H, N, d, device="cuda", generator=g) * torch.linspace(0.2, 6.0, N, device="cuda")[None, :, None]).bfloat16()
    v = torch.randn(H, N, d, device="cuda", generator=g).bfloat16()
    scale = 1.0
    O, L = _run("gd_attn_fwd_sm100", [q], [k], [v], scale)
    s = torch.einsum("hnd,hkd->hnk", q.float(), k.float()) * scale
    ref = torch.softmax(s, -1) @ v.float()
    assert torch.isfinite(O).all()
    assert relerr(O[0].cpu().numpy(), ref.cpu().numpy()) <= 2e-2
    assert relerr(L[0].cpu().numpy(), torch.logsumexp(s, -1).cpu().numpy()) <= 1e-3


def test_sm100_rejects_unsupported_shapes():
    from geodiffuser_b200 import _lib

    q = torch.zeros(1, 100, 40, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.GeoDiffuserB200Error):
        _run("gd_attn_fwd_sm100", [q], [q], [q], 1.0)


@pytest.mark.parametrize("N,d,H,M", [(4096, 40, 8, 0), (4096, 40, 8, 410), (1024, 80, 8, 100), (128, 40, 1, 5), (256, 80, 2, 0), (2304, 40, 2, 64)])
@pytest.mark.timeout(120)
def test_sm100_backward_dq(N, d, H, M):
    """tcgen05 dQ kernel vs the fp32 evaluation of dQ = scale * (P o (dO V^T + extra - delta)) K on the same bf16 inputs, and vs the mma.sync
    kernel it replaces at these shapes; `extra` = dL/dP rows of the removal loss (random here), scaled by a device scalar."""
    from geodiffuser_b200 import _lib
    from geodiffuser_b200._lib import call, ptr, stream

    g = torch.Generator(device="cuda").manual_seed(N + d + M)
    mk = lambda s=1.5: (torch.randn(H, N, d, device="cuda", generator=g) * s).bfloat16()
    q, k, v, do = mk(), mk(), mk(), mk(1.0)
    scale = d ** -0.5
    O, L = _run("gd_attn_fwd_sm100", [q], [k], [v], scale)
    s = torch.einsum("hnd,hkd->hnk", q.float(), k.float()) * scale
    p = torch.softmax(s, -1)
    dp = torch.einsum("hnd,hkd->hnk", do.float(), v.float())
    ld = (N + 7) // 8 * 8
    extra = rowmap = dl = None
    if M:
        rows = torch.randperm(N, device="cuda", generator=g)[:M].sort().values.int()
        rowmap = torch.full((N,), -1, device="cuda", dtype=torch.int32)
        rowmap[rows.long()] = torch.arange(M, device="cuda", dtype=torch.int32)
        extra = torch.randn(H, M, ld, device="cuda", generator=g) * 0.05
        dl = torch.full((1,), 0.7, device="cuda")
        dp[:, rows.long(), :] += 0.7 * extra[:, :, :N]
    delta = (p * dp).sum(-1).contiguous()          # the row term of the softmax Jacobian (what gd_attn_bwd_prep produces on the path)
    ref = torch.einsum("hnk,hkd->hnd", p * (dp - delta[..., None]), k.float()) * scale
    out = []
    for entry in ("gd_attn_bwd_sm100", "gd_attn_bwd"):
        dq = torch.full((H, N, d), float("nan"), device="cuda", dtype=torch.float32)
        if entry == "gd_attn_bwd":
            call(entry, 0, ptr(q), ptr(k), ptr(v), ptr(do), ptr(L[0]), ptr(delta), ptr(extra), ptr(dl), ptr(rowmap), ld, M, ptr(dq), H, N, N, d,
                 float(scale), stream())
        else:
            call(entry, ptr(q), ptr(k), ptr(v), ptr(do), ptr(L[0]), ptr(delta), ptr(extra), ptr(dl), ptr(rowmap), ld, M, ptr(dq), H, N, d, float(scale),
                 stream())
        torch.cuda.synchronize()
        out.append(dq)
    assert torch.isfinite(out[0]).all()
    assert relerr(out[0].cpu().numpy(), ref.cpu().numpy()) <= 1e-2       # bf16 operands / bf16 dS vs fp32 math (tolerance of the path: 2e-2)
    assert relerr(out[0].cpu().numpy(), out[1].cpu().numpy()) <= 1e-2
