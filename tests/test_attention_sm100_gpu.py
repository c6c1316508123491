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
