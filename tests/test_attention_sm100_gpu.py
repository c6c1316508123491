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
         _lib.ptr_array([L[g] for g in range(G)]), None, G, H, N, N, d, float(scale), None, 0, stream())
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
    # the tcgen05 kernel with the removal rows row-major (H, M, ld) and key-major (H, N, Mp) [what the product path passes], then the mma.sync kernel
    extra_t, Mp = None, (M + 3) // 4 * 4
    if M:
        extra_t = torch.zeros(H, N, Mp, device="cuda")
        extra_t[:, :, :M] = extra[:, :, :N].transpose(1, 2)
    for entry, key_major in (("gd_attn_bwd_sm100", 0), ("gd_attn_bwd_sm100", 1), ("gd_attn_bwd", 0)):
        dq = torch.full((H, N, d), float("nan"), device="cuda", dtype=torch.float32)
        if entry == "gd_attn_bwd":
            call(entry, 0, ptr(q), ptr(k), ptr(v), ptr(do), ptr(L[0]), ptr(delta), ptr(extra), ptr(dl), ptr(rowmap), ld, M, ptr(dq), H, N, N, d,
                 float(scale), None, 0, stream())
        else:
            call(entry, ptr(q), ptr(k), ptr(v), ptr(do), ptr(L[0]), ptr(delta), ptr(extra_t if key_major else extra), ptr(dl), ptr(rowmap),
                 Mp if key_major else ld, M, ptr(dq), H, N, d, float(scale), None, 0, key_major if M else 0, stream())
        torch.cuda.synchronize()
        out.append(dq)
    assert torch.isfinite(out[0]).all()
    assert relerr(out[0].cpu().numpy(), ref.cpu().numpy()) <= 1e-2       # bf16 operands / bf16 dS vs fp32 math (tolerance of the path: 2e-2)
    assert torch.equal(out[0], out[1])                                   # the layout of the removal rows does not change a bit
    assert relerr(out[0].cpu().numpy(), out[2].cpu().numpy()) <= 1e-2


@pytest.mark.parametrize("N,Nk,d,H", [(4096, 4096, 40, 8), (1024, 1024, 80, 8), (256, 256, 160, 8), (1024, 77, 80, 8), (200, 77, 40, 4)])
@pytest.mark.timeout(120)
def test_projection_layout_is_read_and_written_in_place(N, Nk, d, H):
    """q / k / v as the projections produce them, (B, N, H*d), addressed through slab strides (no head_to_batch_dim copy), and the output /
    dQ / dK written straight into (B, N, H*d) tensors (no batch_to_head_dim copy): bit-identical to the same kernels on the permuted,
    contiguous (B*H, N, d) tensors, forward (bf16 strided output + fp32 output + lse) and backward."""
    from geodiffuser_b200 import _lib, functional as Fn
    from geodiffuser_b200._lib import call, ptr, stream

    B = 2
    g = torch.Generator(device="cuda").manual_seed(N + d)
    qp = (torch.randn(B, N, H * d, device="cuda", generator=g) * 1.5).bfloat16()
    kp = (torch.randn(B, Nk, H * d, device="cuda", generator=g) * 1.5).bfloat16()
    vp = (torch.randn(B, Nk, H * d, device="cuda", generator=g) * 1.5).bfloat16()
    perm = lambda t: t.reshape(t.shape[0], t.shape[1], H, d).permute(0, 2, 1, 3).reshape(t.shape[0] * H, t.shape[1], d).contiguous()
    unperm = lambda t: t.reshape(B, H, t.shape[1], d).permute(0, 2, 1, 3).reshape(B, t.shape[1], H * d)
    scale = d ** -0.5
    out_p = Fn.plain_attention(Fn.ProjView(qp, H), Fn.ProjView(kp, H), Fn.ProjView(vp, H), scale, H)
    out_h = Fn.plain_attention(perm(qp), perm(kp), perm(vp), scale, H)
    assert out_p.shape == (B, N, H * d) and out_p.dtype == torch.bfloat16
    assert torch.equal(out_p, unperm(out_h))
    s = torch.einsum("bhnd,bhkd->bhnk", perm(qp).float().reshape(B, H, N, d), perm(kp).float().reshape(B, H, Nk, d)) * scale
    ref = torch.softmax(s, -1) @ perm(vp).float().reshape(B, H, Nk, d)
    assert relerr(out_h.float().cpu().numpy(), ref.reshape(B * H, N, d).cpu().numpy()) <= 1.5e-2

    # backward entry points: dQ (and dK for Nk != N) of batch entry 1, strided bf16 outputs vs contiguous fp32 outputs
    lay_p, lay_h = Fn._Layout(qp, kp, H, True), Fn._Layout(perm(qp), perm(kp), H, False)
    qh, kh, vh = perm(qp), perm(kp), perm(vp)
    O, L = Fn.attention_forward([lay_h.sl(qh, 1)], [lay_h.sl(kh, 1)], [lay_h.sl(vh, 1)], scale)
    do = torch.randn(H, N, d, device="cuda", generator=g).bfloat16()
    delta = (do.float() * O[0]).sum(-1).contiguous()
    sm100 = Nk == N and N % 128 == 0 and d in (40, 80)
    res = {}
    for name, lay, (q_, k_, v_) in (("proj", lay_p, (qp, kp, vp)), ("heads", lay_h, (qh, kh, vh))):
        dq = torch.zeros_like(q_)
        args = (_lib.base_ptr(lay.sl(q_, 1)), _lib.base_ptr(lay.sl(k_, 1)), _lib.base_ptr(lay.sl(v_, 1)), ptr(do), ptr(L[0]), ptr(delta), None,
                None, None, (Nk + 7) // 8 * 8, 0, _lib.base_ptr(lay.sl(dq, 1)))
        if sm100:
            call("gd_attn_bwd_sm100", *args, H, N, d, float(scale), lay.strides(), 1, 0, stream())
        else:
            call("gd_attn_bwd", 0, *args, H, N, Nk, d, float(scale), lay.strides(), 1, stream())
        dk = torch.zeros_like(k_)
        ws = torch.empty(4, H, Nk, d, device="cuda")
        call("gd_attn_bwd_dk_split", *args[:11], _lib.base_ptr(lay.sl(dk, 1)), ptr(ws), 4, H, N, Nk, d, float(scale), lay.strides(out=lay.kv), 1,
             stream())
        torch.cuda.synchronize()
        res[name] = (dq, dk)
    assert torch.equal(res["proj"][0], unperm(res["heads"][0])) and float(res["proj"][0][1].abs().max()) > 0
    assert float(res["proj"][0][0].abs().max()) == 0.0
    assert torch.equal(res["proj"][1], unperm(res["heads"][1])) and float(res["proj"][1][1].abs().max()) > 0


@pytest.mark.parametrize("d", [40, 80])
@pytest.mark.parametrize("pattern", ["staircase", "ramp_then_flat", "spike"])
def test_forward_reference_moves_inside_a_tile(d, pattern):
    """The online softmax rescales O lazily (only when a row's running max grows by more than 2^8).  Adversarial score profiles exercise that
    path on every tile: a staircase rising by ~17 nats every 16 keys, a ramp that stops, one spike in the middle of a tile.  Against fp32 torch."""
    from geodiffuser_b200 import functional as Fn

    H, N = 2, 1024
    g = torch.Generator(device="cuda").manual_seed(d + len(pattern))
    u = torch.nn.functional.normalize(torch.randn(d, device="cuda", generator=g), dim=0)
    n = torch.arange(N, device="cuda", dtype=torch.float32)
    if pattern == "staircase":
        prof = torch.floor(n / 16) * 2.5
    elif pattern == "ramp_then_flat":
        prof = torch.clamp(n * 0.4, max=120.0)
    else:
        prof = torch.zeros(N, device="cuda"); prof[N // 2 + 37] = 90.0
    k = (prof[:, None] * u[None, :] + 0.3 * torch.randn(N, d, device="cuda", generator=g))[None].repeat(H, 1, 1).bfloat16()
    q = ((d ** 0.5) * 0.5 * u[None, :] + 0.3 * torch.randn(N, d, device="cuda", generator=g))[None].repeat(H, 1, 1).bfloat16()
    v = torch.randn(H, N, d, device="cuda", generator=g).bfloat16()
    scale = d ** -0.5
    O, LSE = Fn.attention_forward([q], [k], [v], scale)
    s = torch.einsum("hnd,hkd->hnk", q.float(), k.float()) * scale
    ref = torch.softmax(s, -1) @ v.float()
    assert torch.isfinite(O[0]).all() and torch.isfinite(LSE[0]).all()
    assert relerr(O[0].cpu().numpy(), ref.cpu().numpy()) <= 1e-2
    assert relerr(LSE[0].cpu().numpy(), torch.logsumexp(s, -1).cpu().numpy()) <= 1e-3


@pytest.mark.parametrize("bnk", [64, 128])
@pytest.mark.parametrize("N,d,H,G", [(1024, 40, 8, 3), (1024, 80, 4, 2)])
@pytest.mark.timeout(120)
def test_sm100_forward_both_step_sizes(bnk, N, d, H, G):
    """gd_attn_sm100_config key 2: 64- and 128-key steps are the same computation (the default picks 128 at head_dim 40, 64 at head_dim 80; the
    other two instances exist for A/B measurements and must stay correct): head_dim 40 at 64 keys runs on two TMEM allocations, three CTAs per SM."""
    from geodiffuser_b200._lib import call

    g = torch.Generator(device="cuda").manual_seed(N + d + bnk)
    mk = lambda: (torch.randn(H, N, d, device="cuda", generator=g) * 1.5).bfloat16()
    qs = [mk() for _ in range(G)]
    k, v = mk(), mk()
    scale = d ** -0.5
    try:
        call("gd_attn_sm100_config", 2, bnk)
        O, L = _run("gd_attn_fwd_sm100", qs, [k] * G, [v] * G, scale)
    finally:
        call("gd_attn_sm100_config", 2, 0)
    for i in range(G):
        s = torch.einsum("hnd,hkd->hnk", qs[i].float(), k.float()) * scale
        ref = torch.softmax(s, -1) @ v.float()
        assert relerr(O[i].cpu().numpy(), ref.cpu().numpy()) <= 1e-2, i
        assert relerr(L[i].cpu().numpy(), torch.logsumexp(s, -1).cpu().numpy()) <= 1e-3, i


@pytest.mark.parametrize("N,d,H,M,proj,dq_dtype", [(4096, 40, 8, 76, False, torch.float32), (1024, 80, 8, 18, True, torch.bfloat16),
                                                   (1024, 40, 2, 37, True, torch.float32), (2304, 80, 2, 300, False, torch.bfloat16)])
@pytest.mark.timeout(120)
def test_removal_dq_rows_is_the_weighted_contraction(N, d, H, M, proj, dq_dtype):
    """gd_removal_weighted_rows + gd_removal_dq_rows: dq[h, rows[m], :] += gscale * scale * sum_k (A_e[m,k] (g_bg P2[m,k] + g_in P2[M+m,k])) K[h,k,:]
    -- the removal term of dQ (attention_processors.py:262-280 through autograd in the reference) -- against fp32 on the same bf16 operands, with
    contiguous slabs and with the projection layout (B, N, H*d) the processors hand over, accumulating into fp32 and bf16 gradients."""
    from geodiffuser_b200._lib import call, ptr, stream, host_longs

    g = torch.Generator(device="cuda").manual_seed(N + d + M)
    ld = (N + 7) // 8 * 8
    a_e = (torch.rand(H, M, ld, device="cuda", generator=g) * 0.02).bfloat16()
    p2 = (torch.rand(H, 2 * M, ld, device="cuda", generator=g) * 0.02).bfloat16()
    g2 = torch.randn(H * M, 2, device="cuda", generator=g)
    rows = torch.randperm(N, device="cuda", generator=g)[:M].sort().values.int()
    gs = torch.full((1,), 0.7, device="cuda")
    scale = d ** -0.5
    if proj:
        kfull = (torch.randn(N, H * d, device="cuda", generator=g) * 1.5).bfloat16()
        k = kfull.reshape(N, H, d).permute(1, 0, 2)                                   # (H, N, d) view of the projection layout
        dqfull = (torch.randn(N, H * d, device="cuda", generator=g) * 0.01).to(dq_dtype)
        dq = dqfull.reshape(N, H, d).permute(1, 0, 2)
        strides = host_longs([H * d, d, H * d, d])
        kbase, dqbase = kfull, dqfull
    else:
        k = (torch.randn(H, N, d, device="cuda", generator=g) * 1.5).bfloat16()
        dq = (torch.randn(H, N, d, device="cuda", generator=g) * 0.01).to(dq_dtype)
        strides, kbase, dqbase = None, k, dq
    before = dq.float().clone()
    w = torch.empty(H, M, ld, device="cuda", dtype=torch.bfloat16)
    call("gd_removal_weighted_rows", ptr(a_e), ptr(p2), ptr(g2), H, M, N, ld, ptr(w), stream())
    call("gd_removal_dq_rows", ptr(w), ptr(kbase), ptr(rows), ptr(gs), ptr(dqbase), H, M, N, N, d, float(scale), ld, strides, int(dq_dtype == torch.bfloat16), stream())
    torch.cuda.synchronize()
    gg = g2.reshape(H, M, 2)
    wr = a_e.float() * (gg[..., :1] * p2[:, :M].float() + gg[..., 1:] * p2[:, M:].float())
    assert relerr(w[:, :, :N].float().cpu().numpy(), wr[:, :, :N].cpu().numpy()) <= 1e-2          # bf16 rounding of W
    add = torch.einsum("hmk,hkd->hmd", wr[:, :, :N], k.float()) * scale * 0.7
    got = dq.float() - before
    tol = 1e-2 if dq_dtype == torch.float32 else 5e-2                                 # (bf16 accumulation rounds the sum to 8 bits)
    assert relerr(got[:, rows.long()].cpu().numpy(), add.cpu().numpy()) <= tol
    untouched = torch.ones(N, dtype=torch.bool, device="cuda")
    untouched[rows.long()] = False
    assert torch.equal(got[:, untouched], torch.zeros_like(got[:, untouched]))         # rows outside the inpaint set are not written
