"""diagnostic: mma backward (dQ, dK) at small head dims / few heads vs fp32 torch"""
import sys
import torch
sys.path.insert(0, ".")
from geodiffuser_b200._lib import call, ptr, stream

def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))

for (H, N, Nk, d, M) in [(2, 1024, 1024, 16, 0), (2, 1024, 1024, 16, 100), (8, 1024, 1024, 16, 0), (2, 1024, 1024, 40, 100), (2, 1024, 1024, 24, 0), (2, 256, 256, 32, 0), (2, 1024, 77, 16, 50)]:
    g = torch.Generator(device="cuda").manual_seed(N + d + M)
    q, do = [(torch.randn(H, N, d, device="cuda", generator=g) * s).bfloat16() for s in (1.5, 1.0)]
    k, v = [(torch.randn(H, Nk, d, device="cuda", generator=g) * 1.5).bfloat16() for _ in range(2)]
    scale = d ** -0.5
    s = torch.einsum("hnd,hkd->hnk", q.float(), k.float()) * scale
    p = torch.softmax(s, -1)
    L = torch.logsumexp(s, -1).contiguous()
    dp = torch.einsum("hnd,hkd->hnk", do.float(), v.float())
    ld = (Nk + 7) // 8 * 8
    extra = rowmap = dl = None
    if M:
        rows = torch.randperm(N, device="cuda", generator=g)[:M].sort().values.int()
        rowmap = torch.full((N,), -1, device="cuda", dtype=torch.int32)
        rowmap[rows.long()] = torch.arange(M, device="cuda", dtype=torch.int32)
        extra = torch.randn(H, M, ld, device="cuda", generator=g) * 0.05
        dl = torch.full((1,), 0.7, device="cuda")
        dp[:, rows.long(), :] += 0.7 * extra[:, :, :Nk]
    delta = (p * dp).sum(-1).contiguous()
    ds = p * (dp - delta[..., None])
    ref_q = torch.einsum("hnk,hkd->hnd", ds, k.float()) * scale
    ref_k = torch.einsum("hnk,hnd->hkd", ds, q.float()) * scale
    dq = torch.full((H, N, d), float("nan"), device="cuda")
    dk = torch.full((H, Nk, d), float("nan"), device="cuda")
    call("gd_attn_bwd", 0, ptr(q), ptr(k), ptr(v), ptr(do), ptr(L), ptr(delta), ptr(extra), ptr(dl), ptr(rowmap), ld, M, ptr(dq), H, N, Nk, d, float(scale), None, 0, stream())
    call("gd_attn_bwd", 1, ptr(q), ptr(k), ptr(v), ptr(do), ptr(L), ptr(delta), ptr(extra), ptr(dl), ptr(rowmap), ld, M, ptr(dk), H, N, Nk, d, float(scale), None, 0, stream())
    torch.cuda.synchronize()
    print(f"H={H} N={N} Nk={Nk} d={d} M={M}: dQ err {rel(dq, ref_q):.3e} dK err {rel(dk, ref_k):.3e} finite {bool(torch.isfinite(dq).all())}", flush=True)
