"""Micro-benchmark + accuracy check of the attention kernels of the path at the shapes the edit loop launches (not a bench.py number;
used to iterate on a kernel and as the short command for `ncu --set full`).

    python scripts/bench_kernels.py [--iters 20] [--only fwd|bwd] [--quick]
"""
import sys

import torch

sys.path.insert(0, ".")
from geodiffuser_b200 import _lib  # noqa: E402
from geodiffuser_b200._lib import call, ptr, stream  # noqa: E402

iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 20
only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
quick = "--quick" in sys.argv


def timed(fn, n):
    """mean GPU time of one call, ms: `n` calls recorded into one CUDA graph and replayed, so that the host's per-call cost (ctypes + 3 G tensor-map
    encodes: ~30-45 us, longer than the 32^2-token kernels themselves) stays out of the measurement"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def mk(g, *shape, s=1.5):
    return (torch.randn(*shape, device="cuda", generator=g) * s).bfloat16()


def fwd_case(entry, G, H, N, d, check=True):
    g = torch.Generator(device="cuda").manual_seed(N + d + G)
    qs = [mk(g, H, N, d) for _ in range(G)]
    k, v = mk(g, H, N, d), mk(g, H, N, d)
    ks, vs = [k] * G, [v] * G
    O = torch.empty(G, H, N, d, device="cuda", dtype=torch.float32)
    L = torch.empty(G, H, N, device="cuda", dtype=torch.float32)
    scale = d ** -0.5

    def run():
        call(entry, _lib.ptr_array(qs), _lib.ptr_array(ks), _lib.ptr_array(vs), _lib.ptr_array([O[i] for i in range(G)]),
             _lib.ptr_array([L[i] for i in range(G)]), None, G, H, N, N, d, float(scale), None, 0, stream())

    ms = timed(run, iters)
    fl = 4.0 * G * H * N * N * d
    err = errl = float("nan")
    if check:
        s = torch.einsum("hnd,hkd->hnk", qs[G - 1].float(), k.float()) * scale
        ref = torch.softmax(s, -1) @ v.float()
        err, errl = rel(O[G - 1], ref), rel(L[G - 1], torch.logsumexp(s, -1))
    print(f"{entry:22s} G={G} H={H} N={N:5d} d={d:3d}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s  relerr O {err:.2e} LSE {errl:.2e}", flush=True)


def bwd_case(H, N, d, M=0, sm100=False, clustered=False):
    g = torch.Generator(device="cuda").manual_seed(N + d)
    q, k, v = mk(g, H, N, d), mk(g, H, N, d), mk(g, H, N, d)
    do = mk(g, H, N, d, s=1.0)
    scale = d ** -0.5
    O = torch.empty(1, H, N, d, device="cuda", dtype=torch.float32)
    L = torch.empty(1, H, N, device="cuda", dtype=torch.float32)
    entry = "gd_attn_fwd_sm100" if (N % 128 == 0 and d in (40, 80)) else "gd_attn_fwd_generic"
    call(entry, _lib.ptr_array([q]), _lib.ptr_array([k]), _lib.ptr_array([v]), _lib.ptr_array([O[0]]), _lib.ptr_array([L[0]]), None, 1, H, N, N, d,
         float(scale), None, 0, stream())
    delta = (do.float() * O[0]).sum(-1).contiguous()
    dq = torch.empty(H, N, d, device="cuda", dtype=torch.float32)
    ld = (N + 7) // 8 * 8
    extra = rowmap = dl = None
    if M:
        rows = (torch.arange(M, device="cuda") + N // 3) if clustered else torch.randperm(N, device="cuda")[:M].sort().values
        rows = rows.int()
        rowmap = torch.full((N,), -1, device="cuda", dtype=torch.int32)
        rowmap[rows.long()] = torch.arange(M, device="cuda", dtype=torch.int32)
        extra = torch.randn(H, M, ld, device="cuda") * 0.01
        dl = torch.ones(1, device="cuda")
        Mp = (M + 3) // 4 * 4
        extra_t = torch.zeros(H, N, Mp, device="cuda")          # key-major copy: the layout the product path hands to the tcgen05 kernel
        extra_t[:, :, :M] = extra[:, :, :N].transpose(1, 2)

    def run():
        if sm100:
            call("gd_attn_bwd_sm100", ptr(q), ptr(k), ptr(v), ptr(do), ptr(L[0]), ptr(delta), ptr(extra_t if M else None), ptr(dl), ptr(rowmap), Mp if M else ld, M,
                 ptr(dq), H, N, d, float(scale), None, 0, 1 if M else 0, stream())
        else:
            call("gd_attn_bwd", 0, ptr(q), ptr(k), ptr(v), ptr(do), ptr(L[0]), ptr(delta), ptr(extra), ptr(dl), ptr(rowmap), ld, M, ptr(dq), H, N, N,
                 d, float(scale), None, 0, stream())

    ms = timed(run, iters)
    fl = 6.0 * H * N * N * d
    # fp32 reference
    s = torch.einsum("hnd,hkd->hnk", q.float(), k.float()) * scale
    p = torch.softmax(s, -1)
    dp = torch.einsum("hnd,hkd->hnk", do.float(), v.float())
    if M:
        dp[:, rows.long(), :] += extra[:, :, :N]
        # delta gets the extra term too in the real path; keep the same delta on both sides here
    ds = p * (dp - delta[..., None])
    ref = torch.einsum("hnk,hkd->hnd", ds, k.float()) * scale
    print(f"gd_attn_bwd{'_sm100' if sm100 else '      '}(dQ)  H={H} N={N:5d} d={d:3d} M={M:4d}{' (clustered)' if clustered else ''}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s  relerr dQ {rel(dq, ref):.2e}",
          flush=True)


def corr_case(H, N, d, M):
    """removal-loss correlation + masked arg-max: round-1 path (materialised base map, mma.sync GEMMs) against the tcgen05 kernel"""
    g = torch.Generator(device="cuda").manual_seed(N + d + M)
    q_b, k_b, q_e = mk(g, H, N, d), mk(g, H, N, d), mk(g, H, N, d)
    scale = d ** -0.5
    lse_b = torch.logsumexp(torch.einsum("hnd,hkd->hnk", q_b.float(), k_b.float()) * scale, -1).contiguous()
    lse_e = torch.logsumexp(torch.einsum("hnd,hkd->hnk", q_e.float(), k_b.float()) * scale, -1).contiguous()
    rows = (torch.arange(M, device="cuda") + N // 3).int()
    m_in = torch.zeros(N, device="cuda"); m_in[rows.long()] = 1.0
    m_bg = ((torch.rand(N, device="cuda", generator=g) > 0.3).float() * (1 - m_in)).contiguous()
    ld = (N + 7) // 8 * 8
    a_e = torch.empty(H, M, ld, device="cuda", dtype=torch.bfloat16)
    call("gd_attn_probs", ptr(q_e), ptr(k_b), ptr(lse_e), ptr(rows), M, H, N, N, d, float(scale), ptr(a_e), ld, None, stream())
    a_b = torch.empty(H, N, ld, device="cuda", dtype=torch.bfloat16)
    p_old = torch.empty(H, (N + 63) // 64, M, 4, device="cuda")
    p_new = torch.empty(H, N // 32, M, 4, device="cuda")

    def old():
        call("gd_attn_probs", ptr(q_b), ptr(k_b), ptr(lse_b), None, N, H, N, N, d, float(scale), ptr(a_b), ld, None, stream())
        call("gd_corr_max_partial", ptr(a_e), ptr(a_b), H, M, N, N, ld, ptr(m_in), ptr(m_bg), ptr(p_old), stream())

    def new():
        call("gd_removal_corr_sm100", ptr(q_b), ptr(k_b), ptr(lse_b), ptr(a_e), H, M, N, d, float(scale), ld, None, ptr(m_in), ptr(m_bg), ptr(p_new), stream())

    ms_old, ms_new = timed(old, iters), timed(new, iters)
    fl = 2.0 * H * M * N * N
    red = lambda p: (p[..., 0].amax(1), p[..., 2].amax(1))          # max over the column tiles: (H, M) for the inpaint and the background mask
    (oi, ob), (ni, nb) = red(p_old), red(p_new)
    # fp32 reference of the two maxima on the same bf16 maps
    corr = torch.einsum("hmk,hnk->hmn", a_e[:, :, :N].float(), a_b[:, :, :N].float())
    ri, rb = (corr * m_in).amax(-1), (corr * m_bg).amax(-1)
    print(f"removal correlation    H={H} N={N:5d} d={d:3d} M={M:4d}: materialised map + mma.sync {ms_old * 1e3:8.1f} us ({fl / ms_old / 1e9:6.1f} TFLOP/s) | "
          f"tcgen05 {ms_new * 1e3:8.1f} us ({fl / ms_new / 1e9:6.1f} TFLOP/s)  relerr vs fp32: max_in {rel(ni, ri):.2e} max_bg {rel(nb, rb):.2e} "
          f"(mma.sync path: {rel(oi, ri):.2e} {rel(ob, rb):.2e})", flush=True)


def rows_case(H, N, d, M):
    """removal term of dQ as its own contraction: W = A_e[rows] o (g_bg P2[m] + g_in P2[M+m]), dq[rows] += gscale * scale * W @ K"""
    g = torch.Generator(device="cuda").manual_seed(N + d + M)
    ld = (N + 7) // 8 * 8
    a_e = torch.rand(H, M, ld, device="cuda", generator=g).bfloat16() * 0.01
    p2 = torch.rand(H, 2 * M, ld, device="cuda", generator=g).bfloat16() * 0.01
    g2 = torch.randn(H * M, 2, device="cuda", generator=g)
    k = mk(g, H, N, d)
    rows = (torch.arange(M, device="cuda") + N // 3).int()
    gs = torch.full((1,), 0.7, device="cuda")
    w = torch.empty(H, M, ld, device="cuda", dtype=torch.bfloat16)
    dq = torch.zeros(H, N, d, device="cuda")
    scale = d ** -0.5
    f_w = lambda: call("gd_removal_weighted_rows", ptr(a_e), ptr(p2), ptr(g2), H, M, N, ld, ptr(w), stream())
    f_d = lambda: call("gd_removal_dq_rows", ptr(w), ptr(k), ptr(rows), ptr(gs), ptr(dq), H, M, N, N, d, float(scale), ld, None, 0, stream())
    ms_w, ms_d = timed(f_w, iters), timed(f_d, iters)
    dq.zero_()
    f_w(); f_d()
    gg = g2.reshape(H, M, 2)
    wr = a_e.float() * (gg[..., :1] * p2[:, :M].float() + gg[..., 1:] * p2[:, M:].float())
    ref = torch.einsum("hmk,hkd->hmd", wr[:, :, :N], k.float()) * scale * 0.7
    got = dq[:, rows.long(), :]
    print(f"removal dQ rows        H={H} N={N:5d} d={d:3d} M={M:4d}: weighted rows {ms_w * 1e3:6.1f} us + W@K {ms_d * 1e3:6.1f} us  relerr {rel(got, ref):.2e}  "
          f"(rows outside the set untouched: {bool((dq.abs().sum((0, 2)) > 0).sum() == M)})", flush=True)


cfg = lambda key, value: call("gd_attn_sm100_config", key, value)

if only in ("corr", "sweep"):
    corr_case(8, 4096, 40, 410)
    corr_case(8, 4096, 40, 76)
    corr_case(8, 4096, 40, 640)
    corr_case(8, 1024, 80, 100)

if only == "fwd40":
    for np_ in (2, 3, 4):
        cfg(0, np_)
        print(f"--- fwd: {np_}/8 pairs on the polynomial", flush=True)
        fwd_case("gd_attn_fwd_sm100", 3, 8, 4096, 40)
        fwd_case("gd_attn_fwd_sm100", 4, 8, 4096, 40)
        fwd_case("gd_attn_fwd_sm100", 2, 8, 4096, 40)
        fwd_case("gd_attn_fwd_sm100", 3, 8, 1024, 80)
        fwd_case("gd_attn_fwd_sm100", 1, 2, 9216, 40)
    cfg(0, 2)

if only == "rows":
    rows_case(8, 4096, 40, 76)
    rows_case(8, 4096, 40, 410)
    rows_case(8, 1024, 80, 18)
    rows_case(8, 1024, 80, 100)
    rows_case(2, 9216, 40, 900)

if only == "ab":
    for bnk in (128, 64):
        cfg(2, bnk)
        print(f"--- fwd: {bnk} keys per step", flush=True)
        fwd_case("gd_attn_fwd_sm100", 3, 8, 4096, 40)
        fwd_case("gd_attn_fwd_sm100", 4, 8, 4096, 40)
        fwd_case("gd_attn_fwd_sm100", 3, 8, 1024, 80)
        fwd_case("gd_attn_fwd_sm100", 4, 8, 1024, 80)
    bwd_case(8, 4096, 40, sm100=True)
    bwd_case(8, 4096, 40, M=410, sm100=True)
    bwd_case(8, 4096, 40, M=76, sm100=True, clustered=True)
    bwd_case(8, 1024, 80, sm100=True)
    bwd_case(8, 1024, 80, M=18, sm100=True, clustered=True)
    corr_case(8, 4096, 40, 410)
    corr_case(8, 4096, 40, 76)
    corr_case(8, 1024, 80, 18)

if only == "fwdsweep":
    for bnk in (128, 64):
        cfg(2, bnk)
        for np_ in ((2, 3) if quick else (1, 2, 3)):
            cfg(0, np_)
            print(f"--- fwd: {bnk} keys per step, {np_}/8 pairs on the polynomial", flush=True)
            fwd_case("gd_attn_fwd_sm100", 3, 8, 4096, 40)
            fwd_case("gd_attn_fwd_sm100", 4, 8, 4096, 40)
            fwd_case("gd_attn_fwd_sm100", 2, 8, 4096, 40)
            fwd_case("gd_attn_fwd_sm100", 3, 8, 1024, 80)
            fwd_case("gd_attn_fwd_sm100", 4, 8, 1024, 80)
            fwd_case("gd_attn_fwd_sm100", 1, 2, 9216, 40)
    cfg(0, 2)
    cfg(2, 64)

if only == "sweep":
    for np_ in (0, 1, 2, 3, 4):
        cfg(0, np_)
        print(f"--- fwd: packed arithmetic, {np_}/8 pairs on the polynomial", flush=True)
        fwd_case("gd_attn_fwd_sm100", 3, 8, 4096, 40)
        fwd_case("gd_attn_fwd_sm100", 4, 8, 4096, 40)
        fwd_case("gd_attn_fwd_sm100", 3, 8, 1024, 80)
        fwd_case("gd_attn_fwd_sm100", 1, 2, 9216, 40)
    cfg(0, 2)
    for variant, nps in ((1, (0, 1)),):
        for np_ in nps:
            cfg(3, np_)
            print(f"--- bwd: variant {variant} ({'128-key steps, 1 CTA/SM' if variant == 0 else f'64-key steps, 2 CTA/SM, {np_}/8 pairs on the polynomial'})", flush=True)
            bwd_case(8, 4096, 40, sm100=True)
            bwd_case(8, 4096, 40, M=410, sm100=True)
            bwd_case(8, 1024, 80, sm100=True)
            bwd_case(8, 1024, 80, M=100, sm100=True)
            bwd_case(2, 9216, 40, sm100=True)
    cfg(3, 0)
if only in (None, "fwd"):
    fwd_case("gd_attn_fwd_sm100", 3, 8, 4096, 40)
    fwd_case("gd_attn_fwd_sm100", 4, 8, 4096, 40)
    fwd_case("gd_attn_fwd_sm100", 3, 8, 1024, 80)
    fwd_case("gd_attn_fwd_sm100", 4, 8, 1024, 80)
    if not quick:
        fwd_case("gd_attn_fwd_sm100", 3, 8, 9216, 40, check=False)
        fwd_case("gd_attn_fwd_generic", 3, 8, 4096, 40)
        fwd_case("gd_attn_fwd_generic", 3, 8, 256, 160)
if only in (None, "bwd"):
    bwd_case(8, 4096, 40, sm100=True)
    bwd_case(8, 4096, 40, M=410, sm100=True)
    bwd_case(8, 4096, 40, M=410, sm100=True, clustered=True)
    bwd_case(8, 4096, 40, M=76, sm100=True, clustered=True)
    bwd_case(8, 1024, 80, sm100=True)
    bwd_case(8, 1024, 80, M=100, sm100=True)
    bwd_case(2, 9216, 40, sm100=True)
    if not quick:
        bwd_case(8, 4096, 40)
        bwd_case(8, 4096, 40, M=410)
        bwd_case(8, 1024, 80)
        bwd_case(8, 256, 160)
