"""Achieved bandwidth of the HBM-bound kernels of the path and of the caller-side fused ops (CUDA events, algorithmic bytes / time), at the
size the edit loop launches them and on a batched synthetic of >= 256 MB traffic (at batch-1 sizes these kernels are
launch-latency / L2 bound; the batched figure is the one to read against the measured HBM peak).  Not a bench.py number.
    python scripts/bench_hbm_kernels.py > gpurun_out/hbm_kernels.log
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from geodiffuser_b200 import body_ops, geometry as G, image_processing as IP, synth  # noqa: E402

peak = 6554.9
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def timed(fn, n=20, cold=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if cold:
            flush.zero_()           # > 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, nbytes, fn, cold=False):
    ms = timed(fn, cold=cold)
    gbs = nbytes / ms / 1e6
    print(f"{name:64s} {nbytes / 1e6:9.1f} MB  {ms * 1e3:8.1f} us  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of measured HBM peak ({peak:.0f} GB/s)", flush=True)


g = torch.Generator(device="cuda").manual_seed(0)
cl = torch.channels_last
for B, C, S in ((2, 320, 64), (2, 960, 64), (3, 1280, 32), (64, 960, 64)):
    x = torch.randn(B, C, S, S, device="cuda", generator=g).bfloat16().contiguous(memory_format=cl)
    norm = torch.nn.GroupNorm(32, C).cuda().bfloat16().requires_grad_(False)
    big = x.numel() * 4 >= 256e6
    report(f"group_norm+silu nhwc fwd (2 launches)   B={B} C={C} S={S}", x.numel() * 4, lambda: body_ops.group_norm_act(norm, x, silu=True), cold=big)
    report(f"  stock torch group_norm + silu         B={B} C={C} S={S}", x.numel() * 4, lambda: torch.nn.functional.silu(norm(x)), cold=big)
for rows, Fd in ((2 * 4096, 1280), (64 * 4096, 1280)):
    proj = torch.randn(rows, 2 * Fd, device="cuda", generator=g).bfloat16()
    big = rows * Fd * 6 >= 256e6
    report(f"geglu fwd                               rows={rows} F={Fd}", rows * Fd * 6, lambda: body_ops.geglu(proj), cold=big)
    a, b = proj.chunk(2, -1)
    report(f"  stock torch a * gelu(g)               rows={rows} F={Fd}", rows * Fd * 6, lambda: a * torch.nn.functional.gelu(b), cold=big)
image, depth, mask, T = synth.edit_inputs("rotate3d")
geo = G.correspondence_field(depth.copy(), mask.copy(), T)
for S, H, d in ((64, 8, 40), (32, 8, 80), (16, 8, 160)):
    cS = G.reshape_transform_coords(geo["coords"][None], in_mat_shape=(1, 1, S, S))
    idx, _, d2 = G.splat_index(cS)
    q = torch.randn(H, S * S, d, device="cuda", generator=g).bfloat16()
    m = torch.rand(S * S, device="cuda", generator=g)
    nb = H * S * S * d * 2 * 2 + S * S * 15 * 8
    report(f"query splat composite (rows kernel)     S={S} H={H} d={d}", nb, lambda: G.splat_composite(q, idx, d2, channels_last=True, blend_mask=m, out_dtype=torch.bfloat16))
src = torch.randint(0, 256, (2048, 2048, 3), device="cuda", dtype=torch.uint8, generator=g)
tm = torch.randint(0, 200, (2048, 2048, 3), device="cuda", dtype=torch.uint8, generator=g)
mk = (torch.rand(2048, 2048, device="cuda", generator=g) > 0.3).float()
report("masked histogram matching (3 launches)  2048x2048x3", 2048 * 2048 * (3 + 3 + 8 + 3 + 24), lambda: IP.masked_histogram_matching(src, tm, mk, mk), cold=False)
s5 = torch.randint(0, 256, (512, 512, 3), device="cuda", dtype=torch.uint8, generator=g)
t5 = torch.randint(0, 200, (512, 512, 3), device="cuda", dtype=torch.uint8, generator=g)
m5 = (torch.rand(512, 512, device="cuda", generator=g) > 0.3).float()
report("masked histogram matching (3 launches)  512x512x3", 512 * 512 * (3 + 3 + 8 + 3 + 24), lambda: IP.masked_histogram_matching(s5, t5, m5, m5))
