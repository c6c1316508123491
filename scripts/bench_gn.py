"""GroupNorm(+SiLU) forward per shape of the SD-1.5 body: one-launch cluster kernel vs two-launch scheme vs a plain device copy of the same
tensor (the streaming floor), back-to-back launches through the C ABI with preallocated buffers (CUDA events).  Diagnostic."""
import sys
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import _lib
from geodiffuser_b200._lib import call, ptr, stream
from geodiffuser_b200.body_ops import ptr_cl

cl = torch.channels_last
g = torch.Generator(device="cuda").manual_seed(0)
L = _lib.lib()
for B, C, S in ((2, 320, 64), (3, 320, 64), (2, 640, 32), (2, 640, 64), (2, 960, 64), (2, 1280, 16), (2, 1280, 8), (2, 2560, 16), (2, 1920, 32)):
    x = torch.randn(B, C, S, S, device="cuda", generator=g).bfloat16().contiguous(memory_format=cl)
    y = torch.empty_like(x)
    w = torch.ones(C, device="cuda").bfloat16(); b = torch.zeros(C, device="cuda").bfloat16()
    stats = torch.empty(B, 32, 2, device="cuda")
    n = L.gd_group_norm_nhwc_workspace(B, S * S, C, 32)
    ws = torch.empty(n, device="cuda"); cnt = torch.zeros(64, device="cuda", dtype=torch.int32)
    res = []
    for scheme in (1, 0, -1):
        if scheme >= 0:
            call("gd_group_norm_config", scheme)
            fn = lambda: call("gd_group_norm_nhwc_fwd", ptr_cl(x), None, ptr(w), ptr(b), 1, B, S * S, C, 32, 1e-5, 1, ptr(ws), n, ptr(cnt), ptr(stats), ptr_cl(y), stream())
        else:
            fn = lambda: y.copy_(x)
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()          # 50 launches replayed from a graph: device time, not host launch rate
        with torch.cuda.graph(gr):
            for _ in range(50):
                fn()
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 50 * 1e3)
    mb = x.numel() * 4 / 1e6
    print(f"B={B} C={C:4d} S={S:2d}  {mb:6.1f} MB r+w   cluster {res[0]:6.1f} us ({mb / res[0] * 1e3:6.0f} GB/s)   two-launch {res[1]:6.1f} us   copy {res[2]:6.1f} us ({mb / res[2] * 1e3:6.0f} GB/s)", flush=True)
call("gd_group_norm_config", 1)
