import sys, torch, numpy as np
sys.path.insert(0, ".")
from geodiffuser_b200._lib import call, ptr, stream
from geodiffuser_b200 import geometry as G, synth
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
S = 64; N = S * S
m = torch.zeros(S, S, device="cuda"); m[12:30, 16:36] = 1; m = m.reshape(-1).contiguous()
idx = torch.empty(N, 4, device="cuda", dtype=torch.int32); val = torch.empty(N, 4, device="cuda"); w = torch.empty(N, device="cuda")
print("amodal_knn S=64: %.1f us" % timed(lambda: call("gd_amodal_knn", ptr(m), S, ptr(idx), ptr(val), ptr(w), stream())))
image, depth, mask, T = synth.edit_inputs("rotate3d")
print("correspondence_field (pixel2cam + centroid + project, host part included): %.1f us" % timed(lambda: G.correspondence_field(depth.copy(), mask.copy(), T), 5))
