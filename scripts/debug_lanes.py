"""debug: edits through runner.EditWorkers vs the same edits alone on the tiny model.  argv: share (0|1)"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from geodiffuser_b200 import editor, graphs, runner, unet_sd15

graphs.SHARE_GRAD_GRAPHS = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True


def psnr(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    mse = ((a - ref) ** 2).mean()
    peak = ref.max() - ref.min()
    return float(10 * np.log10(peak * peak / max(mse, 1e-30)))


fn = lambda m, k: editor.perform_synthetic_edit(m, k, num_ddim_steps=6)
kinds3 = ["rotate3d", "remove", "translate2d"]
share = graphs.SHARE_GRAD_GRAPHS
graphs.SHARE_GRAD_GRAPHS = False
model0 = unet_sd15.build_model("cuda", tiny=True)
alone = {k: fn(model0, k).float().cpu() for k in kinds3}      # the baseline: no graph sharing, main thread, default stream
graphs.SHARE_GRAD_GRAPHS = share
model = unet_sd15.build_model("cuda", tiny=True)
if len(sys.argv) > 2 and sys.argv[2] == "stream":
    st = torch.cuda.Stream()
    for r in range(2):
        for k in kinds3:
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                o = fn(model, k)
            torch.cuda.synchronize()
            print(f"share={share} main thread, side stream, round {r} {k}: {psnr(o.float().cpu()[1].numpy(), alone[k][1].numpy()):.1f} dB", flush=True)
for r in range(3):
    for k in kinds3:
        o = fn(model, k).float().cpu()
        print(f"share={graphs.SHARE_GRAD_GRAPHS} main thread round {r} {k}: {psnr(o[1].numpy(), alone[k][1].numpy()):.1f} dB", flush=True)
w1 = runner.EditWorkers(model, lanes=1)
for r in range(2):
    for k in kinds3:
        o = w1.map(fn, [k])[0].float().cpu()
        torch.cuda.synchronize()
        print(f"share={graphs.SHARE_GRAD_GRAPHS} one lane (model 0, lane thread) round {r} {k}: {psnr(o[1].numpy(), alone[k][1].numpy()):.1f} dB", flush=True)
w2 = runner.EditWorkers(model, lanes=2)
for r in range(2):
    for k in kinds3:          # one at a time on lane 1 (the replica)
        outs = w2.map(lambda m, kk: None if kk is None else fn(m, kk), [None, k])
        torch.cuda.synchronize()
        o = outs[1].float().cpu()
        print(f"share={graphs.SHARE_GRAD_GRAPHS} replica lane alone round {r} {k}: {psnr(o[1].numpy(), alone[k][1].numpy()):.1f} dB", flush=True)
kinds = ["rotate3d", "remove", "translate2d", "rotate3d", "remove", "translate2d"]
for r in range(2):
    outs = w2.map(fn, kinds)
    torch.cuda.synchronize()
    for i, (k, o) in enumerate(zip(kinds, outs)):
        o = o.float().cpu()
        print(f"share={graphs.SHARE_GRAD_GRAPHS} two lanes concurrent round {r} job {i} (lane {i % 2}) {k}: {psnr(o[1].numpy(), alone[k][1].numpy()):.1f} dB", flush=True)
