"""debug: two edit lanes on the tiny model; argv: capture_error_mode (thread_local|relaxed|global) cudnn_benchmark (0|1)"""
import sys, threading, time, traceback
sys.path.insert(0, ".")
import torch
from geodiffuser_b200 import editor, graphs, runner, unet_sd15

mode, bench = sys.argv[1], int(sys.argv[2])
graphs.CAPTURE_ERROR_MODE = mode
model = unet_sd15.build_model("cuda", tiny=True)
torch.backends.cudnn.benchmark = bool(bench)
kinds = ["rotate3d", "remove", "translate2d", "rotate3d", "remove", "translate2d"]
w = runner.EditWorkers(model, lanes=2)
t0 = time.time()
try:
    for r in range(3):
        outs = w.map(lambda m, k: editor.perform_synthetic_edit(m, k, num_ddim_steps=6), kinds)
        torch.cuda.synchronize()
        print(f"mode={mode} benchmark={bench} round {r}: ok, {time.time() - t0:.1f}s, finite={all(bool(torch.isfinite(o).all()) for o in outs)}", flush=True)
except BaseException:
    traceback.print_exc()
    print(f"mode={mode} benchmark={bench}: FAILED", flush=True)
