"""scratch: run one synthetic edit end to end on cuda:0 and print timings (not a bench)"""
import sys, time
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import unet_sd15, editor, _lib

tiny = "--tiny" in sys.argv
steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 50
kind = sys.argv[sys.argv.index("--kind") + 1] if "--kind" in sys.argv else "rotate3d"
t0 = time.time()
model = unet_sd15.build_model("cuda", tiny=tiny)
torch.cuda.synchronize()
print("model built", time.time() - t0, flush=True)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.time(); l0 = _lib.LAUNCHES
    lat, log = editor.perform_synthetic_edit(model, kind, num_ddim_steps=steps, return_log=True)
    torch.cuda.synchronize()
    print(f"edit {it}: {time.time()-t0:.2f}s launches={_lib.LAUNCHES-l0} finite={bool(torch.isfinite(lat).all())} norm={float(lat[-1].norm()):.3f}", flush=True)
    for i in sorted(log)[:3]:
        print(i, {k: round(v, 4) for k, v in log[i]["self"].items()}, round(log[i]["loss"], 4))
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)
