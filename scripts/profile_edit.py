"""Profiling driver (run under ncu, never a bench number): one warm-up edit (graphs captured, caches built), then one edit between
cudaProfilerStart / cudaProfilerStop -- which, unlike an NVTX range, also covers the kernels launched from the autograd thread and the
kernel nodes of the replayed CUDA graphs.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python scripts/profile_edit.py --steps 10
"""
import sys
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import unet_sd15, editor

steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 10
kind = sys.argv[sys.argv.index("--kind") + 1] if "--kind" in sys.argv else "rotate3d"
model = unet_sd15.build_model("cuda")
for _ in range(2):
    editor.perform_synthetic_edit(model, kind, num_ddim_steps=steps)
torch.cuda.synchronize()
torch.cuda.profiler.start()
editor.perform_synthetic_edit(model, kind, num_ddim_steps=steps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
