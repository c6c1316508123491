"""Profiling driver (run under ncu, never a bench number): one warm-up edit, then one edit inside the NVTX range `timed`.
    ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python scripts/profile_edit.py --steps 10
"""
import sys
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import unet_sd15, editor

steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 10
kind = sys.argv[sys.argv.index("--kind") + 1] if "--kind" in sys.argv else "rotate3d"
model = unet_sd15.build_model("cuda")
editor.perform_synthetic_edit(model, kind, num_ddim_steps=steps)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
editor.perform_synthetic_edit(model, kind, num_ddim_steps=steps)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
