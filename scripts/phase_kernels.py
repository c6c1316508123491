"""Per-phase kernel table (torch.profiler, eager, graphs off): which kernels one inversion pass, one CFG pass and one optimisation pass
of the 50-step edit spend their device time in.  Diagnostic only -- never a bench number.
    python scripts/phase_kernels.py [kind] > gpurun_out/phase_kernels.log
"""
import collections
import sys

import torch

sys.path.insert(0, ".")
from geodiffuser_b200 import editor, graphs, unet_sd15  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "rotate3d"
graphs.ENABLED = False
graphs.GRAD_ENABLED = False
model = unet_sd15.build_model("cuda")
count = collections.Counter()
AT = {"inversion": 5, "cfg": 5, "opt": 3}
ARMED = [False]


def wrap(name, fn):
    def w(*a, **k):
        count[name] += 1
        if not ARMED[0] or count[name] != AT[name]:
            return fn(*a, **k)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            r = fn(*a, **k)
            torch.cuda.synchronize()
        rows = collections.defaultdict(lambda: [0.0, 0])
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                rows[ev.name[:110]][0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
                rows[ev.name[:110]][1] += 1
        tot = sum(v[0] for v in rows.values())
        n = sum(v[1] for v in rows.values())
        print(f"=== {name} pass: {tot / 1e3:.2f} ms of kernels, {n} launches")
        for k_, v in sorted(rows.items(), key=lambda kv: -kv[1][0])[:45]:
            print(f"  {v[0] / 1e3:8.3f} ms {100 * v[0] / tot:5.1f}%  n={v[1]:4d}  avg {v[0] / v[1]:7.1f} us  {k_}")
        sys.stdout.flush()
        return r
    return w


graphs.inversion_pass = wrap("inversion", graphs.inversion_pass)
graphs.edit_pass = wrap("cfg", graphs.edit_pass)
graphs.grad_pass = wrap("opt", graphs.grad_pass)
editor.perform_synthetic_edit(model, kind, num_ddim_steps=50)
count.clear()
ARMED[0] = True
editor.perform_synthetic_edit(model, kind, num_ddim_steps=50)
