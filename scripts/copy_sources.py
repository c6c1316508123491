"""Which Python lines launch the strided-copy kernels of one inversion pass (torch.profiler with stacks).  Diagnostic only."""
import collections, sys
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import editor, graphs, unet_sd15
from torch.profiler import ProfilerActivity, profile

graphs.ENABLED = False
model = unet_sd15.build_model("cuda")
req = editor.synthetic_request("rotate3d")
staged, _ = editor.stage_inputs(req["depth"], req["image_mask"], req["text_embeddings"], req["uncond_embeddings"], req["x0"], model.device)
lat = torch.randn(2, 4, 64, 64, device="cuda")
ctx = torch.randn(2, 77, 768, device="cuda")
for _ in range(2):
    graphs.inversion_pass(model, lat, 500, ctx)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    graphs.inversion_pass(model, lat, 500, ctx)
    torch.cuda.synchronize()
rows = collections.Counter()
for ev in prof.key_averages(group_by_stack_n=12):
    if ev.key in ("aten::copy_", "aten::clone", "aten::contiguous", "aten::add", "aten::add_", "aten::mul", "aten::gelu", "aten::cat"):
        frames = [f for f in ev.stack if "geodiffuser_b200" in f]
        rows[(ev.key, frames[0] if frames else "?")] += ev.count
for (k, f), n in rows.most_common(60):
    print(f"{n:5d}  {k:18s} {f}")
