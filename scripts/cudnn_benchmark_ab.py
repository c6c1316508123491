import sys, time, torch
sys.path.insert(0, ".")
import os
if os.environ.get("GD_CUDNN_BENCHMARK") == "1":
    torch.backends.cudnn.benchmark = True
from geodiffuser_b200 import editor, unet_sd15
model = unet_sd15.build_model("cuda")
req = editor.synthetic_request("rotate3d")
staged, _ = editor.stage_inputs(req["depth"], req["image_mask"], req["text_embeddings"], req["uncond_embeddings"], req["x0"], model.device)
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    editor.run_edit(model, staged, req["transform_in"], req["edit_type"])
    torch.cuda.synchronize(); print(f"edit {it}: {(time.perf_counter()-t0)*1e3:.0f} ms", flush=True)
