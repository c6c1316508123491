"""debug (GPU box): tiny loop vs golden for both body dtypes"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import unet_sd15, editor, diffusion

def psnr(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    mse = ((a - ref) ** 2).mean(); peak = ref.max() - ref.min()
    return float(10 * np.log10(peak * peak / max(mse, 1e-30)))

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model = unet_sd15.build_model("cuda", tiny=True)
for dt in (torch.float32, torch.bfloat16):
    diffusion.set_body_dtype(dt)
    for kind in ("translate2d", "rotate3d", "remove"):
        z = np.load(f"tests/golden/loop_{kind}_tiny.npz")
        lat, log = editor.perform_synthetic_edit(model, kind, num_ddim_steps=int(z["meta"][1]), return_log=True)
        lat = lat.float().cpu().numpy()
        print(f"{dt} {kind}: PSNR ref {psnr(lat[0], z['latents'][0]):.1f} edit {psnr(lat[1], z['latents'][1]):.1f}")
        for i in sorted(log):
            s = f"   step {i}: loss {log[i]['loss']:.5f} / {float(z[f'log{i}_loss']):.5f}"
            for att in ("self", "cross"):
                for k, v in log[i][att].items():
                    s += f" {att[0]}.{k} {v:.5f}/{float(z[f'log{i}_{att}_{k}']):.5f}"
            print(s)
