"""debug: is an edit that replays a persistent optimisation-pass graph reproducible from edit to edit?  argv: share(0|1) corr_sm100(0|1) reuse_ref(0|1)"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from geodiffuser_b200 import editor, graphs, unet_sd15, functional as Fn

graphs.SHARE_GRAD_GRAPHS = bool(int(sys.argv[1])); Fn.CORR_SM100 = bool(int(sys.argv[2])); editor.REUSE_REFERENCE_OF_OPT_PASS = bool(int(sys.argv[3]))
model = unet_sd15.build_model("cuda", tiny=True)
def run(seed, steps=6):
    req = editor.synthetic_request("rotate3d", seed=seed, pin=False)
    staged, _ = editor.stage_inputs(req["depth"], req["image_mask"], req["text_embeddings"], req["uncond_embeddings"], req["x0"], model.device)
    lat, log = editor.run_edit(model, staged, req["transform_in"], req["edit_type"], num_ddim_steps=steps, return_log=True)
    return lat.float().cpu(), log
r = [run(11) for _ in range(5)]
for i in range(1, 5):
    la, lb = r[i - 1][1], r[i][1]
    print(f"share={sys.argv[1]} corr_sm100={sys.argv[2]} reuse_ref={sys.argv[3]} edit {i-1} vs {i}: latents equal {torch.equal(r[i-1][0], r[i][0])}; "
          + " ".join(f"step{k}: loss {la[k]['loss']:.7f}/{lb[k]['loss']:.7f} rem {la[k]['self']['removal']:.7f}/{lb[k]['self']['removal']:.7f}" for k in sorted(la)))
