"""debug (GPU box): trace the first steps of the tiny loop (fp32 body) against the CPU oracle loop: gradients, updated latents, CFG step, warp."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import unet_sd15, editor, diffusion, optimization, synth
from geodiffuser_b200.editor import EXP_PARAMS, synthetic_embeddings
from oracle import loop_oracle as LO, geodiff_oracle as O

kind = sys.argv[1] if len(sys.argv) > 1 else "translate2d"
LIMIT = 3
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-12))

# ---- oracle
tr_o = []
_upd, _step, _warp = O.update_latent, O.ddim_step, O.warp_grid_edit
def upd(lat, g_lat, l_eff, mnw, ctx, g_ctx):
    out = _upd(lat, g_lat, l_eff, mnw, ctx, g_ctx)
    tr_o.append(("g_lat", g_lat.clone())); tr_o.append(("g_ctx", g_ctx.clone())); tr_o.append(("new_lat", out[0].clone())); tr_o.append(("new_ctx", out[1].clone()))
    return out
def stp(*a, **k):
    out = _step(*a, **k); tr_o.append(("cfg_step", out.clone())); return out
def wrp(src, coords, *a, **k):
    out = _warp(src, coords, *a, **k)
    if src.shape[1] == 4: tr_o.append(("warp", torch.from_numpy(np.asarray(out)).clone()))
    return out
O.update_latent, O.ddim_step, O.warp_grid_edit = upd, stp, wrp
model_c = unet_sd15.build_model("cpu", tiny=True)
unet = model_c.unet.float()
edit_type = "geometry_remover" if kind == "remove" else "geometry_editor"
hp = dict(EXP_PARAMS[edit_type])
geo = LO.geometry_inputs(kind, synth)
text, uncond, x0 = synthetic_embeddings(device="cpu")
ddim = LO.ddim_inversion(unet, x0, torch.cat([uncond[:1], text[:1]]), hp["guidance_scale"], 10)
tr_o.clear()
lat_o, log_o = LO.edit_loop(unet, kind, geo, text, uncond, ddim[-1], ddim, hp, 10, step_limit=LIMIT)

# ---- ours
tr = []
diffusion.set_body_dtype(torch.float32)
model = unet_sd15.build_model("cuda", tiny=True)
_apply = optimization.apply_latent_update
def apply(latents, g, step, mask, context, gc):
    out = _apply(latents, g, step, mask, context, gc)
    tr.append(("g_lat", g.detach().cpu().clone())); tr.append(("g_ctx", gc.detach().cpu().clone())); tr.append(("new_lat", out[0].detach().cpu().clone())); tr.append(("new_ctx", out[1].detach().cpu().clone()))
    return out
optimization.apply_latent_update = apply
_scfg = model.scheduler.step_cfg
def scfg(*a, **k):
    out = _scfg(*a, **k); tr.append(("cfg_step", out.detach().cpu().clone())); return out
model.scheduler.step_cfg = scfg
_lw = editor._latent_warp_replace
def lw(controller, latents, tc, fast=False):
    out = _lw(controller, latents, tc, fast)
    tr.append(("warp_out", out.detach().cpu().clone())); return out
editor._latent_warp_replace = lw
# run with the oracle's inversion trajectory so both loops start from the same x_T
req = editor.synthetic_request(kind, pin=False)
staged, _ = editor.stage_inputs(req["depth"], req["image_mask"], req["text_embeddings"], req["uncond_embeddings"], req["x0"], model.device)
_inv = editor.ddim_inversion_loop
editor.ddim_inversion_loop = lambda *a, **k: [d.cuda() for d in ddim]
class Stop(Exception): pass
cnt = [0]
def prog(x):
    cnt[0] += 1
    if cnt[0] >= LIMIT: raise Stop
import geodiffuser_b200.editor as E
_t2i = E.text2image_ldm_stable
def t2i(*a, **k):
    k["progress"] = prog
    return _t2i(*a, **k)
E.text2image_ldm_stable = t2i
try:
    editor.run_edit(model, staged, req["transform_in"], req["edit_type"], num_ddim_steps=10)
except Stop:
    pass
names_o = [n for n, _ in tr_o]; names = [n for n, _ in tr]
print("oracle trace:", names_o); print("ours trace:", names)
io = 0
for n, v in tr:
    if n == "warp_out":
        # compare with oracle's latents after warp: reconstruct = last cfg_step then warp... compare the warped source only
        continue
    while io < len(tr_o) and tr_o[io][0] != n: io += 1
    if io >= len(tr_o): break
    ref = tr_o[io][1]; io += 1
    if n in ("g_lat", "new_lat", "cfg_step"):
        print(f"{n:9s} edit-sample relerr {rel(v[-1].float(), ref[-1]):.3e}  (|ref|max {float(ref[-1].abs().max()):.4g}) base-sample relerr {rel(v[0].float(), ref[0]) if float(ref[0].abs().max())>0 else 0:.3e}")
    else:
        print(f"{n:9s} edit-sample relerr {rel(v[-1].float(), ref[-1]):.3e}  (|ref|max {float(ref[-1].abs().max()):.4g})")
print("log ours vs oracle: ", {i: round(v["loss"], 4) for i, v in log_o.items()})
