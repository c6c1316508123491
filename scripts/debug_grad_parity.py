"""debug (GPU box): ONE optimisation pass (loss + d loss / d latents, d loss / d context) of the tiny UNet, ours vs the CPU oracle, at a state
where the edit latent differs from the reference latent by `eps` * noise (eps = 0 is the degenerate step-0 state: L1 terms at their kink)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import unet_sd15, editor, diffusion, synth
from geodiffuser_b200.attention_processors import register_attention_control_diffusers, set_attn_processor_for_edit
from geodiffuser_b200.editor import EXP_PARAMS, synthetic_embeddings
from oracle import loop_oracle as LO, geodiff_oracle as O
import copy

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-12))
cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
kinds = sys.argv[1:] or ["translate2d", "remove"]
model_c = unet_sd15.build_model("cpu", tiny=True)
unet = model_c.unet.float()
model = unet_sd15.build_model("cuda", tiny=True)
text, uncond, x0 = synthetic_embeddings(device="cpu")
g = torch.Generator().manual_seed(5)
noise = torch.randn(1, 4, 64, 64, generator=g)
cnoise = torch.randn(1, 77, 768, generator=g)
for kind in kinds:
    edit_type = "geometry_remover" if kind == "remove" else "geometry_editor"
    hp = dict(EXP_PARAMS[edit_type])
    geo = LO.geometry_inputs(kind, synth)
    req = editor.synthetic_request(kind, pin=False)
    staged, _ = editor.stage_inputs(req["depth"], req["image_mask"], req["text_embeddings"], req["uncond_embeddings"], req["x0"], model.device)
    for eps in (0.0, 0.05, 0.3):
        lat = torch.cat([x0, x0 + eps * noise])
        ctx = torch.cat([text[:1], text[:1] + eps * cnoise])
        step_i, num_steps = 2, 10
        ts = O.ddim_timesteps(num_steps).tolist()
        t = ts[step_i]
        # oracle
        ctl_kind = "remove" if kind == "remove" else "edit"
        oc = LO.OracleController(ctl_kind, num_steps, hp["self_replace_steps"], hp["obj_edit_step"], geo["mask"], geo["coords"], geo["mnw"], geo["amodal"],
                                 copy.deepcopy(hp["loss_weights_dict"]))
        LO.register(unet, oc)
        oc.cur_step = step_i
        LO.set_mode(oc, (0, 1), (1, 2), False)
        li, ci = lat.clone().requires_grad_(True), ctx.clone().requires_grad_(True)
        with torch.enable_grad():
            unet(li, t, encoder_hidden_states=ci)
            go_l, go_c = torch.autograd.grad(oc.loss, [li, ci], allow_unused=True)
        if go_c is None: go_c = torch.zeros_like(ci)
        log_o = LO.log_to_float(oc.loss_log_dict)
        for dt in (torch.float32, torch.bfloat16):
            diffusion.set_body_dtype(dt)
            c, tc = editor.make_controller(model, staged, req["transform_in"], edit_type, hp, num_steps)
            register_attention_control_diffusers(model, c, tc)
            c._ensure_mask_new_warped(tc, model.device)
            c.cur_step = step_i
            model.scheduler.set_timesteps(num_steps)
            set_attn_processor_for_edit(model, coords_base=(0, 1), coords_edit=(1, 2), use_cfg=False)
            lg, cg = lat.cuda().requires_grad_(True), ctx.cuda().requires_grad_(True)
            editor.clear_controller_loss(c)
            with torch.enable_grad():
                diffusion.diffusion_step(model, c, lg, cg, t, 3.0, transform_coords=tc, use_cfg=False, return_noise=True)
                g_l, g_c = torch.autograd.grad(c.loss, [lg, cg], allow_unused=True)
            if g_c is None: g_c = torch.zeros_like(cg)
            log = editor.convert_loss_log_to_numpy(c.loss_log_dict)
            print(f"{kind} eps={eps} body={str(dt)[6:]}: loss {float(c.loss):.4f}/{float(oc.loss):.4f}  dL/dlat[edit] relerr {rel(g_l[-1].cpu(), go_l[-1]):.3e} cos {cos(g_l[-1].cpu(), go_l[-1]):.5f}"
                  f"  dL/dctx[edit] relerr {rel(g_c[-1].cpu(), go_c[-1]):.3e} cos {cos(g_c[-1].cpu(), go_c[-1]):.5f}  |g| {float(go_l[-1].abs().max()):.3g} {float(go_c[-1].abs().max()):.3g}")
            print("     self ", {k: (round(v, 5), round(log_o['self'][k], 5)) for k, v in log["self"].items()})
