"""oracle/loop_oracle.py -- TEST INFRASTRUCTURE ONLY (checker + timed CPU baseline; never imported by geodiffuser_b200).

CPU fp32 restatement of the reference's edit loop for the hot path, built on the layer restatements of oracle/geodiff_oracle.py:
  text2image_ldm_stable          /root/reference/GeoDiffuser/utils/editor.py:65-423
  diffusion_step                 diffusion.py:40-59
  _update_latent                 optimization.py:165-253
  adaptive_optimization_step_*   optimization.py:7-105
  EditProcessor / controllers    attention_processors.py:141-228, 633-664, 931-959
  ddim_loop                      inversion.py:131-196
The UNet is the same random-init SD-1.5-topology module the product is measured on (geodiffuser_b200.unet_sd15, plain torch: it is
the caller, not the path); here it runs in fp32 on the host with attention maps fully materialised, like the reference.

Pinning: `python -m oracle.make_golden_loop` runs this loop with the REAL reference controller classes substituted for
`OracleController` (oracle/ref_import.py) and asserts both agree, then writes tests/golden/loop_*.npz.
"""
import copy
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import geodiff_oracle as O


def _project(attn, hidden_states, encoder_hidden_states):
    """attention_processors.py:164-203 for the module attributes this UNet has (no spatial/group norm, no norm_cross)"""
    is_cross = encoder_hidden_states is not None
    ehs = encoder_hidden_states if is_cross else hidden_states
    q = attn.head_to_batch_dim(attn.to_q(hidden_states))
    k = attn.head_to_batch_dim(attn.to_k(ehs))
    v = attn.head_to_batch_dim(attn.to_v(ehs))
    return q, k, v, is_cross


def _finish(attn, x):
    x = attn.batch_to_head_dim(x)
    return attn.to_out[1](attn.to_out[0](x))


class OracleVanillaProcessor:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        q, k, v, _ = _project(attn, hidden_states, encoder_hidden_states)
        _, o = O.attention(q, k, v, attn.scale)
        return _finish(attn, o)


class OracleEditProcessor:
    def __init__(self, controller, place):
        self.controller, self.place, self.perform_edit = controller, place, True

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        q, k, v, is_cross = _project(attn, hidden_states, encoder_hidden_states)
        out = self.controller(q, k, v, is_cross, self.place, scale=attn.scale)
        return _finish(attn, out)


class OracleController:
    """State machine of AttentionControl (attention_sharing.py:127-144) + AttentionGeometryEdit/Remover.forward"""

    def __init__(self, kind, num_steps, self_replace_steps, obj_edit_step, image_mask, coords512, mask_new_warped, amodal, weights):
        self.kind, self.num_steps, self.obj_edit_step = kind, num_steps, obj_edit_step
        self.num_self_replace = (0, int(num_steps * self_replace_steps))
        self.image_mask = np.asarray(image_mask, np.float32)
        self.image_mask_dilated = O.dilate(self.image_mask, 5) if kind == "remove" else None
        self.coords512, self.mask_new_warped, self.amodal = coords512, mask_new_warped, amodal
        self.default_loss_weights = weights
        self.loss_weight_dict = weights  # aliased, like the reference
        self.cur_step, self.cur_att_layer, self.num_att_layers = 0, 0, -1
        self.coords_base, self.coords_edit, self.use_cfg = (2, 3), (3, 4), True
        self.batch_size = 2
        self.loss = 0.0
        self._cache = {}
        self.initialize_loss_log_dict()

    def initialize_default_loss_weights(self):
        self.loss_weight_dict = self.default_loss_weights

    def initialize_loss_log_dict(self):
        keys = ("sim", "movement", "removal", "smoothness") if self.kind == "edit" else ("sim", "removal", "smoothness")
        self.loss_log_dict = {"self": {k: 0.0 for k in keys}, "cross": {k: 0.0 for k in keys}, "num_layers": 0}

    def _res(self, S):
        if S not in self._cache:
            masks = O.build_masks(self.image_mask, self.mask_new_warped, self.amodal, S) if self.kind == "edit" else None
            coords = O.resize_coords(self.coords512, S) if self.kind == "edit" else None
            self._cache[S] = (masks, coords)
        return self._cache[S]

    def __call__(self, q, k, v, is_cross, place, scale=None):
        h = q.shape[0] // (2 * self.batch_size if self.use_cfg else self.batch_size)
        in_window = is_cross or (self.num_self_replace[0] <= self.cur_step < self.num_self_replace[1])
        if not in_window:
            _, out = O.attention(q, k, v, scale)
        else:
            N = q.shape[1]
            S = int(round(math.sqrt(N)))
            blend = self.cur_step < int(self.num_steps * self.obj_edit_step)
            if self.kind == "edit":
                masks, coords = self._res(S)
                res = O.edit_layer(q, k, v, is_cross, scale, h, self.coords_base, self.coords_edit, masks, coords, self.use_cfg, blend,
                                   weights=self.loss_weight_dict)
            else:
                res = O.remover_layer(q, k, v, is_cross, scale, h, self.coords_base, self.coords_edit, self.image_mask_dilated, self.use_cfg,
                                      blend, weights=self.loss_weight_dict)
            out = res["out"]
            if res["loss"] is not None:
                self.loss = self.loss + res["loss"]
                d = self.loss_log_dict["cross" if is_cross else "self"]
                for key in d:
                    d[key] = d[key] + res["terms"][key].detach()
                self.loss_log_dict["num_layers"] += 1
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
        return out


def register(unet, controller):
    procs, n = {}, 0
    for name in unet.attn_processors.keys():
        place = "mid" if name.startswith("mid_block") else ("up" if name.startswith("up_blocks") else "down")
        procs[name] = OracleEditProcessor(controller, place)
        n += 1
    unet.set_attn_processor(procs)
    controller.num_att_layers = n


def set_mode(controller, coords_base, coords_edit, use_cfg):
    controller.coords_base, controller.coords_edit, controller.use_cfg = coords_base, coords_edit, use_cfg


def clear_loss(controller):
    controller.loss = 0.0
    controller.initialize_loss_log_dict()


def log_to_float(d):
    out = {"self": {}, "cross": {}}
    for att in ("self", "cross"):
        for k, v in d[att].items():
            out[att][k] = float(v)
    out["num_layers"] = d["num_layers"]
    return out


def adaptive_step(controller, i, skip, log, num_ddim_steps, removal_loss_value_in, kind):
    """optimization.py:7-105"""
    frac = i / num_ddim_steps
    if frac < 0.4:
        remaining = int((0.4 - frac) * num_ddim_steps / skip)
        expected = removal_loss_value_in / (1.25) ** remaining
        cur = log["self"]["removal"]
        if expected < cur:
            controller.loss_weight_dict["self"]["removal"] *= 1.3
        elif 2.5 * expected > cur:
            controller.loss_weight_dict["self"]["removal"] /= (2.0 if kind == "edit" else 2.5)
    elif 0.4 < frac < 0.8:
        if (removal_loss_value_in - 0.3) < log["self"]["removal"]:
            controller.loss_weight_dict["self"]["removal"] *= 2.0
        else:
            controller.initialize_default_loss_weights()
    else:
        controller.initialize_default_loss_weights()


def geometry_inputs(kind, synth):
    """A1-A3 on the host: correspondence field, amodal mask, warped mask (same as oracle/make_golden.py:geometry_case)"""
    image, depth, mask, T = synth.edit_inputs(kind)
    g = O.corr_build(depth.copy(), mask.copy(), T)
    amodal = O.erode3(O.mesh_mask(g["coords"], g["mask"]))
    idx512, _, d2 = O.splat_index(g["coords"][None])
    # editor.py:147-149 warps controller.image_mask, which the remover's constructor has dilated by 5 px (attention_processors.py:986)
    src = O.dilate(mask.astype(np.float32), 5) if kind == "remove" else mask.astype(np.float32)
    mnw = O.binarize(O.splat_composite(np.asarray(src, np.float32)[None, None], idx512, d2))[0, 0]
    return dict(coords=g["coords"], mask=mask.astype(np.float32), amodal=amodal, mnw=mnw)


def inverse_step(x, eps, t, alphas, num_steps, num_train=1000):
    ratio = num_train // num_steps
    cur = min(t - ratio, num_train - 1)
    a_t = alphas[cur] if cur >= 0 else torch.tensor(1.0)
    a_n = alphas[t]
    x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    return a_n ** 0.5 * x0 + (1 - a_n) ** 0.5 * eps


@torch.no_grad()
def ddim_inversion(unet, x0, context, guidance, num_steps):
    unet.set_attn_processor(OracleVanillaProcessor())
    al, ts = O.ddim_alphas(), O.ddim_timesteps(num_steps)
    lat, out = x0.clone(), [x0]
    for t in ts[::-1].tolist():
        eps = unet(torch.cat([lat] * 2), t, encoder_hidden_states=context)["sample"]
        eu, ec = eps.chunk(2)
        lat = inverse_step(lat, O.cfg_combine(eu, ec, guidance), t, al, num_steps)
        out.append(lat)
    return out


def edit_loop(unet, kind, geo, text, uncond, x_t, ddim_latents, hp, num_steps, make_controller=None, step_limit=None, timings=None):
    """editor.py:65-423 on the host.  unet: fp32 CPU module.  Returns (latents, log)."""
    import time

    ctl_kind = "remove" if kind == "remove" else "edit"
    weights = copy.deepcopy(hp["loss_weights_dict"])
    if make_controller is None:
        controller = OracleController(ctl_kind, num_steps, hp["self_replace_steps"], hp["obj_edit_step"], geo["mask"], geo["coords"], geo["mnw"],
                                      geo["amodal"], weights)
        register(unet, controller)
    else:
        controller = make_controller(weights)
    al, ts = O.ddim_alphas(), O.ddim_timesteps(num_steps)
    latents = x_t[:1].expand(2, *x_t.shape[1:]).clone()
    context_save, log = None, {}
    skip, n_t, gs = hp["skip_optim_steps"], num_steps, hp["guidance_scale"]
    for i, t in enumerate(ts.tolist()):
        if step_limit is not None and i >= step_limit:
            break
        context = torch.cat([uncond, text])
        clear_loss(controller)
        if (i < hp["optimize_steps"] * n_t) and (i % skip == 0):
            t0 = time.perf_counter()
            l_eff = hp["lr"] * (50 - i) * skip * (50 / (num_steps + 1e-8))
            set_mode(controller, (0, 1), (1, 2), False)
            lat_in = latents.detach().float().requires_grad_(True)
            orig_norm = O.norm_tensor(lat_in[-1:].detach()).item()
            ctx_in = (context if context_save is None else context_save).detach().float().requires_grad_(True)
            with torch.enable_grad():
                unet(lat_in, t, encoder_hidden_states=ctx_in[2:])
                loss = controller.loss
                g_lat, g_ctx = torch.autograd.grad(loss, [lat_in, ctx_in])
            new_lat, new_ctx = O.update_latent(lat_in.detach(), g_lat, l_eff, geo["mnw"], ctx_in.detach(), g_ctx)
            out_log = log_to_float(controller.loss_log_dict)
            adaptive_step(controller, i, skip, out_log, num_steps, hp.get("removal_loss_value_in", -1.5), ctl_kind)
            out_log["loss"] = float(loss)
            log[i] = out_log
            clear_loss(controller)
            controller.cur_step -= 1
            if hp["optimize_latents"]:
                latents = new_lat.detach().clone()
                latents[-1:] = latents[-1:] * orig_norm / O.norm_tensor(latents[-1:]).item()
            if hp["optimize_embeddings"]:
                context = new_ctx.detach()
                context_save = context
            if timings is not None:
                timings.setdefault("opt", []).append(time.perf_counter() - t0)
        elif context_save is not None:
            context = context_save
        t0 = time.perf_counter()
        with torch.no_grad():
            set_mode(controller, (2, 3), (3, 4), True)
            eps = unet(torch.cat([latents] * 2), t, encoder_hidden_states=context)["sample"]
            eu, ec = eps.chunk(2)
            latents = O.ddim_step(latents, O.cfg_combine(eu, ec, gs), t, al, num_steps)
            if ddim_latents is not None:
                latents = torch.cat([ddim_latents[len(ddim_latents) - 2 - i], latents[-1:]], 0)
            if ctl_kind == "edit" and i < n_t * hp["latent_replace"]:
                S = latents.shape[-1]
                tc = O.resize_coords(geo["coords"], S)
                m = O.binarize(O.resize_bilinear(geo["mnw"][None], S))[0]
                warped = torch.from_numpy(O.warp_grid_edit(latents[-2:-1].numpy(), tc[None]))
                mt = torch.from_numpy(m)[None, None]
                latents = torch.cat([latents[:-1], latents[-1:] * (1 - mt) + mt * warped], 0)
        if timings is not None:
            timings.setdefault("cfg", []).append(time.perf_counter() - t0)
    return latents, log
