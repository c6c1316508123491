"""The DDIM edit loop around the hot path, restated from /root/reference/GeoDiffuser/utils/editor.py:65-423
(`text2image_ldm_stable`), inversion.py:131-196 (`ddim_loop`) and generic.py:34-60 (loss-log helpers), with the per-edit-type
hyper-parameters of large_scale_editor.py:199-299 (`perform_exp`).  Text encoder / VAE / image post-processing are out of scope
(SURVEY 2 rows 8, 12): prompts are replaced by synthetic context embeddings and images by synthetic latents.

Control flow, step conditions, learning-rate schedule, norm preservation, reference-latent replacement and the latent warp are the
reference's; every tensor op on the path is a C-ABI CUDA launch (attention layers, DDIM step, update, warp), the UNet body is stock torch.
"""
import numpy as np
import torch

from . import geometry, graphs, synth
from .attention_processors import (AttentionGeometryEdit, AttentionGeometryRemover, VanillaAttentionProcessor,
                                   register_attention_control_diffusers, set_attn_processor_for_edit)
from .diffusion import body_autocast, diffusion_step
from .optimization import (_update_latent, adaptive_optimization_step_editing, adaptive_optimization_step_remover, apply_latent_update,
                           norm_tensor, norm_tensor_dev, rescale_to_norm_)
from ._lib import call, ptr, stream

NUM_DDIM_STEPS = 50
IMAGE_SIZE = 512
SEED = 1234  # editor.py:47
SKIP_DEAD_UNCOND_REFERENCE = True   # diffusion.diffusion_step(skip_uncond_reference=...): result-preserving, 1/4 of every CFG pass
REUSE_REFERENCE_OF_OPT_PASS = True  # on a timestep with an optimisation pass, the CFG pass reuses the reference sample's per-layer
                                    # K / V / warped-stream output from that pass instead of re-evaluating the sample (1/3 of those CFG passes)


def clear_controller_loss(controller):
    """generic.py:41-47"""
    controller.loss = 0.0
    if controller.loss_log_dict is not None:
        controller.initialize_loss_log_dict()


def convert_loss_log_to_numpy(loss_log_dict):
    """generic.py:50-60 (one small D2H per logged term, as the reference's .item() calls)"""
    out = {"self": {}, "cross": {}}
    for att_type in loss_log_dict:
        if att_type in ("self", "cross"):
            for key in loss_log_dict[att_type]:
                v = loss_log_dict[att_type][key]
                out[att_type][key] = v.item() if torch.is_tensor(v) else float(v)
        else:
            out[att_type] = loss_log_dict[att_type]
    return out


@torch.no_grad()
def ddim_inversion_loop(model, latent, context, guidance_scale=3.0, num_ddim_steps=NUM_DDIM_STEPS):
    """inversion.py:131-196: x_0 -> [x_0, x_1, ..., x_T] with classifier-free guidance and vanilla attention (N1 in SURVEY 8(f))."""
    model.unet.set_attn_processor(VanillaAttentionProcessor())
    sched = model.scheduler
    sched.set_timesteps(num_ddim_steps)
    all_latent = [latent]
    latents = latent.clone().detach().float()
    for t in reversed(sched.timesteps.tolist()):  # 0, 20, ..., 980 (DDIMInverseScheduler, leading spacing)
        with body_autocast():
            noise_pred = graphs.inversion_pass(model, torch.cat([latents] * 2), t, context)
        eu, ec = noise_pred.chunk(2)
        latents = sched.next_step_cfg(eu, ec, guidance_scale, t, latents)
        all_latent.append(latents.detach())
    return all_latent


@torch.no_grad()
def text2image_ldm_stable(model, prompt, controller, num_inference_steps=50, guidance_scale=7.5, generator=None, latent=None,
                          uncond_embeddings=None, text_embeddings=None, start_time=50, return_type="latent",
                          transform_coordinates=None, mask_obj=None, optimize_steps=0.2, latent_replace=0.2, lr=0.0,
                          optimize_embeddings=False, optimize_latents=False, ddim_latents=None, ddim_noise=None,
                          edit_type="geometry_editor", fast_start_steps=0.0, num_first_optim_steps=5, use_adaptive_optimization=True,
                          adain_latents_steps=0.95, use_optimizer=False, removal_loss_value_in=-1.5, skip_optim_steps=2,
                          progress=None):
    """editor.py:65-423.  `uncond_embeddings` / `text_embeddings`: (B,77,768) tensors standing in for the CLIP encodings (:106-121).
    Returns (latents, x_T, global_loss_log_dict)."""
    if edit_type not in ("geometry_editor", "geometry_remover"):
        raise NotImplementedError(edit_type)  # the stitch controllers do not exist in the reference either
    if use_optimizer:
        raise NotImplementedError("use_optimizer is never forwarded by the reference drivers (editor.py:651-652)")
    global_loss_log_dict = {}
    batch_size = len(prompt)
    device = model.device
    register_attention_control_diffusers(model, controller, transform_coordinates)
    text_embeddings = text_embeddings.to(device).float()
    uncond_embeddings_ = uncond_embeddings.to(device).float()
    latents = latent[:1].expand(batch_size, *latent.shape[1:]).to(device).float().contiguous()
    model.scheduler.set_timesteps(num_inference_steps)
    timesteps = model.scheduler.timesteps[-start_time:].tolist()
    n_t = len(timesteps)
    context_save = None
    first_optim_complete = False
    if transform_coordinates is not None and hasattr(controller, "_ensure_mask_new_warped"):
        controller._ensure_mask_new_warped(transform_coordinates, device)  # editor.py:147-149
    is_remover = type(controller).__name__ == "AttentionGeometryRemover"

    for i, t in enumerate(timesteps):
        context = torch.cat([uncond_embeddings_, text_embeddings])
        clear_controller_loss(controller)
        do_opt = (i < optimize_steps * n_t) and (i % skip_optim_steps == 0) and (i >= fast_start_steps * n_t)
        if do_opt:
            if not first_optim_complete and fast_start_steps > 0.0:
                num_optim_steps, first_optim_complete = num_first_optim_steps, True
            else:
                num_optim_steps = 1
            best_loss, best_latents, best_context = 1e8, None, None
            l_eff = lr * (50 - i) * skip_optim_steps * (50 / (NUM_DDIM_STEPS + 1e-8))  # editor.py:207
            set_attn_processor_for_edit(model, coords_base=(0, 1), coords_edit=(1, 2), use_cfg=False)
            reuse = REUSE_REFERENCE_OF_OPT_PASS and SKIP_DEAD_UNCOND_REFERENCE and ddim_latents is not None and hasattr(controller, "base_mode")
            controller.base_mode = "write" if reuse else None
            latents_in = latents.detach().float().requires_grad_(True)
            orig_norm = norm_tensor_dev(latents_in[-1:].detach())   # stays on the device (no host sync in front of the optimisation pass)
            context_in = (context if context_save is None else context_save).detach().float().requires_grad_(True)
            for _ in range(num_optim_steps):
                with torch.enable_grad():
                    # forward with the loss-bearing layers + autograd.grad (diffusion_step(use_cfg=False) + optimization.py:201); replayed from
                    # a CUDA graph after the first two passes of an edit (graphs.grad_pass)
                    g_lat, g_ctx = graphs.grad_pass(model, controller, latents_in, context_in, t)
                    loss_val = controller.loss.detach().item()
                    if loss_val < best_loss:
                        best_latents, best_context, best_loss = latents_in, context_in, loss_val
                    latents_in, context_new = apply_latent_update(
                        latents_in, g_lat if g_lat is not None else torch.zeros_like(latents_in), l_eff, controller.mask_new_warped[:1], context_in,
                        g_ctx if g_ctx is not None else torch.zeros_like(context_in))
                    if num_optim_steps == 1:
                        best_latents, best_context = latents_in, context_new
                    context_in = context_new.detach().float().requires_grad_(True)
                    latents_in = latents_in.detach().float().requires_grad_(True)
                    out_log = convert_loss_log_to_numpy(controller.loss_log_dict)
                    if use_adaptive_optimization:
                        fn = adaptive_optimization_step_remover if is_remover else adaptive_optimization_step_editing
                        fn(controller, i, skip_optim_steps, out_log, num_ddim_steps=NUM_DDIM_STEPS, removal_loss_value_in=removal_loss_value_in)
                    out_log["loss"] = loss_val
                    global_loss_log_dict[i] = out_log
                    clear_controller_loss(controller)
                    controller.cur_step -= 1
            if optimize_latents:
                latents = best_latents.detach().clone()
                rescale_to_norm_(latents[-1], orig_norm)  # editor.py:316
            if best_context is not None and optimize_embeddings:
                context = best_context.detach()
                context_save = context
        elif i < fast_start_steps * n_t:
            pass    # editor.py:354-355: only the diffusion step is skipped; the reference-latent replacement and the fast-start warp below still run
        elif context_save is not None:
            context = context_save
        # the reference latent is overwritten below whenever the inversion trajectory is given, which makes its unconditional evaluation dead work
        skip = SKIP_DEAD_UNCOND_REFERENCE and ddim_latents is not None
        # the reference sample of this timestep was evaluated by the optimisation pass above (same latent, same conditional context): reuse it
        cached = do_opt and getattr(controller, "base_mode", None) == "write"
        if hasattr(controller, "base_mode"):
            controller.base_mode = "read" if cached else None
        if cached:
            set_attn_processor_for_edit(model, coords_base=(0, 1), coords_edit=(1, 2), use_cfg=True)     # batch [uncond edit, cond edit]
        elif skip:
            set_attn_processor_for_edit(model, coords_base=(1, 2), coords_edit=(2, 3), use_cfg=True)
        else:
            set_attn_processor_for_edit(model, coords_base=(2, 3), coords_edit=(3, 4), use_cfg=True)
        if not i < fast_start_steps * n_t:
            latents = diffusion_step(model, controller, latents, context, t, guidance_scale, transform_coords=transform_coordinates,
                                     skip_uncond_reference=skip, cached_reference=cached)
        if hasattr(controller, "base_mode"):
            controller.base_mode = None
        if ddim_latents is not None:
            i_n = len(ddim_latents) - 2 - i
            latents = torch.cat([ddim_latents[i_n].to(latents), latents[-1:].detach()], 0)  # editor.py:375-377
        if progress is not None:
            progress(i / NUM_DDIM_STEPS)
        if not is_remover and ((i < n_t * latent_replace and mask_obj is not None) or (i < n_t * fast_start_steps)):
            latents = _latent_warp_replace(controller, latents, transform_coordinates, fast=i < n_t * fast_start_steps)
    return latents, latent, global_loss_log_dict


def _latent_warp_replace(controller, latents, transform_coordinates, fast=False):
    """editor.py:382-399: paste the reference latent, splat-warped through the 64^2 correspondence field, inside the warped mask."""
    S = latents.shape[-1]
    dev = latents.device
    cache = controller._get_cache(S, transform_coordinates, dev)  # same coords@S / splat index the attention layers use
    m = geometry.resize_bilinear(controller.mask_new_warped[:1].float().contiguous(), S)[0, 0].contiguous()
    src = latents[-2:-1].detach().float().contiguous()
    warped = geometry.splat_composite(src, cache.idx, cache.dist2)
    out = latents.clone()
    base = latents[:1] if fast else latents[-1:]
    call("gd_latent_blend", ptr(base.contiguous()), ptr(warped), ptr(m), S * S, 1, warped.numel(), ptr(out[-1]), stream())
    return out


# hyper-parameters of large_scale_editor.perform_exp (:199-299) per edit type -- the canonical benchmark settings
EXP_PARAMS = {
    "geometry_editor": dict(cross_replace_steps={"default_": 0.95}, self_replace_steps=0.95, optimize_steps=0.65, lr=0.03,
                            latent_replace=0.1, optimize_embeddings=True, optimize_latents=True, obj_edit_step=0.9, skip_optim_steps=2,
                            guidance_scale=3.0,
                            loss_weights_dict={"self": {"sim": 55, "movement": 30.5, "removal": 2.6, "smoothness": 30.0, "amodal": 80.5},
                                               "cross": {"sim": 45, "movement": 30.34, "removal": 2.6, "smoothness": 15.0, "amodal": 3.5}}),
    "geometry_remover": dict(cross_replace_steps={"default_": 0.9}, self_replace_steps=0.9, optimize_steps=0.85, lr=0.03,
                             latent_replace=0.4, optimize_embeddings=True, optimize_latents=True, obj_edit_step=1.0, skip_optim_steps=2,
                             guidance_scale=5.0,
                             loss_weights_dict={"self": {"sim": 55, "removal": 4.6, "smoothness": 30.0},
                                                "cross": {"sim": 45, "removal": 4.6, "smoothness": 15.0}}),
}


def synthetic_embeddings(seed=SEED, device="cuda", image_size=IMAGE_SIZE):
    """context stand-ins: randn(2,77,768) text + randn uncond, CPU generator so every box sees the same values"""
    g = torch.Generator().manual_seed(seed + 1)
    text = torch.randn(1, 77, 768, generator=g).expand(2, 77, 768).contiguous()
    uncond = torch.randn(1, 77, 768, generator=g).expand(2, 77, 768).contiguous()
    x0 = torch.randn(1, 4, image_size // 8, image_size // 8, generator=g)
    return text.to(device), uncond.to(device), x0.to(device)


def stage_inputs(depth, image_mask, text_embeddings, uncond_embeddings, x0, device):
    """Host -> device copies of one edit request (the only H2D traffic of an edit).  Returns (staged dict, bytes copied)."""
    d_dev, m_dev = geometry.stage_depth_mask(depth, image_mask, device)
    obj = torch.from_numpy(np.ascontiguousarray(image_mask, dtype=np.float32)).to(device, non_blocking=True)
    text = text_embeddings.to(device, non_blocking=True).float()
    uncond = uncond_embeddings.to(device, non_blocking=True).float()
    x0 = x0.to(device, non_blocking=True).float()
    staged = dict(depth=d_dev, mask=m_dev, obj_mask=obj, text=text, uncond=uncond, x0=x0)
    nbytes = sum(t.numel() * t.element_size() for t in staged.values())
    return staged, nbytes


def make_controller(model, staged, transform_in, edit_type, hp, num_ddim_steps=50):
    """editor.py:508-638: correspondence field, amodal mesh mask and the controller of one edit.  Returns (controller, transform_coordinates)."""
    device = model.device
    g = geometry.correspondence_field_device(staged["depth"], staged["mask"], transform_in)
    mesh = geometry.mesh_mask(g["coords"], g["mask"])
    transform_coordinates = g["coords"][None]  # stays on the device (the reference round-trips through numpy, editor.py:546-547)
    cls = AttentionGeometryRemover if edit_type == "geometry_remover" else AttentionGeometryEdit
    controller = cls(["", ""], num_ddim_steps, cross_replace_steps=hp["cross_replace_steps"], self_replace_steps=hp["self_replace_steps"],
                     image_mask=None, empty_scale=0.0, use_all=False, obj_edit_step=hp["obj_edit_step"], device=device)
    # cache buffers and captured graphs of the gradient-free passes live on the model, per controller kind, and are reused by the next edit
    # (edits on one model are sequential: one live controller per model)
    controller._arena = model.__dict__.setdefault("_arenas", {}).setdefault(cls.__name__, {})
    controller._unet_graphs = model.__dict__.setdefault("_edit_graphs", {}).setdefault(cls.__name__, {})
    controller._grad_graphs = model.__dict__.setdefault("_grad_graph_store", {}).setdefault(cls.__name__, {})
    controller._grad_graphs_shared = True    # graphs.grad_pass: keyed on the edit's fingerprint (inpaint-row counts, mask sums), reused across edits
    controller.image_mask = staged["obj_mask"][None].tile(2, 1, 1)
    controller.amodal_mask = geometry.torch_erode(mesh[None, None])  # editor.py:633
    if hp.get("loss_weights_dict") is not None:
        import copy
        lw = copy.deepcopy(hp["loss_weights_dict"])
        controller.loss_weight_dict = lw
        controller.default_loss_weights = lw  # aliased exactly like editor.py:637-638
    return controller, transform_coordinates


def run_edit(model, staged, transform_in, edit_type="geometry_editor", num_ddim_steps=50, perform_ddim_inversion=True, seed=SEED,
             return_log=False, **overrides):
    """editor.py:428-711 on device-resident inputs: correspondence field -> DDIM inversion -> controller -> edit loop.
    Returns the final (2,4,64,64) latents [reference, edited] on the device."""
    global NUM_DDIM_STEPS
    NUM_DDIM_STEPS = num_ddim_steps
    device = model.device
    hp = dict(EXP_PARAMS[edit_type])
    hp.update(overrides)
    controller, transform_coordinates = make_controller(model, staged, transform_in, edit_type, hp, num_ddim_steps)
    text, uncond, x0 = staged["text"], staged["uncond"], staged["x0"]
    model.scheduler.set_timesteps(num_ddim_steps)
    # The per-resolution caches depend on the masks and the correspondence field only, and building them synchronises: do it now, before the
    # inversion passes are queued, so that the host can record (capture) this edit's optimisation-pass graph while the GPU is still busy
    # with the 50 inversion replays (graphs.grad_pass)
    if graphs.ENABLED and graphs.GRAD_ENABLED and x0.is_cuda:
        graphs.prebuild_caches(controller, transform_coordinates, device, int(x0.shape[-1]))
    if perform_ddim_inversion:
        ddim_latents = ddim_inversion_loop(model, x0, torch.cat([uncond[:1], text[:1]]), hp["guidance_scale"], num_ddim_steps)
    else:
        gen = torch.Generator().manual_seed(seed + 2)
        ddim_latents = [x0] + [torch.randn(*x0.shape, generator=gen).to(device) for _ in range(num_ddim_steps)]
    x_t = ddim_latents[-1]
    latents, _, log = text2image_ldm_stable(
        model, ["", ""], controller, num_inference_steps=num_ddim_steps, guidance_scale=hp["guidance_scale"], latent=x_t,
        uncond_embeddings=uncond, text_embeddings=text, transform_coordinates=transform_coordinates, mask_obj=staged["obj_mask"],
        optimize_steps=hp["optimize_steps"], latent_replace=hp["latent_replace"], lr=hp["lr"], optimize_embeddings=hp["optimize_embeddings"],
        optimize_latents=hp["optimize_latents"], ddim_latents=ddim_latents, edit_type=edit_type, skip_optim_steps=hp["skip_optim_steps"],
        removal_loss_value_in=hp.get("removal_loss_value_in", -1.5), fast_start_steps=hp.get("fast_start_steps", 0.0),
        num_first_optim_steps=hp.get("num_first_optim_steps", 5))
    model.unet.set_attn_processor(VanillaAttentionProcessor())  # editor.py:698
    if return_log:
        return latents, log
    return latents


def perform_geometric_edit(model, depth, image_mask, transform_in, text_embeddings, uncond_embeddings, x0, edit_type="geometry_editor",
                           **kw):
    """Public entry for one edit request with HOST inputs (numpy depth / mask, host tensors for the embeddings and the image latent):
    host->device staging, the edit, and the device->host copy of the result.  Returns (latents on the host, h2d bytes, d2h bytes)."""
    staged, h2d = stage_inputs(depth, image_mask, text_embeddings, uncond_embeddings, x0, model.device)
    latents = run_edit(model, staged, transform_in, edit_type, **kw)
    out = latents.cpu()
    return out, h2d, out.numel() * out.element_size()


def synthetic_request(kind="rotate3d", seed=SEED, pin=True, image_size=IMAGE_SIZE):
    """host-side inputs of one synthetic edit request; image_size 768 = BASELINE.json configs[3] (96^2 latent)"""
    image, depth, mask, T = synth.edit_inputs(kind, size=image_size)
    text, uncond, x0 = synthetic_embeddings(seed, "cpu", image_size)
    if pin and torch.cuda.is_available():
        text, uncond, x0 = text.pin_memory(), uncond.pin_memory(), x0.pin_memory()
    edit_type = "geometry_remover" if kind == "remove" else "geometry_editor"
    return dict(depth=depth, image_mask=mask, transform_in=T, text_embeddings=text, uncond_embeddings=uncond, x0=x0, edit_type=edit_type)


def perform_synthetic_edit(model, kind="rotate3d", num_ddim_steps=NUM_DDIM_STEPS, return_log=False, image_size=IMAGE_SIZE, **kw):
    """convenience wrapper used by the tests: device-side result of one synthetic edit"""
    req = synthetic_request(kind, pin=False, image_size=image_size)
    staged, _ = stage_inputs(req["depth"], req["image_mask"], req["text_embeddings"], req["uncond_embeddings"], req["x0"], model.device)
    return run_edit(model, staged, req["transform_in"], req["edit_type"], num_ddim_steps=num_ddim_steps, return_log=return_log, **kw)
