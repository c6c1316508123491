"""wall-clock (synchronised) time per phase of one 50-step edit: inversion passes, optimisation passes (fwd+bwd+update), CFG passes, rest"""
import sys, time, collections
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import editor, graphs, unet_sd15, diffusion, optimization

kind = sys.argv[1] if len(sys.argv) > 1 else "rotate3d"
model = unet_sd15.build_model("cuda")
req = editor.synthetic_request(kind)
staged, _ = editor.stage_inputs(req["depth"], req["image_mask"], req["text_embeddings"], req["uncond_embeddings"], req["x0"], model.device)
acc = collections.defaultdict(float); cnt = collections.Counter()

def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc[name] += time.perf_counter() - t0; cnt[name] += 1
        return r
    return w

graphs.inversion_pass = timed("inversion_pass", graphs.inversion_pass)
graphs.edit_pass = timed("cfg_pass(unet)", graphs.edit_pass)
graphs.grad_pass = timed("opt_pass fwd+bwd", graphs.grad_pass)
editor.apply_latent_update = timed("opt_pass update", optimization.apply_latent_update)
editor.make_controller = timed("make_controller (geometry)", editor.make_controller)
editor.ddim_inversion_loop = timed("[ddim_inversion_loop whole]", editor.ddim_inversion_loop)
editor.text2image_ldm_stable = timed("[text2image whole]", editor.text2image_ldm_stable)
editor.convert_loss_log_to_numpy = timed("convert_loss_log", editor.convert_loss_log_to_numpy)
editor.set_attn_processor_for_edit = timed("set_attn_processor_for_edit", editor.set_attn_processor_for_edit)
editor.register_attention_control_diffusers = timed("register_attention_control", editor.register_attention_control_diffusers)
editor.clear_controller_loss = timed("clear_controller_loss", editor.clear_controller_loss)
editor._latent_warp_replace = timed("latent_warp_replace", editor._latent_warp_replace)
editor.norm_tensor = timed("norm_tensor", editor.norm_tensor)
editor.rescale_to_norm_ = timed("rescale_to_norm", editor.rescale_to_norm_)
diffusion.DDIMScheduler._launch = timed("scheduler step kernel", diffusion.DDIMScheduler._launch)
for it in range(3):
    acc.clear(); cnt.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    editor.run_edit(model, staged, req["transform_in"], req["edit_type"])
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print(f"--- edit {it}: total {tot*1e3:.0f} ms")
    for k, v in acc.items():
        print(f"  {k:28s} {v*1e3:8.1f} ms  n={cnt[k]:3d}  avg {v/cnt[k]*1e3:7.2f} ms")
    
