"""Latent / context gradient step and the adaptive loss-weight schedule of the optimisation loop, mirroring
/root/reference/GeoDiffuser/utils/optimization.py (`_update_latent` :165-253 optimizer=None branch, `adaptive_optimization_step_editing`
:7-55, `_remover` :58-105) and generic_torch.norm_tensor (:87).  The masked update and the norm-preserving rescale are CUDA launches."""
import torch

from . import geometry
from ._lib import call, ptr, stream


def norm_tensor(A, eps=1e-12):
    """generic_torch.py:87 (returns a python float; one 4-byte D2H like the reference's .item())"""
    x = A.detach().float().contiguous().clone()
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    call("gd_norm_rescale", ptr(x), x.numel(), 0.0, None, ptr(out), stream())
    return float(out)


def norm_tensor_dev(A):
    """the same norm as a 1-element device tensor: the edit loop only hands it back to rescale_to_norm_, so it never needs to reach the host
    (the reference's .item() here is a device -> host sync in front of every optimisation step)"""
    x = A.detach().float().contiguous().clone()
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    call("gd_norm_rescale", ptr(x), x.numel(), 0.0, None, ptr(out), stream())
    return out


def rescale_to_norm_(x, target_norm):
    """in place: x *= target_norm / ||x||   (editor.py:316)"""
    assert x.is_contiguous() and x.dtype == torch.float32
    if torch.is_tensor(target_norm):
        call("gd_norm_rescale", ptr(x), x.numel(), 0.0, ptr(target_norm.float().contiguous()), None, stream())
    else:
        call("gd_norm_rescale", ptr(x), x.numel(), float(target_norm), None, None, stream())
    return x


def apply_latent_update(latents, grad_cond, step_size, mask=None, context=None, context_grad=None):
    """optimization.py:213-253 given the gradients: only the LAST batch entry moves.  mask: mask_new_warped[:1] at image size."""
    latents = latents.detach().float().contiguous()
    g = grad_cond.detach().float().contiguous()
    out = latents.clone()
    n = latents[-1].numel()
    m = None
    hw = latents.shape[-1] * latents.shape[-2]
    if mask is not None:
        m512 = mask.reshape(mask.shape[-2:]).float().contiguous()
        m = geometry.resize_bilinear(m512[None], latents.shape[-1])[0].contiguous()
    call("gd_latent_update", ptr(latents[-1]), ptr(g[-1]), ptr(m), hw, float(step_size), n, ptr(out[-1]), stream())
    context_new = None
    if context is not None:
        context = context.detach().float().contiguous()
        gc = context_grad.detach().float().contiguous()
        context_new = context.clone()
        call("gd_latent_update", ptr(context[-1]), ptr(gc[-1]), None, 0, float(step_size), context[-1].numel(), ptr(context_new[-1]), stream())
    return out, context_new


def _update_latent(latents, loss, step_size, mask=None, context=None, scaler=None, optimizer=None):
    """optimization.py:165-253.  The torch.optim branch is dead in the reference (use_optimizer is never forwarded)."""
    if optimizer is not None or scaler is not None:
        raise NotImplementedError("only the optimizer=None branch is reachable from the reference drivers (editor.py:233-234)")
    # The fused layer returns no gradient (None) where the reference's graph carries a structural zero (base sample, detached K/V of
    # the remover's cross layers), so an input the loss does not reach comes back undefined here instead of as zeros.
    grads = torch.autograd.grad(loss, [latents, context], retain_graph=False, allow_unused=True)
    g_lat = grads[0] if grads[0] is not None else torch.zeros_like(latents)
    g_ctx = grads[1] if grads[1] is not None else torch.zeros_like(context)
    return apply_latent_update(latents, g_lat, step_size, mask, context, g_ctx)


def _adaptive(controller, i, skip_optim_steps, out_loss_log_dict, num_ddim_steps, removal_loss_value_in, reduce_div):
    frac = i / num_ddim_steps
    if frac < 0.4:
        remaining_steps = int((0.4 - frac) * num_ddim_steps / skip_optim_steps)
        expected = removal_loss_value_in / (1.25) ** (remaining_steps)
        cur = out_loss_log_dict["self"]["removal"]
        if expected < cur:
            controller.loss_weight_dict["self"]["removal"] *= 1.3
        elif 2.5 * expected > cur:
            controller.loss_weight_dict["self"]["removal"] /= reduce_div
    elif (frac > 0.4) and (frac < 0.8):
        if (removal_loss_value_in - 0.3) < out_loss_log_dict["self"]["removal"]:
            controller.loss_weight_dict["self"]["removal"] *= 2.0
        else:
            controller.initialize_default_loss_weights()
    else:
        controller.initialize_default_loss_weights()


def adaptive_optimization_step_editing(controller, i, skip_optim_steps, out_loss_log_dict, num_ddim_steps, removal_loss_value_in=-1.5):
    _adaptive(controller, i, skip_optim_steps, out_loss_log_dict, num_ddim_steps, removal_loss_value_in, 2.0)


def adaptive_optimization_step_remover(controller, i, skip_optim_steps, out_loss_log_dict, num_ddim_steps, removal_loss_value_in=-1.5):
    _adaptive(controller, i, skip_optim_steps, out_loss_log_dict, num_ddim_steps, removal_loss_value_in, 2.5)
