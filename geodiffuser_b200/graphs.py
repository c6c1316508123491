"""CUDA-graph replay of the gradient-free UNet evaluations of an edit (the 50 DDIM-inversion passes and the 50 classifier-free-guidance
passes; editor.py:339-366 and inversion.py:131-196 in the reference run them eagerly under torch.no_grad()).

One evaluation of the SD-1.5 body is ~1 700 kernel launches (32 of them the fused shared-attention layers); eagerly the host cannot issue
them as fast as a B200 retires them, so the edit is launch-bound.  Every pass of a given kind has identical shapes, so it is captured once
(inputs in static buffers, the timestep as a device scalar) and replayed.  A pass of the edit loop depends on controller state through
two flags only -- inside / outside the self-attention replace window, blending on / off (attention_processors.py:544, 619, 646) -- so the
captured graphs are keyed on those flags.  They hold pointers into the controller's per-resolution caches; `editor.make_controller` keeps
those caches in per-model buffers that the next edit refreshes in place (functional.ResolutionCache `arena`), so the graphs of one edit
serve every later edit of the same kind on that model.  A hand-made controller without an arena keeps its graphs to itself.
The kernels of the path launch on torch's current stream (`_lib.stream()`), which is the capture stream inside `torch.cuda.graph`.
"""
import torch

from . import _lib

ENABLED = True      # product default; tests compare against the eager path by switching it off


class GraphedUNet:
    """unet(sample, t, context) -> eps for fixed shapes, replayed from a CUDA graph.  `warmup` eager evaluations first (they also build
    every lazily-created cache: resolution caches, tensor maps, cudaFuncSetAttribute), then one capture."""

    def __init__(self, unet, sample, t, context, after_eval=None, warmup=1):
        dev = sample.device
        self.unet = unet
        self.sample = sample.detach().clone()
        self.context = context.detach().clone()
        self.t = torch.zeros(1, device=dev, dtype=torch.int64)
        self.t.fill_(int(t))
        self.graph = None
        self.out = None
        self.path_launches = 0          # kernels of the C-ABI library inside the captured graph (re-counted on every replay)
        self.after_eval = after_eval
        self.warmup_left = warmup

    def _eval(self):
        out = self.unet(self.sample, self.t, encoder_hidden_states=self.context)["sample"]
        if self.after_eval is not None:
            self.after_eval()           # host-side bookkeeping the eager evaluation did (step counters) must not run twice
        return out

    def __call__(self, sample, t, context):
        self.sample.copy_(sample)
        self.context.copy_(context)
        self.t.fill_(int(t))
        if self.graph is None:
            if self.warmup_left > 0:
                self.warmup_left -= 1
                return self.unet(self.sample, self.t, encoder_hidden_states=self.context)["sample"]
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            l0 = _lib.LAUNCHES
            with torch.cuda.graph(g):
                self.out = self._eval()
            self.path_launches = _lib.LAUNCHES - l0
            _lib.LAUNCHES = l0          # capturing launches nothing
            self.graph = g
        self.graph.replay()
        _lib.LAUNCHES += self.path_launches
        return self.out


def controller_key(controller):
    """the controller state a captured edit pass depends on"""
    c = controller
    return (bool(c.use_cfg), tuple(c.coords_base), tuple(c.coords_edit), c.num_self_replace[0] <= c.cur_step < c.num_self_replace[1],
            c.cur_step < int(c.num_steps * c.obj_edit_step))


def edit_pass(model, controller, latents_input, t, context):
    """One gradient-free UNet evaluation of the edit loop (the CFG pass).  Eager on the first occurrence of a controller state, captured on
    the second, replayed afterwards; the controller's step counter advances exactly as in the eager evaluation."""
    # a graph captured for an earlier edit may only be replayed once THIS controller has refreshed the shared cache buffers, i.e. after its
    # first real evaluation (normally the first optimisation pass)
    if not ENABLED or torch.is_grad_enabled() or getattr(controller, "eager_passes", 0) == 0:
        return model.unet(latents_input, t, encoder_hidden_states=context)["sample"]
    store = controller.__dict__.setdefault("_unet_graphs", {})
    arena = controller.__dict__.get("_arena")
    key = (controller_key(controller), tuple(latents_input.shape), id(model.unet), arena.get("generation", 0) if arena is not None else -1)
    g = store.get(key)
    if g is None:
        g = store[key] = GraphedUNet(model.unet, latents_input, t, context)
    step, layer = controller.cur_step, controller.cur_att_layer
    out = g(latents_input, t, context)
    # eager evaluation and capture both walk the 32 layers through AttentionControl.__call__, a replay does not: leave the counters where one
    # evaluation leaves them in every case
    controller.cur_step, controller.cur_att_layer = step + 1, layer
    return out


def inversion_pass(model, latents_input, t, context):
    """One UNet evaluation of the DDIM inversion (vanilla attention, no controller state): the graph lives on the model and is reused by
    every step of every edit."""
    if not ENABLED or torch.is_grad_enabled():
        return model.unet(latents_input, t, encoder_hidden_states=context)["sample"]
    store = model.__dict__.setdefault("_inversion_graphs", {})
    key = (tuple(latents_input.shape), tuple(context.shape), id(model.unet))
    g = store.get(key)
    if g is None:
        g = store[key] = GraphedUNet(model.unet, latents_input, t, context)
    return g(latents_input, t, context)
