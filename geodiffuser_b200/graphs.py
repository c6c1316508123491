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
import contextlib
import gc
import threading

import torch

from . import _lib

_CAPTURE_LOCK = threading.RLock()   # one capture at a time per process: the garbage collector is paused around it (runner.EditWorkers runs
                                    # several edits of one GPU from several threads).  EAGER evaluations of the body take it as well: that is
                                    # where cuDNN autotunes (torch.backends.cudnn.benchmark), and its timing runs (device synchronisation,
                                    # allocations) abort a capture in progress on another thread -- CUDNN_STATUS_INTERNAL_ERROR /
                                    # cudaErrorStreamCaptureInvalidated.  Replays never take it.
ENABLED = True      # product default; tests compare against the eager path by switching it off
GRAD_ENABLED = True  # the optimisation pass (forward + backward) as a graph as well
SHARE_GRAD_GRAPHS = True             # reuse an edit's recorded optimisation pass for later edits with the same fingerprint (False: one capture per edit)
MAX_SHARED_GRAD_GRAPHS = 6           # optimisation-pass graphs kept per (model, controller kind) for reuse by later edits with the same fingerprint
CAPTURE_ERROR_MODE = "thread_local"   # cudaStreamCaptureMode of the hand-driven captures: CUDA calls that OTHER threads make meanwhile (another
                                      # edit lane's allocations, event queries) stay legal


@contextlib.contextmanager
def _capture(graph, pool, stream, sync=True):
    """Stream capture into `graph` on `stream`.  torch.cuda.graph() would also run gc.collect() and torch.cuda.empty_cache() first; an edit
    captures one graph per request, and an emptied allocator cache makes every eager allocation after it pay for cudaMalloc again, so the capture
    is driven by hand.  The cyclic garbage collector is paused: a collection in the middle of a capture may destroy an older CUDAGraph or free
    its tensors, which invalidates the capture in progress."""
    with _CAPTURE_LOCK:
        was = gc.isenabled()
        gc.disable()
        cur = torch.cuda.current_stream(stream.device)
        if sync:
            cur.synchronize()       # (the current stream only: another edit lane of this GPU may be busy on its own streams)
        stream.wait_stream(cur)
        try:
            with torch.cuda.stream(stream):
                graph.capture_begin(pool=pool, capture_error_mode=CAPTURE_ERROR_MODE)
                try:
                    yield
                finally:
                    graph.capture_end()
        finally:
            cur.wait_stream(stream)
            if was:
                gc.enable()


def _pool(model):
    """one memory pool for every graph captured on a model: a dead graph's memory goes back to the pool instead of through cudaFree, so
    re-capturing for the next edit does not pay for allocation"""
    return model.__dict__.setdefault("_graph_pool", torch.cuda.graph_pool_handle())


def _side_stream(model, dev):
    """persistent warm-up stream (its cached allocations are reused by the next edit's warm-up pass)"""
    st = model.__dict__.get("_graph_side_stream")
    if st is None:
        st = model.__dict__["_graph_side_stream"] = torch.cuda.Stream(device=dev)
    return st


class GraphedUNet:
    """unet(sample, t, context) -> eps for fixed shapes, replayed from a CUDA graph.  `warmup` eager evaluations first (they also build
    every lazily-created cache: resolution caches, tensor maps, cudaFuncSetAttribute), then one capture."""

    def __init__(self, unet, sample, t, context, after_eval=None, warmup=1, pool=None, stream=None):
        dev = sample.device
        self.unet = unet
        self.sample = sample.detach().clone()
        self.context = context.detach().clone()
        self.t = torch.zeros(1, device=dev, dtype=torch.int64)
        self.t.fill_(int(t))
        self.graph = None
        self.out = None
        self.path_launches = 0          # kernels of the C-ABI library inside the captured graph (re-counted on every replay)
        self.path_flops = 0.0           # their algorithmic attention-path FLOP (_lib.FLOPS)
        self.after_eval = after_eval
        self.warmup_left = warmup
        self.warm_threads = set()       # host threads that have evaluated this pass eagerly (cuDNN autotune cache / cuBLAS handle are per thread)
        self.pool = pool
        self.stream = stream if stream is not None else torch.cuda.Stream(device=dev)

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
            tid = threading.get_ident()
            if self.warmup_left > 0 or tid not in self.warm_threads:
                self.warmup_left = max(0, self.warmup_left - 1)
                self.warm_threads.add(tid)
                with _CAPTURE_LOCK:
                    return self.unet(self.sample, self.t, encoder_hidden_states=self.context)["sample"]
            g = torch.cuda.CUDAGraph()
            cnt = _lib.counters()                       # (this thread's: a capture is recorded by one thread)
            l0, f0 = cnt.launches, cnt.flops
            with _capture(g, self.pool, self.stream):
                self.out = self._eval()
            self.path_launches, self.path_flops = cnt.launches - l0, cnt.flops - f0
            cnt.launches, cnt.flops = l0, f0            # capturing launches nothing
            self.graph = g
        self.graph.replay()
        cnt = _lib.counters()
        cnt.launches += self.path_launches
        cnt.flops += self.path_flops
        return self.out


def controller_key(controller):
    """the controller state a captured edit pass depends on"""
    c = controller
    return (bool(c.use_cfg), tuple(c.coords_base), tuple(c.coords_edit), c.num_self_replace[0] <= c.cur_step < c.num_self_replace[1],
            c.cur_step < int(c.num_steps * c.obj_edit_step), getattr(c, "base_mode", None))


def edit_pass(model, controller, latents_input, t, context):
    """One gradient-free UNet evaluation of the edit loop (the CFG pass).  Eager on the first occurrence of a controller state, captured on
    the second, replayed afterwards; the controller's step counter advances exactly as in the eager evaluation."""
    # a graph captured for an earlier edit may only be replayed once THIS controller has refreshed the shared cache buffers, i.e. after its
    # first real evaluation (normally the first optimisation pass)
    if not ENABLED or torch.is_grad_enabled() or getattr(controller, "eager_passes", 0) == 0 or getattr(controller, "store_attention_maps", False):
        with _CAPTURE_LOCK:
            return model.unet(latents_input, t, encoder_hidden_states=context)["sample"]
    store = controller.__dict__.setdefault("_unet_graphs", {})
    arena = controller.__dict__.get("_arena")
    key = (controller_key(controller), tuple(latents_input.shape), id(model.unet), arena.get("generation", 0) if arena is not None else -1)
    g = store.get(key)
    if g is None:
        g = store[key] = GraphedUNet(model.unet, latents_input, t, context, pool=_pool(model), stream=_side_stream(model, latents_input.device))
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
        with _CAPTURE_LOCK:
            return model.unet(latents_input, t, encoder_hidden_states=context)["sample"]
    store = model.__dict__.setdefault("_inversion_graphs", {})
    key = (tuple(latents_input.shape), tuple(context.shape), id(model.unet))
    g = store.get(key)
    if g is None:
        g = store[key] = GraphedUNet(model.unet, latents_input, t, context, pool=_pool(model), stream=_side_stream(model, latents_input.device))
    return g(latents_input, t, context)


class GraphedGradPass:
    """One optimisation pass -- UNet forward under autograd with the loss-bearing attention layers, then d loss / d (latents, context)
    (editor.py:239-253 + optimization.py:201 in the reference) -- captured as ONE graph: forward, the fused layers' backward kernels and
    torch's backward of the body.  Static inputs: latents, context, timestep; static outputs: loss, the two gradients; the controller's
    log accumulator is a static buffer already.  Everything that changes between passes lives in device memory (the removal weight:
    controller.sync_device_weights) or in the key (all other weights, the controller flags)."""

    def __init__(self, model, latents, context, t, sync=True):
        self.sync = sync            # False: record the graph without draining the device first (the GPU may still be running earlier passes)
        self.model = model          # (no reference to the controller: the graph is stored on it and must die with it, by refcount)
        self.latents = latents.detach().clone().requires_grad_(True)
        self.context = context.detach().clone().requires_grad_(True)
        self.t = torch.zeros(1, device=latents.device, dtype=torch.int64)
        self.graph = None
        self.loss = self.g_lat = self.g_ctx = None
        self.num_layers = 0
        self.path_launches = 0
        self.path_flops = 0.0

    def _run(self, c):
        with torch.enable_grad():
            self.model.unet(self.latents, self.t, encoder_hidden_states=self.context[self.context.shape[0] // 2:])
            g = torch.autograd.grad(c.loss, [self.latents, self.context], allow_unused=True)
        return c.loss.detach(), g[0], g[1]

    def __call__(self, c, latents, context, t):
        self.latents.data.copy_(latents.detach())
        self.context.data.copy_(context.detach())
        self.t.fill_(int(t))
        if self.graph is None:
            g = torch.cuda.CUDAGraph()
            step, layers = c.cur_step, c.loss_log_dict["num_layers"]
            cnt = _lib.counters()
            l0, f0 = cnt.launches, cnt.flops
            with _capture(g, _pool(self.model), _side_stream(self.model, latents.device), sync=self.sync):
                self.loss, self.g_lat, self.g_ctx = self._run(c)
            # (the layers' backward launches run on the autograd thread and are counted there: add that thread's share of this capture)
            self.path_launches, cnt.launches = cnt.launches - l0, l0
            self.path_flops, cnt.flops = cnt.flops - f0, f0
            self.num_layers = c.loss_log_dict["num_layers"] - layers
            c.cur_step, c.loss_log_dict["num_layers"] = step, layers
            self.graph = g
        self.graph.replay()
        cnt = _lib.counters()
        cnt.launches += self.path_launches
        cnt.flops += self.path_flops
        c.loss = self.loss
        c.cur_step += 1
        c.loss_log_dict["num_layers"] += self.num_layers
        return self.g_lat, self.g_ctx


def prebuild_caches(controller, transform_coords, device, latent_size):
    """Builds the controller's per-resolution caches (masks, splat index, inpaint rows, amodal tables: functional.ResolutionCache) for every
    UNet level, as the first eager evaluation of an edit would (a no-op for the levels that exist already)."""
    if not hasattr(controller, "_get_cache"):
        return False
    controller._ensure_device_state(device)
    for f in (1, 2, 4, 8):
        if latent_size % f == 0:
            controller._get_cache(latent_size // f, transform_coords, device)
    controller.eager_passes = max(1, getattr(controller, "eager_passes", 0))   # the shared cache buffers now belong to this edit
    return True


def _prebuild_caches(model, controller, latents):
    """prebuild_caches with the correspondence field the registered processors hold; False if they are not this package's processors"""
    for proc in model.unet.attn_processors.values():
        if getattr(proc, "controller", None) is controller and getattr(proc, "transform_coords", None) is not None:
            return prebuild_caches(controller, proc.transform_coords, latents.device, int(latents.shape[-1]))
    return False


def grad_pass(model, controller, latents, context, t):
    """-> (d loss / d latents, d loss / d context) of one optimisation pass; controller.loss and controller.loss_log_dict hold the loss and its
    logged terms afterwards, as after the reference's diffusion_step(use_cfg=False).  `context` is the full [uncond, text] stack (the UNet sees
    its text half, editor.py:250).  First occurrence of a controller state: eager, on a side stream (which doubles as the warm-up a backward
    capture needs); second: capture; then replay."""
    dev = latents.device

    def eager():
        with _CAPTURE_LOCK, torch.enable_grad():
            model.unet(latents, t, encoder_hidden_states=context[context.shape[0] // 2:])
            g = torch.autograd.grad(controller.loss, [latents, context], allow_unused=True)
            if len(threading.enumerate()) > 1:
                torch.cuda.current_stream(dev).synchronize()     # the backward's autotune runs on the autograd thread: let it finish under the lock
        return g

    if not (ENABLED and GRAD_ENABLED) or not latents.is_cuda:
        return eager()
    controller.sync_device_weights(dev)
    store = controller.__dict__.setdefault("_grad_graphs", {})
    lw = controller.loss_weight_dict
    weights = tuple(sorted((a, k, float(v)) for a in ("self", "cross") for k, v in lw[a].items() if k != "removal"))
    # What a recorded pass bakes in besides pointers: per resolution the inpaint-row count M (buffer shapes) and the mask sums (the loss
    # normalisers are kernel arguments).  Everything else it reads -- masks, row lists, splat index, amodal tables, base stores, log accumulator,
    # removal weight -- sits in the per-model arena at fixed addresses and is refreshed in place by the next edit, so with an arena the graph
    # of one edit serves every later edit with the same fingerprint (editor.make_controller keeps the store on the model); without one the
    # store is the controller's own and dies with it.
    arena = controller.__dict__.get("_arena")
    shared = SHARE_GRAD_GRAPHS and arena is not None and controller.__dict__.get("_grad_graphs_shared", False)
    if shared:
        _prebuild_caches(model, controller, latents)
    finger = tuple(sorted((S, c.M, c.sum_bg, c.sum_edit, c.sum_inp, c.sum_w_am) for S, c in getattr(controller, "_res_cache", {}).items())) if shared else id(controller)
    key = (controller_key(controller), weights, tuple(latents.shape), tuple(context.shape), id(model.unet), finger,
           arena.get("generation", 0) if arena is not None else -1)
    g = store.get(key)
    if g is not None and g != "warm" and shared:
        store[key] = store.pop(key)             # most recently used last
    while shared and len(store) > MAX_SHARED_GRAD_GRAPHS:
        store.pop(next(iter(store)))            # each holds the activations of one forward + backward in the graph pool
    # what a backward capture needs warmed up (cuBLAS / cuDNN handles and workspaces, cudnn.benchmark choices, the side stream's allocator
    # cache) belongs to the process and the model, not to the edit: once one eager pass of this kind and shape has run on this model, the next
    # edit captures its first optimisation pass directly -- after building its per-resolution caches, which synchronise and so cannot be
    # built inside a capture
    warmed = model.__dict__.setdefault("_grad_warm", set())
    # (cuDNN's benchmark cache is keyed on these global switches as well: a capture after one of them changed would have to autotune inside
    # the capture, which cuDNN cannot do)
    cd = torch.backends.cudnn
    # (... and both cuDNN's autotune cache and the cuBLAS handle are per host thread: runner.EditWorkers drives a model from its lane's thread)
    wkey = (type(controller).__name__, tuple(latents.shape), tuple(context.shape), id(model.unet), cd.benchmark, cd.allow_tf32, cd.deterministic,
            torch.backends.cuda.matmul.allow_tf32, threading.get_ident())
    if g is None and wkey in warmed and _prebuild_caches(model, controller, latents):
        # recorded without a device synchronisation: when the caches were built before the inversion (editor.run_edit) nothing here waits for
        # the GPU, so the ~45 ms of host work of the capture hide under the inversion replays still in flight
        g = store[key] = GraphedGradPass(model, latents, context, t, sync=False)
        return g(controller, latents, context, t)
    if g is None:
        warmed.add(wkey)
        store[key] = "warm"
        side = _side_stream(model, dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            out = eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        return out
    if g == "warm":
        g = store[key] = GraphedGradPass(model, latents, context, t)
    return g(controller, latents, context, t)
