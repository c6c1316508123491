"""Drop-in mirror of the reference's attention-controller / processor hook API for the geometry-warped shared-attention path.

Same names, argument meaning, attributes and error behaviour as /root/reference/GeoDiffuser/utils/attention_processors.py and
attention_sharing.py (line numbers below refer to those files); the arithmetic is a single fused CUDA call per layer
(`functional.shared_attention_layer`) instead of ~60 torch / pytorch3d ops.

  register_attention_control_diffusers   attention_processors.py:26-53
  set_attn_processor_for_edit            attention_processors.py:56-67
  VanillaAttentionProcessor              attention_processors.py:69-139
  EditProcessor                          attention_processors.py:141-228
  AttentionControl / AttentionStore      attention_sharing.py:110-207
  AttentionGeometryEdit                  attention_processors.py:377-736
  AttentionGeometryRemover               attention_processors.py:741-1022
"""
import abc
import math

import numpy as np
import torch

from . import functional as Fn
from . import geometry

# the random-init UNet used here has plain nn.Linear projections: same call convention as diffusers with the PEFT backend
USE_PEFT_BACKEND = True

# The processors hand q / k / v to the controller as functional.ProjView (the projection output (B, N, H*d) presented with the reference's
# (B*H, N, d) shape) and get the attention output back in projection layout: the kernels address the head slabs in place, so the
# reference's head_to_batch_dim / batch_to_head_dim copies (attention_processors.py:201-203, 225) disappear.  False = the reference's
# literal tensor layout through the same kernels (tests run both).
PROJECTION_LAYOUT = True


def register_attention_control_diffusers(model, controller, transform_coords=None):
    attn_procs = {}
    count = 0
    for name in model.unet.attn_processors.keys():
        if name.startswith("mid_block"):
            place_in_unet = "mid"
        elif name.startswith("up_blocks"):
            place_in_unet = "up"
        elif name.startswith("down_blocks"):
            place_in_unet = "down"
        else:
            continue
        count += 1
        attn_procs[name] = EditProcessor(transform_coords, controller, place_in_unet)
    model.unet.set_attn_processor(attn_procs)
    controller.num_att_layers = count


def set_attn_processor_for_edit(model, perform_edit=True, coords_base=(2, 3), coords_edit=(3, 4), use_cfg=True):
    for proc in model.unet.attn_processors.values():
        proc.perform_edit = perform_edit
        proc.controller.coords_base = coords_base
        proc.controller.coords_edit = coords_edit
        proc.controller.use_cfg = use_cfg


def _project_qkv(attn, hidden_states, encoder_hidden_states, attention_mask, temb, scale):
    """The part of the diffusers processor protocol that precedes the attention product (attention_processors.py:164-203)."""
    args = () if USE_PEFT_BACKEND else (scale,)
    if attn.spatial_norm is not None:
        hidden_states = attn.spatial_norm(hidden_states, temb)
    input_ndim = hidden_states.ndim
    shape4 = None
    if input_ndim == 4:
        shape4 = hidden_states.shape
        batch_size, channel, height, width = shape4
        hidden_states = hidden_states.view(batch_size, channel, height * width).transpose(1, 2)
    batch_size, sequence_length, _ = hidden_states.shape if encoder_hidden_states is None else encoder_hidden_states.shape
    attention_mask = attn.prepare_attention_mask(attention_mask, sequence_length, batch_size)
    if attn.group_norm is not None:
        hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
    query = attn.to_q(hidden_states, *args)
    is_cross = True
    if encoder_hidden_states is None:
        encoder_hidden_states = hidden_states
        is_cross = False
    elif attn.norm_cross:
        encoder_hidden_states = attn.norm_encoder_hidden_states(encoder_hidden_states)
    key = attn.to_k(encoder_hidden_states, *args)
    value = attn.to_v(encoder_hidden_states, *args)
    if PROJECTION_LAYOUT and query.is_cuda:
        h = attn.heads
        return Fn.ProjView(query, h), Fn.ProjView(key, h), Fn.ProjView(value, h), is_cross, shape4, args
    return attn.head_to_batch_dim(query), attn.head_to_batch_dim(key), attn.head_to_batch_dim(value), is_cross, shape4, args


def _finish(attn, hidden_states, residual, shape4, args, proj=False):
    if not proj:
        hidden_states = attn.batch_to_head_dim(hidden_states)
    hidden_states = attn.to_out[0](hidden_states, *args)
    hidden_states = attn.to_out[1](hidden_states)
    if shape4 is not None:
        hidden_states = hidden_states.transpose(-1, -2).reshape(shape4)
    if attn.residual_connection:
        hidden_states = hidden_states + residual
    if attn.rescale_output_factor != 1.0:      # x / 1.0 == x: not worth a kernel launch per layer
        hidden_states = hidden_states / attn.rescale_output_factor
    return hidden_states


class VanillaAttentionProcessor:
    """Plain softmax(QK^T)V through the fused kernel (reference :69-139 materialises the score matrix)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale: float = 1.0):
        residual = hidden_states
        q, k, v, _, shape4, args = _project_qkv(attn, hidden_states, encoder_hidden_states, None, temb, scale)
        if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
            raise NotImplementedError("VanillaAttentionProcessor is forward-only (DDIM inversion / final reset, editor.py:698)")
        hidden_states = Fn.plain_attention(q, k, v, attn.scale, attn.heads)
        return _finish(attn, hidden_states, residual, shape4, args, isinstance(q, Fn.ProjView))


class EditProcessor:
    def __init__(self, transform_coords, controller, place_in_unet="down", perform_edit=True, coords_base=(2, 3), coords_edit=(3, 4),
                 use_cfg=True):
        self.transform_coords = transform_coords
        self.place_in_unet = place_in_unet
        self.perform_edit = perform_edit
        self.controller = controller
        self.controller.use_cfg = use_cfg
        self.controller.coords_base = coords_base
        self.controller.coords_edit = coords_edit

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale: float = 1.0):
        residual = hidden_states
        q, k, v, is_cross, shape4, args = _project_qkv(attn, hidden_states, encoder_hidden_states, attention_mask, temb, scale)
        if self.perform_edit:
            hidden_states = self.controller(q, k, v, is_cross=is_cross, place_in_unet=self.place_in_unet,
                                            transform_coords=self.transform_coords, scale=attn.scale, mask=None)
        else:
            if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
                raise NotImplementedError("perform_edit=False is forward-only")
            hidden_states = Fn.plain_attention(q, k, v, attn.scale, attn.heads)
        return _finish(attn, hidden_states, residual, shape4, args, isinstance(q, Fn.ProjView))


# ------------------------------------------------------------------------------------------------------------------
class AttentionControl(abc.ABC):
    """attention_sharing.py:110-155"""

    def step_callback(self, x_t, transform_coords):
        return x_t

    def between_steps(self):
        return

    @property
    def num_uncond_att_layers(self):
        return 0  # LOW_RESOURCE = False (attention_sharing.py:11)

    @abc.abstractmethod
    def forward(self, q, k, v, is_cross: bool, place_in_unet: str, transform_coords=None, scale=None, mask=None):
        raise NotImplementedError

    def __call__(self, q, k, v, is_cross: bool, place_in_unet: str, transform_coords=None, scale=None, mask=None):
        if self.cur_att_layer >= self.num_uncond_att_layers:
            out = self.forward(q, k, v, is_cross, place_in_unet, transform_coords=transform_coords, scale=scale, mask=mask)
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers + self.num_uncond_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
            self.between_steps()
            if not (q.is_cuda and torch.cuda.is_current_stream_capturing()):
                self.eager_passes = getattr(self, "eager_passes", 0) + 1   # graphs.py: caches are complete after one real evaluation
        return out

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    def __init__(self):
        self.cur_step = 0
        self.num_att_layers = -1
        self.cur_att_layer = 0


class AttentionStore(AttentionControl):
    """attention_sharing.py:158-207.  Attention maps are never materialised by the fused kernel; what the reference stores (maps of
    N <= 16^2 query tokens, :166-179; the geometry controllers' `store_attention_maps`, attention_processors.py:452-454, 562-564) is served by an
    explicit probability kernel (functional.attention_maps) when requested."""

    @staticmethod
    def get_empty_store():
        return {"down_cross": [], "mid_cross": [], "up_cross": [], "down_self": [], "mid_self": [], "up_self": []}

    def forward(self, q, k, v, is_cross: bool, place_in_unet: str, transform_coords=None, scale=None, mask=None):
        heads = q.shape[0] // max(1, getattr(self, "batch_size", 1))
        if q.shape[1] <= 16 ** 2:      # :166-169: the maps of the small levels are kept
            self.attn_store(Fn.attention_maps(q, k, scale, heads), is_cross, place_in_unet)
        return Fn.plain_attention(q, k, v, scale, heads)

    def attn_store(self, attn, is_cross: bool, place_in_unet: str):
        key = f"{place_in_unet}_{'cross' if is_cross else 'self'}"
        if attn.shape[1] <= 16 ** 2:
            self.step_store[key].append(attn.detach())

    def between_steps(self):
        if len(self.attention_store) == 0:
            self.attention_store = self.step_store
        else:
            for key in self.step_store:
                self.attention_store[key] = self.attention_store[key] + self.step_store[key]
                if self.cur_step == 1:
                    self.attention_store["length_" + key] = len(self.step_store[key])
        self.step_store = self.get_empty_store()

    def get_average_attention(self):
        return {key: [item / self.cur_step for item in self.attention_store[key]] for key in self.attention_store}

    def reset(self):
        super().reset()
        self.step_store = self.get_empty_store()
        self.attention_store = {}

    def __init__(self):
        super().__init__()
        self.step_store = self.get_empty_store()
        self.attention_store = {}


class _GeometryControllerBase(AttentionStore, abc.ABC):
    KIND = "edit"
    LOG_KEYS = ("sim", "movement", "removal", "smoothness")
    _TERM_SLOT = {"sim": 0, "movement": 1, "removal": 2, "smoothness": 3, "amodal": 4}

    def step_callback(self, x_t, transform_coords=None):
        if self.local_blend is not None:
            x_t = self.local_blend(x_t, self.attention_store, transform_coords)
        return x_t

    # -- loss bookkeeping (same dict shapes as the reference; values are views of one device buffer so that the per-layer
    #    accumulation happens inside the loss-reduce kernel instead of 8 tiny torch adds per layer)
    def initialize_loss_log_dict(self):
        if self._log_accum is not None:
            self._log_accum.zero_()
            self.loss_log_dict = {
                "self": {k: self._log_accum[0, self._TERM_SLOT[k]] for k in self.LOG_KEYS},
                "cross": {k: self._log_accum[1, self._TERM_SLOT[k]] for k in self.LOG_KEYS},
                "num_layers": 0}
        else:
            self.loss_log_dict = {"self": {k: 0.0 for k in self.LOG_KEYS}, "cross": {k: 0.0 for k in self.LOG_KEYS}, "num_layers": 0}

    def initialize_default_loss_weights(self):
        self.loss_weight_dict = self.default_loss_weights  # aliasing is the reference's behaviour

    @staticmethod
    def _on(t, device):
        """is tensor `t` on `device`?  (`device` may come without an index -- torch.device("cuda") -- which never compares equal to a tensor's)"""
        device = torch.device(device)
        return t is not None and t.device.type == device.type and (device.index is None or t.device.index == device.index)

    def _ensure_device_state(self, device):
        if not self._on(self._log_accum, device):
            # with a per-model arena the accumulator (and the removal weight below) sit at a fixed address, like the caches: a recorded
            # optimisation pass of an earlier edit writes into the buffers the current controller reads
            if self._arena is not None:
                buf = self._arena.get("log_accum")
                if not self._on(buf, device):
                    buf = self._arena["log_accum"] = torch.zeros(2, 6, device=device, dtype=torch.float32)
                self._log_accum = buf
            else:
                self._log_accum = torch.zeros(2, 6, device=device, dtype=torch.float32)
            n = self.loss_log_dict["num_layers"] if self.loss_log_dict else 0
            self.initialize_loss_log_dict()
            self.loss_log_dict["num_layers"] = n

    def _common_init(self, prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend, controller, image_mask, empty_scale,
                     use_all, obj_edit_step, mode):
        self.mode = mode
        self.prev_controller = controller
        self.last_cross_mask = None
        self.thre = 0.00001
        self.empty_scale = empty_scale
        self.use_all = use_all
        self.loss = 0.0
        self.batch_size = len(prompts)
        # cross_replace_alpha (ptp_utils.get_time_words_attention_alpha) is read (:654) but never used by the reference path
        self.cross_replace_alpha = None
        self.cross_replace_steps = cross_replace_steps
        if type(self_replace_steps) is float:
            self_replace_steps = 0, self_replace_steps
        self.num_self_replace = int(num_steps * self_replace_steps[0]), int(num_steps * self_replace_steps[1])
        self.local_blend = local_blend
        self.mask_inpaint = None
        self.obj_edit_step = obj_edit_step
        self.num_steps = num_steps
        self.mask_new_warped = None
        self.mask_wo_edit = None
        self.mask_1_empty = None
        self.amodal_mask = None
        self.coords_base = (2, 3)
        self.coords_edit = (3, 4)
        self.use_cfg = True
        self.store_attention_maps = False
        self.masks_cache_dict = {}
        self._res_cache = {}
        self._w_rem_dev = None      # device copy of loss_weight_dict[*]["removal"] (self, cross), see sync_device_weights
        self._w_rem_host = None
        self._arena = None          # editor.make_controller: per-model buffers that keep cache addresses stable across edits
        self.base_mode = None       # None | "write" (optimisation pass stores the base sample's K / V / warped output per layer)
        self._base_stores = {}      #                  | "read" (the CFG pass of the same timestep runs without the base sample); editor drives it
        self._log_accum = None
        self.loss_log_dict = None
        self.loss_weight_dict = None

    def sync_device_weights(self, device):
        """Mirrors the removal-loss weights (the only ones the adaptive schedule changes, optimization.py:7-105) into device memory, so that a
        CUDA graph captured for one optimisation pass stays valid after the schedule has moved them.  Two 4-byte fills; no host sync."""
        w = (float(self.loss_weight_dict["self"].get("removal", 0.0)), float(self.loss_weight_dict["cross"].get("removal", 0.0)))
        if not self._on(self._w_rem_dev, device):
            if self._arena is not None:
                buf = self._arena.get("w_rem_dev")
                if not self._on(buf, device):
                    buf = self._arena["w_rem_dev"] = torch.zeros(2, device=device, dtype=torch.float32)
                self._w_rem_dev = buf
            else:
                self._w_rem_dev = torch.zeros(2, device=device, dtype=torch.float32)
            self._w_rem_host = None
        if self._w_rem_host != w:
            self._w_rem_dev[0].fill_(w[0])
            self._w_rem_dev[1].fill_(w[1])
            self._w_rem_host = w

    # -- per-resolution cache -----------------------------------------------------------------------------------------
    def _coords512(self, transform_coords, device):
        tc = transform_coords
        if not torch.is_tensor(tc):
            tc = torch.as_tensor(np.asarray(tc))
        return tc.to(device=device, dtype=torch.float32).contiguous()

    def _ensure_dilated(self, device):
        """the remover dilates its image mask in the constructor (:986); the editor does not"""

    def _ensure_mask_new_warped(self, transform_coords, device):
        """editor.py:147-149 / attention_processors.py:517-523: binarised forward-splat of the object mask at image resolution
        (the controller's image mask as its constructor left it: dilated by 5 px for the remover, :986)."""
        self._ensure_dilated(device)
        if self.mask_new_warped is None:
            tc = self._coords512(transform_coords, device)[:1]
            img_mask = self.image_mask.to(device=device, dtype=torch.float32)
            idx, _, d2 = geometry.splat_index(tc)
            w = geometry.splat_composite(img_mask[:1, None].contiguous(), idx, d2, binarize=True)
            self.mask_new_warped = w.tile(img_mask.shape[0], 1, 1, 1).detach()
        return self.mask_new_warped

    def _get_cache(self, S, transform_coords, device):
        raise NotImplementedError

    def forward(self, q, k, v, is_cross: bool, place_in_unet: str, transform_coords=None, scale=None, mask=None):
        if self.use_cfg:
            h = q.shape[0] // (2 * self.batch_size)
        else:
            h = q.shape[0] // self.batch_size
        n_entries = int(self.coords_edit[1])          # the edit sample is the last batch entry (attention_processors.py:56-67)
        if q.shape[0] % n_entries == 0 and q.shape[0] // n_entries != h:
            h = q.shape[0] // n_entries               # batch without the dead unconditional reference sample (diffusion.diffusion_step)
        base_mode = self.base_mode if (self.base_mode == "write" and not self.use_cfg) or (self.base_mode == "read" and self.use_cfg) else None
        if not (is_cross or (self.num_self_replace[0] <= self.cur_step < self.num_self_replace[1])):
            if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
                # the reference back-propagates through compute_attention here (:646-647); the batch driver never optimises past the
                # self-replace window (optimize_steps < self_replace_steps), and the plain kernel is forward-only: refuse, do not drop it
                raise NotImplementedError("gradient through a self-attention layer outside the self-replace window "
                                          "(optimize_steps > self_replace_steps) is not served by this path")
            return Fn.plain_attention(q, k, v, scale, h)
        self._ensure_device_state(q.device)
        N = q.shape[1]
        S = int(np.sqrt(N))
        cache = self._get_cache(S, transform_coords, q.device)
        with_loss = N >= 32 ** 2 and (not self.use_cfg)
        att = "cross" if is_cross else "self"
        # the device copy of the removal weight is used only while it is known to be current (sync_device_weights); otherwise the host value
        w_dev = None
        if with_loss and self._w_rem_dev is not None and self._w_rem_dev.device == q.device and self._w_rem_host == (
                float(self.loss_weight_dict["self"].get("removal", 0.0)), float(self.loss_weight_dict["cross"].get("removal", 0.0))):
            w_dev = self._w_rem_dev[1:2] if is_cross else self._w_rem_dev[0:1]
        spec = Fn.LayerSpec(kind=self.KIND, is_cross=is_cross, heads=h, cb=tuple(self.coords_base), ce=tuple(self.coords_edit),
                            scale=float(scale), blend=self.cur_step < int(self.num_steps * self.obj_edit_step), with_loss=with_loss,
                            weights=self.loss_weight_dict[att], cache=cache, log_accum=self._log_accum[1 if is_cross else 0], w_rem_dev=w_dev)
        if base_mode is not None:
            # one store per attention layer (cur_att_layer walks the 32 layers in the same order in both passes); it lives with the per-model
            # arena when there is one, so the CFG-pass graphs that read it survive across edits
            stores = self._arena.setdefault("base_stores", {}) if self._arena is not None else self._base_stores
            spec.base_mode, spec.base_store = base_mode, stores.setdefault((self.cur_att_layer, int(q.shape[1]), int(k.shape[1])), {})
        out, loss, _ = Fn.shared_attention_layer(q, k, v, spec)
        if N >= 32 ** 2:
            self.mask_wo_edit = cache.masks["mask_wo_edit"][None, None]
            self.mask_1_empty = cache.masks["mask_1_empty"][None, None]
            if not is_cross:
                self.mask_inpaint = cache.masks["mask_1_empty"].clone()
        if with_loss:
            self.loss = self.loss + loss
            self.loss_log_dict["num_layers"] += 1
        if self.use_cfg and self.store_attention_maps and N <= 16 ** 2 and base_mode != "read":
            # attention_processors.py:452-454, 562-564: the edit stream's map (edit queries against the base keys; its own text keys on cross
            # layers).  A pass that stores maps runs eagerly (graphs.edit_pass): the store is host state.
            ce, cb = int(self.coords_edit[0]), int(self.coords_base[0])
            lay = Fn._Layout(q.t if isinstance(q, Fn.ProjView) else q, k.t if isinstance(k, Fn.ProjView) else k, h, isinstance(q, Fn.ProjView))
            qt, kt = (q.t if isinstance(q, Fn.ProjView) else q), (k.t if isinstance(k, Fn.ProjView) else k)
            q_e, k_e = lay.sl(qt, ce), lay.sl(kt, ce if is_cross else cb)
            if isinstance(q, Fn.ProjView):
                q_e, k_e = Fn.ProjView(q_e[None], h), Fn.ProjView(k_e[None], h)
            self.attn_store(Fn.attention_maps(q_e, k_e, scale, h), is_cross, place_in_unet)
        return out

    # reference-named entry points kept for callers that invoke them directly (attention_processors.py:384, 513)
    def replace_self_attention(self, q, k, v, place_in_unet, transform_coords=None, add_empty=None, scale=None, mask=None,
                               coords_base=None, coords_edit=None, old_attention_map=None, old_attention_out=None):
        return self._replace(q, k, v, False, transform_coords, scale, coords_base, coords_edit)

    def replace_cross_attention(self, q, k, v, place_in_unet, transform_coords=None, add_empty=False, scale=None, mask=None,
                                coords_base=None, coords_edit=None, old_attention_map=None, old_attention_out=None):
        return self._replace(q, k, v, True, transform_coords, scale, coords_base, coords_edit)

    def _replace(self, q, k, v, is_cross, transform_coords, scale, coords_base, coords_edit):
        cb, ce = self.coords_base, self.coords_edit
        if coords_base is not None:
            self.coords_base, self.coords_edit = coords_base, coords_edit
        saved = self.num_self_replace
        self.num_self_replace = (0, 10 ** 9)
        try:
            out = self.forward(q, k, v, is_cross, "", transform_coords=transform_coords, scale=scale)
        finally:
            self.num_self_replace = saved
            self.coords_base, self.coords_edit = cb, ce
        h = out.shape[0] - self.coords_base[-1] * (q.shape[0] // ((2 if self.use_cfg else 1) * self.batch_size))
        return out[-h:][None]


class AttentionGeometryEdit(_GeometryControllerBase):
    KIND = "edit"
    LOG_KEYS = ("sim", "movement", "removal", "smoothness")  # `amodal` is absent from the reference's initialiser (:626-630)

    def __init__(self, prompts, num_steps: int, cross_replace_steps, self_replace_steps, equalizer=None, local_blend=None,
                 controller=None, image_mask=None, empty_scale=0.2, use_all=True, obj_edit_step=0.0, tokenizer=None, device="cuda:0",
                 mode="bilinear"):
        super().__init__()
        self._common_init(prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend, controller, image_mask, empty_scale,
                          use_all, obj_edit_step, mode)
        if image_mask is not None:
            image_mask = torch.from_numpy(np.asarray(image_mask)[None])
            self.image_mask = image_mask.tile((len(prompts), 1, 1))
        self.default_loss_weights = {"self": {"sim": 110, "movement": 13.5, "removal": 1.67, "smoothness": 35.0, "amodal": 80.5},
                                     "cross": {"sim": 60, "movement": 6.34, "removal": 1.6, "smoothness": 20.0, "amodal": 3.5}}
        self.initialize_loss_log_dict()
        self.initialize_default_loss_weights()

    def _get_cache(self, S, transform_coords, device):
        c = self._res_cache.get(S)
        if c is None:
            mnw = self._ensure_mask_new_warped(transform_coords, device)
            amodal = self.amodal_mask
            if amodal is None:
                raise RuntimeError("controller.amodal_mask must be set before the first attention call (editor.py:633)")
            amodal = torch.as_tensor(amodal).to(device=device, dtype=torch.float32)
            img_mask = self.image_mask.to(device=device, dtype=torch.float32)
            masks = geometry.build_masks(img_mask[-1].contiguous(), mnw[-1, 0].contiguous(), amodal.reshape(amodal.shape[-2:]).contiguous(), S)
            coords_S = geometry.reshape_transform_coords(self._coords512(transform_coords, device)[:1], in_mat_shape=(1, 1, S, S))[0]
            c = Fn.ResolutionCache(S, masks, coords_S=coords_S, need_amodal=S * S > 32 ** 2, arena=self._arena)
            self._res_cache[S] = c
            self.masks_cache_dict[S] = dict(masks, t_coords_q=coords_S)
        return c


class AttentionGeometryRemover(_GeometryControllerBase):
    KIND = "remove"
    LOG_KEYS = ("sim", "removal", "smoothness")

    def __init__(self, prompts, num_steps: int, cross_replace_steps, self_replace_steps, equalizer=None, local_blend=None,
                 controller=None, image_mask=None, empty_scale=0.2, use_all=True, obj_edit_step=0.0, tokenizer=None, device="cuda:0",
                 mode="bilinear"):
        super().__init__()
        self._common_init(prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend, controller, image_mask, empty_scale,
                          use_all, obj_edit_step, mode)
        self._dilated = False
        if image_mask is not None:
            image_mask = torch.from_numpy(np.asarray(image_mask)[None])
            self.image_mask = image_mask.tile((len(prompts), 1, 1))  # dilated by 5 px on first device use (:986)
        self.default_loss_weights = {"self": {"sim": 110.0, "removal": 3.6, "smoothness": 35.0},
                                     "cross": {"sim": 60.0, "removal": 3.6, "smoothness": 20.0}}
        self.initialize_default_loss_weights()
        self.initialize_loss_log_dict()

    def _ensure_dilated(self, device):
        """:986 -- done once, on first device use, whichever of the warped-mask build (editor.py:148) and the first attention call comes first"""
        if not self._dilated:
            self.image_mask = geometry.torch_dilate(self.image_mask.to(device=device, dtype=torch.float32)[:, None], 5)[:, 0]
            self._dilated = True

    def _get_cache(self, S, transform_coords, device):
        c = self._res_cache.get(S)
        if c is None:
            self._ensure_dilated(device)
            masks = geometry.build_masks(self.image_mask[-1].contiguous(), None, None, S)
            c = Fn.ResolutionCache(S, masks, coords_S=None, need_amodal=False, arena=self._arena)
            self._res_cache[S] = c
            self.masks_cache_dict[S] = dict(masks)
        return c
