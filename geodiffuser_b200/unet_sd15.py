"""Random-init UNet with the Stable-Diffusion-1.x topology, in plain PyTorch -- the CALLER of the hot path, present only so that
the edit loop can be measured and tested end to end (diffusers is not installed; BASELINE.json configs use random-init weights).

Not part of the product path: convolutions / linears / norms run through stock torch (cuDNN / cuBLAS), exactly as they do under
the reference.  What matters is that every `Attention` module exposes the diffusers-0.25 interface the reference's processors use
(spatial_norm, group_norm, norm_cross, to_q/to_k/to_v/to_out, head_to_batch_dim, batch_to_head_dim, prepare_attention_mask, scale,
heads, residual_connection, rescale_output_factor; vendored copy of that class: Evaluation/DiffusionHandles/diffhandles/model/
attention_processor.py:36-640) and that `unet.attn_processors` / `unet.set_attn_processor` name the 32 attention layers
`{down_blocks,mid_block,up_blocks}...attn{1,2}.processor` as register_attention_control_diffusers expects.

Topology (SD-1.5 config): in/out 4 ch, block_out (320, 640, 1280, 1280), 2 layers per block, 8 heads, cross-attention dim 768,
GroupNorm(32), GEGLU feed-forward, conv proj_in/proj_out, timestep embedding 320 -> 1280.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


import os

from .body_ops import add_bias_residual, conv1x1, fast_body, geglu, group_norm_act, layer_norm

BODY_CHANNELS_LAST = os.environ.get("GD_BODY_NCHW", "0") != "1"   # bf16 body layout: NHWC (cuDNN's native tensor-core layout) unless overridden


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.is_cross = cross_attention_dim is not None
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.processor = None

    def set_processor(self, processor):
        self.processor = processor

    def head_to_batch_dim(self, tensor):
        b, n, c = tensor.shape
        h = self.heads
        return tensor.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, tensor):
        bh, n, d = tensor.shape
        h = self.heads
        return tensor.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        return attention_mask  # always None on this path

    def get_attention_scores(self, query, key, attention_mask=None):
        raise RuntimeError("score materialisation is not part of this build: use the fused kernels")

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask, **kw)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        return geglu(self.proj(x))


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = self.attn1(layer_norm(self.norm1, x)) + x
        x = self.attn2(layer_norm(self.norm2, x), encoder_hidden_states=context) + x
        return self.ff(layer_norm(self.norm3, x)) + x


class Transformer2DModel(nn.Module):
    def __init__(self, channels, heads, cross_attention_dim, groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.proj_in = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, channels // heads, cross_attention_dim)])
        self.proj_out = nn.Conv2d(channels, channels, 1)

    def forward(self, x, context):
        b, c, h, w = x.shape
        res = x
        x = conv1x1(self.proj_in, group_norm_act(self.norm, x))
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            x = blk(x, context)
        x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
        return conv1x1(self.proj_out, x) + res


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_ch, groups=32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def _shift_bias(self):
        # conv1.bias + time_emb_proj.bias: one (C) vector, so that the per-sample shift below is a single GEMM with a bias epilogue
        b = self.__dict__.get("_tb")
        if b is None or b.device != self.conv1.bias.device or b.dtype != self.conv1.bias.dtype:
            b = self.__dict__["_tb"] = (self.conv1.bias.detach().float() + self.time_emb_proj.bias.detach().float()).to(self.conv1.bias.dtype)
        return b

    def forward(self, x, temb):
        if fast_body(x) and not self.conv1.bias.requires_grad:
            # product setting (bf16, channels-last, frozen weights): conv1's bias and the time-embedding shift are folded into norm2
            # (one (B, C) vector added while the norm reads its input), conv2's bias into the residual add
            h = F.conv2d(group_norm_act(self.norm1, x, silu=True), self.conv1.weight, None, padding=1)
            shift = F.linear(F.silu(temb), self.time_emb_proj.weight, self._shift_bias())
            h = F.conv2d(group_norm_act(self.norm2, h, silu=True, pre_bias=shift), self.conv2.weight, None, padding=1)
            if self.conv_shortcut is not None:
                x = conv1x1(self.conv_shortcut, x)
            return add_bias_residual(x, h, self.conv2.bias)
        h = self.conv1(group_norm_act(self.norm1, x, silu=True))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(group_norm_act(self.norm2, h, silu=True))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb_ch, heads, ctx_dim, has_attn, add_down, layers=2):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_ch) for i in range(layers)])
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, ctx_dim) for _ in range(layers)]) if has_attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, context):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, ch, temb_ch, heads, ctx_dim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch), ResnetBlock2D(ch, ch, temb_ch)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, ctx_dim)])

    def forward(self, x, temb, context):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, context)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin, cout, prev_out, temb_ch, heads, ctx_dim, has_attn, add_up, layers=3):
        super().__init__()
        res = []
        for i in range(layers):
            skip_ch = cin if i == layers - 1 else cout
            res.append(ResnetBlock2D((prev_out if i == 0 else cout) + skip_ch, cout, temb_ch))
        self.resnets = nn.ModuleList(res)
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, ctx_dim) for _ in range(layers)]) if has_attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, context):
        for i, r in enumerate(self.resnets):
            x = r(torch.cat([x, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


def timestep_embedding(t, dim):
    """diffusers Timesteps(320, flip_sin_to_cos=True, downscale_freq_shift=0)"""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class UNet2DConditionModel(nn.Module):
    def __init__(self, block_out=(320, 640, 1280, 1280), heads=8, ctx_dim=768, in_ch=4, out_ch=4):
        super().__init__()
        temb_ch = block_out[0] * 4
        self.block_out = block_out
        self.conv_in = nn.Conv2d(in_ch, block_out[0], 3, padding=1)
        self.time_embedding = nn.ModuleDict(dict(linear_1=nn.Linear(block_out[0], temb_ch), linear_2=nn.Linear(temb_ch, temb_ch)))
        downs, ch = [], block_out[0]
        for i, co in enumerate(block_out):
            last = i == len(block_out) - 1
            downs.append(DownBlock(ch, co, temb_ch, heads, ctx_dim, has_attn=not last, add_down=not last))
            ch = co
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(block_out[-1], temb_ch, heads, ctx_dim)
        ups, rev = [], list(reversed(block_out))
        prev = rev[0]
        for i, co in enumerate(rev):
            cin = rev[min(i + 1, len(rev) - 1)]
            ups.append(UpBlock(cin, co, prev, temb_ch, heads, ctx_dim, has_attn=i > 0, add_up=i < len(rev) - 1))
            prev = co
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(32, block_out[0], eps=1e-5)
        self.conv_out = nn.Conv2d(block_out[0], out_ch, 3, padding=1)
        from .attention_processors import VanillaAttentionProcessor

        self.set_attn_processor(VanillaAttentionProcessor())

    # ---- the processor registry the reference's hook API drives (diffusers UNet2DConditionModel.attn_processors) ----
    def _attention_modules(self):
        # the module tree is fixed after construction: walk it once (diffusers re-walks it on every access of `attn_processors`,
        # which the reference's set_attn_processor_for_edit does 33 times per call and 67 times per edit)
        mods = self.__dict__.get("_attn_mods")
        if mods is None:
            mods = [(name, m) for name, m in self.named_modules() if isinstance(m, Attention)]
            self.__dict__["_attn_mods"] = mods
        return mods

    @property
    def attn_processors(self):
        return {f"{name}.processor": m.processor for name, m in self._attention_modules()}

    def set_attn_processor(self, processor):
        mods = dict(self._attention_modules())
        if isinstance(processor, dict):
            if len(processor) != len(mods):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does not match the "
                                 f"number of attention layers: {len(mods)}.")
            for name, m in mods.items():
                m.set_processor(processor[f"{name}.processor"])
        else:
            for m in mods.values():
                m.set_processor(processor)

    def forward(self, sample, timestep, encoder_hidden_states):
        # the body runs in the dtype of its own weights (bf16 copy in channels_last = product setting, fp32 = parity setting); inputs are
        # cast on entry (differentiable), so no autocast pass re-casts the weights on every evaluation
        wdtype = self.conv_in.weight.dtype
        sample = sample.to(wdtype)
        if wdtype != torch.float32 and BODY_CHANNELS_LAST:
            sample = sample.contiguous(memory_format=torch.channels_last)
        encoder_hidden_states = encoder_hidden_states.to(wdtype)
        t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=sample.device)
        t = t.reshape(-1).to(sample.device).expand(sample.shape[0])
        temb = timestep_embedding(t, self.block_out[0]).to(sample.dtype)
        temb = self.time_embedding["linear_2"](F.silu(self.time_embedding["linear_1"](temb)))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states)
            skips += outs
        x = self.mid_block(x, temb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states)
        x = self.conv_out(group_norm_act(self.conv_norm_out, x, silu=True))
        return {"sample": x}


class EditModel:
    """The `model` object the reference loop passes around (`ldm_stable`): .unet, .scheduler, .device (diffusion.py:99-140).
    Tokenizer / text encoder / VAE are out of scope (synthetic context embeddings and latents)."""

    def __init__(self, unet, scheduler, device):
        self._unets = {torch.float32: unet}
        self.scheduler, self.device = scheduler, device

    @property
    def unet(self):
        """The UNet body in the caller precision selected by diffusion.set_body_dtype: the fp32 master (parity runs) or a bf16,
        channels_last copy of it made on first use (the product / bench setting; same weight rounding autocast would apply per call)."""
        from . import diffusion

        dt = diffusion.AUTOCAST_DTYPE if self.device.type == "cuda" else torch.float32
        if dt not in self._unets:
            import copy

            master = self._unets[torch.float32]
            procs = master.attn_processors
            master.set_attn_processor(None)          # processors (and the controller state behind them) are shared, not copied
            u = copy.deepcopy(master).to(dt)
            if BODY_CHANNELS_LAST:
                u = u.to(memory_format=torch.channels_last)
            master.set_attn_processor(procs)
            u.set_attn_processor(procs)
            self._unets[dt] = u
        return self._unets[dt]


def replicate_model(model):
    """A second EditModel over the SAME weight tensors: own module objects, attention processors, scheduler state, CUDA graphs and cache
    arenas (everything an edit mutates), shared parameters and buffers (nothing an edit mutates).  runner.EditWorkers uses one replica per
    concurrent edit lane of a GPU."""
    import copy
    from .attention_processors import VanillaAttentionProcessor
    from .diffusion import DDIMScheduler

    model.unet      # materialise the body in the current caller precision before copying the module tree
    new = EditModel.__new__(EditModel)
    new.scheduler, new.device, new._unets = DDIMScheduler(), model.device, {}
    for dt, u in model._unets.items():
        procs = u.attn_processors
        u.set_attn_processor(None)
        memo = {id(t): t for t in list(u.parameters()) + list(u.buffers())}      # deepcopy maps every weight tensor to itself: shared
        c = copy.deepcopy(u, memo)
        u.set_attn_processor(procs)
        c.set_attn_processor(VanillaAttentionProcessor())
        new._unets[dt] = c
    return new


def build_model(device="cuda", seed=1234, tiny=False):
    """Random-init SD-1.5 topology under torch.manual_seed(seed) (CPU generator => identical weights on every box)."""
    from .diffusion import DDIMScheduler

    g = torch.get_rng_state()
    torch.manual_seed(seed)
    unet = UNet2DConditionModel(block_out=(64, 128, 256, 256)) if tiny else UNet2DConditionModel()
    torch.set_rng_state(g)
    for p in unet.parameters():
        p.requires_grad = False
    unet = unet.to(device).eval()
    if torch.device(device).type == "cuda" and os.environ.get("GD_CUDNN_BENCHMARK", "1") != "0":
        # the body's ~100 convolutions have a handful of fixed shapes: let cuDNN time its algorithms once per shape (first evaluation)
        # instead of using the heuristic pick: 1006 -> 972 ms per 50-step edit on a B200
        torch.backends.cudnn.benchmark = True
    return EditModel(unet, DDIMScheduler(), torch.device(device))
