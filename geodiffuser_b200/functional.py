"""Fused warped shared-attention layer: the operator the controllers call once per UNet attention layer.

One call replaces, for one layer, the reference's sequence (attention_processors.py:633-664 with :513-624 / :384-508 / :748-928):
plain attention of the untouched batch entries, splat-warp of the base queries, attention of warped and edit queries against
the shared base K/V, the five attention-map loss terms, the output blend -- and, in backward, dQ (dK for cross layers) through
softmax and through the loss terms.  All arithmetic is in the C-ABI CUDA library (`_lib`); torch provides memory and autograd
bookkeeping only.  There is no CPU path.
"""
import math

import numpy as np
import torch

from . import _lib, geometry
from ._lib import call, ptr, stream

TERM_NAMES = ("sim", "movement", "removal", "smoothness", "amodal")
REMOVAL_DQ_GEMM = True   # removal term of dQ as its own (M x Nk) @ (Nk x d) contraction after the tcgen05 backward (False: `extra` rows folded into that kernel)
CORR_SM100 = True   # removal-loss correlation of the self-attention levels on the tcgen05 kernel (False: materialised base map + mma.sync, for A/B tests)


def gaussian_kernel5():
    """generic_torch.py:27-54 with kernel_size=5, sigma = 5//2*2/6"""
    size, std = 5, (5 // 2 * 2 / 6.0)
    g = torch.arange(size, dtype=torch.float32)
    mean = (size - 1) / 2
    k1 = 1 / (std * math.sqrt(2 * math.pi)) * torch.exp(-(((g - mean) / (2 * std)) ** 2))
    k = k1[:, None] * k1[None, :]
    return (k / k.sum()).reshape(-1).tolist()


_GAUSS25 = None


def _gauss25():
    global _GAUSS25
    if _GAUSS25 is None:
        _GAUSS25 = _lib.host_f32(gaussian_kernel5())
    return _GAUSS25


class ResolutionCache:
    """Everything that depends only on (masks, correspondence field, S): built once per edit per attention resolution
    (the reference caches masks/coords per S, attention_processors.py:319-373, but re-rasterises the splat on every call)."""

    def __init__(self, S, masks, coords_S=None, need_amodal=False, arena=None):
        """`arena`: optional dict that outlives the edit (one per model and controller kind).  Everything a gradient-free pass reads from the
        cache (masks, blend coefficients, splat index) is then kept at a fixed address and refreshed in place, so CUDA graphs captured for
        one edit stay valid for the next (graphs.py)."""
        def keep(name, t):
            t = t.contiguous()
            if arena is None:
                return t
            buf = arena.get((S, name))
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype or buf.device != t.device:
                arena[(S, name)] = buf = t.clone()
                arena["generation"] = arena.get("generation", 0) + 1    # graphs that captured the old address are stale
            else:
                buf.copy_(t)
            return buf

        self.S = S
        self.N = S * S
        self.masks = {k: keep(k, v) for k, v in masks.items()}
        dev = masks["mask_wo_edit"].device
        f = lambda n: self.masks[n].reshape(-1)
        self.m_edit, self.m_bg, self.m_inp, self.m_am = f("mask_new_warped"), f("mask_wo_edit"), f("mask_1_empty"), f("amodal_mask")
        self.one_minus_m_edit = keep("one_minus_m_edit", 1.0 - self.m_edit)
        self.m_inp_plus_bg = keep("m_inp_plus_bg", self.m_inp + self.m_bg)
        rows = torch.nonzero(self.m_inp > 0.5).reshape(-1).to(torch.int32)
        self.M = int(rows.numel())
        rowmap = torch.full((self.N,), -1, device=dev, dtype=torch.int32)
        if self.M:
            rowmap[rows.long()] = torch.arange(self.M, device=dev, dtype=torch.int32)
        # (in the arena as well: an optimisation-pass graph recorded for one edit then serves every later edit with the same inpaint-row
        #  counts and mask sums -- graphs.grad_pass keys on them)
        self.rows, self.rowmap = keep("rows", rows), keep("rowmap", rowmap)
        self.sum_bg, self.sum_edit, self.sum_inp = float(self.m_bg.sum()), float(self.m_edit.sum()), float(self.m_inp.sum())
        self.idx = self.dist2 = None
        if coords_S is not None:
            idx, _, dist2 = geometry.splat_index(coords_S[None])
            self.idx, self.dist2 = keep("idx", idx), keep("dist2", dist2)
        self.knn_idx = self.knn_val = self.knn_w = None
        self.sum_w_am = 0.0
        if need_amodal:
            def buf(name, shape, dtype):
                b = arena.get((S, name)) if arena is not None else None
                if b is None or b.shape != torch.Size(shape) or b.device != dev:
                    b = torch.empty(*shape, device=dev, dtype=dtype)
                    if arena is not None:
                        arena[(S, name)] = b
                        arena["generation"] = arena.get("generation", 0) + 1
                return b
            self.knn_idx, self.knn_val, self.knn_w = buf("knn_idx", (self.N, 4), torch.int32), buf("knn_val", (self.N, 4), torch.float32), buf("knn_w", (self.N,), torch.float32)
            call("gd_amodal_knn", ptr(self.m_edit), S, ptr(self.knn_idx), ptr(self.knn_val), ptr(self.knn_w), stream())
            self.sum_w_am = float((self.knn_w * self.m_am).sum())


def _bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


class ProjView:
    """q / k / v exactly as the to_q / to_k / to_v projections produce them -- one contiguous (B, N, H*d) tensor -- presented with the
    (B*H, N, d) shape the reference's controllers receive (attention_sharing.py:127-144).  The kernels address each (H, N, d) slab of it in
    place (row stride H*d, head stride d), so the head_to_batch_dim / batch_to_head_dim permute copies of the reference (4 per attention
    layer, 128 per UNet evaluation) do not exist on this path.  A controller that is handed ProjViews returns the attention output in the
    same projection layout, (B, N, H*d), ready for to_out."""
    __slots__ = ("t", "heads")

    def __init__(self, t, heads):
        assert t.dim() == 3 and t.shape[2] % heads == 0
        self.t = t if t.is_contiguous() else t.contiguous()
        self.heads = heads

    @property
    def shape(self):
        b, n, c = self.t.shape
        return torch.Size((b * self.heads, n, c // self.heads))

    device = property(lambda self: self.t.device)
    is_cuda = property(lambda self: self.t.is_cuda)
    dtype = property(lambda self: self.t.dtype)
    requires_grad = property(lambda self: self.t.requires_grad)


class _Layout:
    """element strides of the (H, N, d) slabs of q, k / v and of the outputs, for the two layouts an operand tensor can have:
    heads-major (B*H, N, d) [reference] or projection (B, N, H*d) [ProjView]"""

    def __init__(self, q, k, heads, proj):
        self.proj, self.h = proj, heads
        if proj:
            self.N, C = q.shape[1], q.shape[2]
            self.Nk = k.shape[1]
            self.d = C // heads
            self.q = (C, self.d)
            self.kv = (k.shape[2], self.d)
        else:
            self.N, self.d = q.shape[1], q.shape[2]
            self.Nk = k.shape[1]
            self.q = (self.d, self.N * self.d)
            self.kv = (self.d, self.Nk * self.d)

    def sl(self, t, i):
        return t[i] if self.proj else t[i * self.h:(i + 1) * self.h]

    def new_batch(self, like, entries, n_tokens):
        """an uninitialised tensor of `entries` batch entries in this layout"""
        if self.proj:
            return torch.empty(entries, n_tokens, self.h * self.d, device=like.device, dtype=like.dtype)
        return torch.empty(entries * self.h, n_tokens, self.d, device=like.device, dtype=like.dtype)

    def strides(self, out=None):
        o = self.q if out is None else out
        return _lib.host_longs([self.q[0], self.q[1], self.kv[0], self.kv[1], o[0], o[1]])


def attention_forward(qs, ks, vs, scale, dims=None, strides=None, want32=None, os=None, os_is_bf16=False):
    """G query streams against their K/V slabs.  -> (O32, LSE): O32[g] (H,N,d) fp32 for the streams in `want32` (default: all) else None;
    LSE (G,H,N) fp32.  `os[g]`: optional strided output slab per stream (bf16 / fp32), written by the kernel in the layout of q.
    Without `dims`/`strides` the operands are contiguous (H,N,d) tensors."""
    G = len(qs)
    if dims is None:
        H, N, d = qs[0].shape
        Nk = ks[0].shape[1]
    else:
        H, N, Nk, d = dims
    dev = qs[0].device
    want32 = range(G) if want32 is None else want32
    buf = torch.empty(len(want32), H, N, d, device=dev, dtype=torch.float32) if len(want32) else None
    O32 = [None] * G
    for i, g in enumerate(want32):
        O32[g] = buf[i]
    LSE = torch.empty(G, H, N, device=dev, dtype=torch.float32)
    entry = "gd_attn_fwd_generic"
    if _lib.HAS_SM100 and Nk == N and N % 128 == 0 and d in (40, 80) and N >= 1024:
        entry = "gd_attn_fwd_sm100"
    call(entry, _lib.ptr_array(qs), _lib.ptr_array(ks), _lib.ptr_array(vs), _lib.ptr_array(O32), _lib.ptr_array([LSE[g] for g in range(G)]),
         _lib.ptr_array(os) if os is not None else None, G, H, N, Nk, d, float(scale), strides, int(os_is_bf16), stream(), tag=(G, H, N, Nk, d))
    return O32, LSE


def warp_queries(q_base, cache, lay=None, out=None):
    """q_base: one (H,N,d) bf16 slab -> q*(1-M) + M*splat(q)  (attention_processors.py:544), bf16, in the layout of q"""
    if lay is None:
        return geometry.splat_composite(q_base, cache.idx, cache.dist2, channels_last=True, blend_mask=cache.m_edit, out_dtype=torch.bfloat16)
    st = _lib.host_longs(lay.q)
    S = cache.S
    r2 = float(np.float32(pow(geometry.splat_radius_ndc(S, None), 2)))
    call("gd_splat_composite_rows", _lib.base_ptr(q_base), 1, st, ptr(cache.idx), ptr(cache.dist2), lay.h, lay.N, lay.d, cache.idx.shape[-1], r2,
         float(geometry.SPLATTER.tau), ptr(cache.m_edit), 0, _lib.base_ptr(out), 1, st, stream())
    return out


class LayerSpec:
    """Static description of one controller call."""
    __slots__ = ("kind", "is_cross", "heads", "cb", "ce", "scale", "blend", "with_loss", "weights", "cache", "log_accum", "w_rem_dev", "base_mode",
                 "base_store")

    def __init__(self, **kw):
        # the base (reference) sample is identical in the optimisation pass and the CFG pass of one timestep.  base_mode "write":
        # this call stores the base K / V slabs and the warped-stream output in `base_store` (a dict of persistent tensors, filled on first use);
        # "read": the batch holds no base sample -- [.., edit] -- and those tensors are taken from `base_store` instead of being recomputed.
        self.base_mode, self.base_store = None, None
        self.w_rem_dev = None   # optional device scalar holding weights["removal"], kept current by the controller: lets a captured pass
        for k, v in kw.items():  # follow the adaptive schedule
            setattr(self, k, v)


def _forward_impl(q, k, v, spec, proj=False):
    """Runs the fused layer on q, k, v tensors in heads-major (B*H, N, d) or (proj) projection (B, N, H*d) layout.
    Returns (out in the same layout, terms6 or None, saved-for-backward dict or None)."""
    h, (cb0, cb1), (ce0, ce1) = spec.heads, spec.cb, spec.ce
    assert ce1 - ce0 == 1 and cb1 - cb0 == 1
    cache = spec.cache
    lay = _Layout(q, k, h, proj)
    N, d, Nk = lay.N, lay.d, lay.Nk
    qb, kb, vb = _bf16(q), _bf16(k), _bf16(v)
    sl = lay.sl
    bp = _lib.base_ptr
    qs = [sl(qb, i) for i in range(cb1)]
    ks = [sl(kb, i) for i in range(cb1)]
    vs = [sl(vb, i) for i in range(cb1)]
    q_e = sl(qb, ce0)
    assert q.dtype in (torch.bfloat16, torch.float32)
    is_bf16 = q.dtype == torch.bfloat16
    out = lay.new_batch(q, cb1 + 1, N)                   # plain entries 0..cb1-1, then the edit entry
    os = [sl(out, i) for i in range(cb1)]                # plain streams: written by the attention kernel itself, in place, in q's dtype
    q_w = None
    if spec.kind == "edit":
        q_w = warp_queries(sl(qb, cb0), cache, lay, lay.new_batch(qb, 1, N))
        k_e = sl(kb, ce0) if spec.is_cross else sl(kb, cb0)
        qs += [sl(q_w, 0), q_e]
        ks += [sl(kb, cb0), k_e]
        vs += [sl(vb, cb0), sl(vb, cb0)]
        g_e = cb1 + 1
        want32 = [cb1, cb1 + 1]
    else:
        k_e = sl(kb, cb0)
        qs += [q_e]
        ks += [k_e]
        vs += [sl(vb, cb0)]
        g_e = cb1
        want32 = [cb0, cb1]
        if not spec.blend:
            qs += [q_e]
            ks += [sl(kb, ce0)]
            vs += [sl(vb, ce0)]
            want32.append(cb1 + 1)
    os += [None] * (len(qs) - cb1)
    O, LSE = attention_forward(qs, ks, vs, spec.scale, dims=(h, N, Nk, d), strides=lay.strides(), want32=want32, os=os, os_is_bf16=is_bf16)
    r = O[g_e]
    e = O[cb1] if spec.kind == "edit" else O[cb0]
    out_edit = sl(out, cb1)
    ost = _lib.host_longs(lay.q)
    if spec.kind == "edit":
        if spec.blend:
            coef = cache.one_minus_m_edit
            call("gd_blend_rows", ptr(e), ptr(cache.m_edit), ptr(r), ptr(coef), h, N, d, bp(out_edit), int(is_bf16), ost, stream())
        else:
            coef = None
            call("gd_blend_rows", None, None, ptr(r), None, h, N, d, bp(out_edit), int(is_bf16), ost, stream())
    else:
        if spec.blend:
            coef = cache.m_inp_plus_bg
            call("gd_blend_rows", None, None, ptr(r), ptr(coef), h, N, d, bp(out_edit), int(is_bf16), ost, stream())
        else:
            coef = cache.m_bg
            call("gd_blend_rows", ptr(O[g_e + 1]), ptr(cache.m_inp), ptr(r), ptr(coef), h, N, d, bp(out_edit), int(is_bf16), ost, stream())
    if spec.base_mode == "write":
        _base_store_write(spec.base_store, lay, sl(kb, cb0), sl(vb, cb0), e if spec.kind == "edit" else None)
    saved = dict(q_e=q_e, k_e=k_e, v_e=sl(vb, cb0), o_e=r, lse_e=LSE[g_e], g_loss=None, extra=None, delta_extra=None, coef=coef,
                 ld=(Nk + 7) // 8 * 8, M=0, lay=lay, keep=(qb, kb, vb, q_w))
    if not spec.with_loss:
        return out, None, saved

    # ---- loss terms + their gradient field ---------------------------------------------------------------------------
    dev = q.device
    S = cache.S
    lw = spec.weights
    w_sim, w_mov = float(lw.get("sim", 0.0)), float(lw.get("movement", 0.0))
    w_rem, w_sm, w_amo = float(lw.get("removal", 0.0)), float(lw.get("smoothness", 0.0)), float(lw.get("amodal", 0.0))
    if spec.kind != "edit":
        w_mov = w_amo = 0.0
    use_amodal = spec.kind == "edit" and N > 32 ** 2
    inv_sim = 1.0 / (cache.sum_bg * h * d + 1e-8)
    inv_mov = 1.0 / (cache.sum_edit * h * d + 1e-8)
    inv_amo = 1.0 / (cache.sum_w_am * h * d + 1e-8) if use_amodal else 0.0
    inv_smh = 1.0 / (h * (S - 1) * S * d)
    inv_rem = 1.0 / (cache.sum_inp * h + 1e-8)
    t = None
    if use_amodal:
        scratch = torch.empty(h, N, d, device=dev, dtype=torch.float32)
        t = torch.empty(h, N, d, device=dev, dtype=torch.float32)
        call("gd_amodal_target", ptr(e), ptr(cache.m_edit), ptr(cache.knn_idx), ptr(cache.knn_val), _gauss25(), h, S, d, ptr(scratch), ptr(t), stream())
    n_part = 296
    partials = torch.empty(n_part, 5, device=dev, dtype=torch.float32)
    g_loss = torch.empty(h, N, d, device=dev, dtype=torch.float32)
    call("gd_attn_l1_losses", ptr(e), ptr(r), ptr(t), ptr(cache.m_bg), ptr(cache.m_edit) if spec.kind == "edit" else None,
         ptr(cache.m_am) if use_amodal else None, ptr(cache.knn_w) if use_amodal else None, w_sim * inv_sim, w_mov * inv_mov,
         (w_amo * inv_amo) if use_amodal else 0.0, w_sm * inv_smh, w_sm * inv_smh, h, S, d, ptr(g_loss), ptr(partials), n_part, stream())
    # removal loss: materialise the base map and the inpaint rows of the edit map (bf16), correlate, masked arg-max
    M = cache.M
    rem_terms = extra = delta_extra = None
    ld = (Nk + 7) // 8 * 8
    if M > 0:
        qk_st = _lib.host_longs([lay.q[0], lay.q[1], lay.kv[0], lay.kv[1]])
        a_e = torch.empty(h, M, ld, device=dev, dtype=torch.bfloat16)
        call("gd_attn_probs", bp(q_e), bp(k_e), ptr(LSE[g_e]), ptr(cache.rows), M, h, N, Nk, d, float(spec.scale), ptr(a_e), ld, qk_st, stream(), tag=(h, M, Nk, d))
        rem_terms = torch.empty(h * M, device=dev, dtype=torch.float32)
        g2 = torch.empty(h * M, 2, device=dev, dtype=torch.float32)
        j2 = torch.empty(h * M, 2, device=dev, dtype=torch.int32)
        delta_extra = torch.empty(h, M, device=dev, dtype=torch.float32)
        wdev = spec.w_rem_dev
        coef = inv_rem if wdev is not None else w_rem * inv_rem
        if CORR_SM100 and _lib.HAS_SM100 and Nk == N and N % 128 == 0 and d in (40, 80) and N >= 1024:
            # self-attention levels: the base map is never materialised -- its tiles are recomputed inside the tcgen05 correlation kernel
            # from q_b, k_b and the stored LSE, and the two rows per (h, m) the backward needs are recomputed as softmax rows
            n_tiles = N // 32
            partial = torch.empty(h, n_tiles, M, 4, device=dev, dtype=torch.float32)
            call("gd_removal_corr_sm100", bp(sl(qb, cb0)), bp(sl(kb, cb0)), ptr(LSE[cb0]), ptr(a_e), h, M, N, d, float(spec.scale), ld, qk_st,
                 ptr(cache.m_inp), ptr(cache.m_bg), ptr(partial), stream(), tag=(h, M, N, Nk))
            call("gd_removal_finalize", ptr(partial), n_tiles, h, M, S, ptr(cache.rows), ptr(cache.m_inp), ptr(cache.m_bg), coef, ptr(wdev), None,
                 N, Nk, ld, ptr(rem_terms), ptr(g2), ptr(j2), ptr(delta_extra), None, stream())
            p2 = torch.empty(h, 2 * M, ld, device=dev, dtype=torch.bfloat16)
            call("gd_attn_probs_rows2", bp(sl(qb, cb0)), bp(sl(kb, cb0)), ptr(LSE[cb0]), ptr(j2), M, h, N, Nk, d, float(spec.scale), ptr(p2), ld,
                 qk_st, stream())
            if REMOVAL_DQ_GEMM and Nk % 256 == 0:
                # the removal term of dS lives on the M inpaint rows only: W = A_e[rows] o dL/dA_e[rows] (bf16), contracted with K after the
                # tcgen05 backward (gd_removal_dq_rows), which then runs without `extra`
                ex_key_major, ex_ld = 2, ld
                extra = torch.empty(h, M, ld, device=dev, dtype=torch.bfloat16)
                call("gd_removal_weighted_rows", ptr(a_e), ptr(p2), ptr(g2), h, M, Nk, ld, ptr(extra), stream())
            else:
                # dL/dA_e rows, key-major (H, Nk, Mp): the layout the tcgen05 backward reads coalesced
                ex_key_major, ex_ld = 1, (M + 3) // 4 * 4
                extra = torch.empty(h, Nk, ex_ld, device=dev, dtype=torch.float32)
                call("gd_removal_extra_rows", ptr(p2), ptr(g2), h, M, Nk, ld, ptr(extra), 1, stream())
            del p2
        else:
            # cross layers (Nk = 77) and ragged shapes: materialise the base map (bf16, H x N x ld: 5 MB at Nk = 77) and correlate with mma.sync
            ex_key_major, ex_ld = 0, ld
            extra = torch.empty(h, M, ld, device=dev, dtype=torch.float32)
            a_b = torch.empty(h, N, ld, device=dev, dtype=torch.bfloat16)
            call("gd_attn_probs", bp(sl(qb, cb0)), bp(sl(kb, cb0)), ptr(LSE[cb0]), None, N, h, N, Nk, d, float(spec.scale), ptr(a_b), ld, qk_st, stream(), tag=(h, N, Nk, d))
            n_tiles = (N + 63) // 64
            partial = torch.empty(h, n_tiles, M, 4, device=dev, dtype=torch.float32)
            call("gd_corr_max_partial", ptr(a_e), ptr(a_b), h, M, N, Nk, ld, ptr(cache.m_inp), ptr(cache.m_bg), ptr(partial), stream(), tag=(h, M, N, Nk))
            call("gd_removal_finalize", ptr(partial), n_tiles, h, M, S, ptr(cache.rows), ptr(cache.m_inp), ptr(cache.m_bg), coef, ptr(wdev), ptr(a_b),
                 N, Nk, ld, ptr(rem_terms), ptr(g2), ptr(j2), ptr(delta_extra), ptr(extra), stream())
            del a_b
        del a_e, partial
    terms = torch.empty(6, device=dev, dtype=torch.float32)
    call("gd_loss_reduce", ptr(partials), n_part, ptr(rem_terms), h * M if M > 0 else 0,
         _lib.host_f32([inv_sim, inv_mov, inv_amo, inv_smh, inv_smh, inv_rem]), _lib.host_f32([w_sim, w_mov, w_amo, w_sm, w_rem]),
         ptr(spec.w_rem_dev), 1.0 if use_amodal else 0.0, ptr(terms), ptr(spec.log_accum), stream())
    saved.update(g_loss=g_loss, extra=extra, delta_extra=delta_extra, M=M)
    if M > 0:
        saved.update(ex_ld=ex_ld, ex_key_major=ex_key_major)
    return out, terms, saved


def _base_store_write(store, lay, k_base, v_base, e):
    """keeps one layer's base K / V slabs (bf16, in the operand layout, as a 1-entry batch) and, for the edit controller, the warped-stream
    output e (fp32 (H, N, d)) at fixed addresses: the CFG pass of the same timestep -- usually a CUDA graph -- reads them"""
    def put(name, src, make):
        buf = store.get(name)
        if buf is None or buf.shape != src.shape or buf.dtype != src.dtype:
            buf = store[name] = make()
            store["generation"] = store.get("generation", 0) + 1
        buf.copy_(src)
    put("k", k_base, lambda: lay.sl(lay.new_batch(k_base, 1, lay.Nk), 0))
    put("v", v_base, lambda: lay.sl(lay.new_batch(v_base, 1, lay.Nk), 0))
    if e is not None:
        put("e", e, lambda: torch.empty_like(e))


def _forward_cached_impl(q, k, v, spec, proj=False):
    """The CFG pass of a timestep whose optimisation pass has just run (spec.base_mode == "read").  q, k, v hold [plain entries .., edit] WITHOUT
    the base sample: its K / V and the warped-stream output come from spec.base_store.  Streams: the plain entries + the edit queries against the
    stored base K / V -- G = 2 instead of 4 -- then the same blend (attention_processors.py:617-624 / 922-925).  No loss terms (use_cfg)."""
    h = spec.heads
    cache, st = spec.cache, spec.base_store
    lay = _Layout(q, k, h, proj)
    N, d, Nk = lay.N, lay.d, lay.Nk
    n_plain = (q.shape[0] if proj else q.shape[0] // h) - 1
    qb, kb, vb = _bf16(q), _bf16(k), _bf16(v)
    sl, bp = lay.sl, _lib.base_ptr
    is_bf16 = q.dtype == torch.bfloat16
    kc, vc = st["k"], st["v"]
    assert kc.shape == sl(kb, 0).shape and kc.stride() == sl(kb, 0).stride(), "base store does not match this layer"
    q_e = sl(qb, n_plain)
    qs = [sl(qb, i) for i in range(n_plain)] + [q_e]
    ks = [sl(kb, i) for i in range(n_plain)] + [sl(kb, n_plain) if (spec.is_cross and spec.kind == "edit") else kc]
    vs = [sl(vb, i) for i in range(n_plain)] + [vc]
    want32 = [n_plain]
    if spec.kind != "edit" and not spec.blend:      # remover outside the blend window: the edit sample's own attention fills the inpaint rows (:925)
        qs.append(q_e); ks.append(sl(kb, n_plain)); vs.append(sl(vb, n_plain))
        want32.append(n_plain + 1)
    out = lay.new_batch(q, n_plain + 1, N)
    os = [sl(out, i) for i in range(n_plain)] + [None] * (len(qs) - n_plain)
    O, _ = attention_forward(qs, ks, vs, spec.scale, dims=(h, N, Nk, d), strides=lay.strides(), want32=want32, os=os, os_is_bf16=is_bf16)
    r, out_edit, ost = O[n_plain], sl(out, n_plain), _lib.host_longs(lay.q)
    if spec.kind == "edit":
        if spec.blend:
            call("gd_blend_rows", ptr(st["e"]), ptr(cache.m_edit), ptr(r), ptr(cache.one_minus_m_edit), h, N, d, bp(out_edit), int(is_bf16), ost, stream())
        else:
            call("gd_blend_rows", None, None, ptr(r), None, h, N, d, bp(out_edit), int(is_bf16), ost, stream())
    elif spec.blend:
        call("gd_blend_rows", None, None, ptr(r), ptr(cache.m_inp_plus_bg), h, N, d, bp(out_edit), int(is_bf16), ost, stream())
    else:
        call("gd_blend_rows", ptr(O[n_plain + 1]), ptr(cache.m_inp), ptr(r), ptr(cache.m_bg), h, N, d, bp(out_edit), int(is_bf16), ost, stream())
    return out


class _SharedAttentionLayerFn(torch.autograd.Function):
    """(q, k, v) -> (out, terms[6]); terms[5] is the weighted layer loss (differentiable), terms[0:5] the logged terms."""

    @staticmethod
    def forward(ctx, q, k, v, spec, proj):
        out, terms, saved = _forward_impl(q, k, v, spec, proj)
        ctx.spec, ctx.saved_dict = spec, saved
        ctx.counters = _lib.counters()
        ctx.shapes = (q.shape, k.shape, q.dtype, k.dtype)
        if terms is None:
            terms = torch.zeros(6, device=q.device, dtype=torch.float32)
            ctx.mark_non_differentiable(terms)
        return out, terms

    @staticmethod
    def backward(ctx, d_out, d_terms):
        with _lib.count_into(ctx.counters):
            return _SharedAttentionLayerFn._backward(ctx, d_out, d_terms)

    @staticmethod
    def _backward(ctx, d_out, d_terms):
        spec, s = ctx.spec, ctx.saved_dict
        h, (cb0, cb1), (ce0, ce1) = spec.heads, spec.cb, spec.ce
        q_shape, k_shape, q_dtype, k_dtype = ctx.shapes
        lay = s["lay"]
        N, d, Nk = lay.N, lay.d, lay.Nk
        bp = _lib.base_ptr
        dev = s["q_e"].device
        if spec.kind != "edit" and not spec.blend:
            raise NotImplementedError("gradient through the remover's identity branch is never requested by the reference loop "
                                      "(optimisation ends before obj_edit_step, editor.py:189)")
        g_out = None
        if d_out is not None:
            if d_out.dtype not in (torch.bfloat16, torch.float32):
                d_out = d_out.float()
            if not d_out.is_contiguous():
                d_out = d_out.contiguous()
            g_out = lay.sl(d_out, cb1)
        has_loss = s["g_loss"] is not None and d_terms is not None
        d_loss = d_terms[5:6].to(torch.float32).contiguous() if has_loss else None
        M = s["M"] if has_loss else 0
        has_extra = M > 0
        rowmap = ptr(spec.cache.rowmap) if has_extra else None
        extra = ptr(s["extra"]) if has_extra else None
        dq = torch.zeros(q_shape, device=dev, dtype=q_dtype)
        if g_out is None and not has_loss:
            return dq, None, None, None, None
        d_o = torch.empty(h, N, d, device=dev, dtype=torch.bfloat16)
        delta = torch.empty(h, N, device=dev, dtype=torch.float32)
        call("gd_attn_bwd_prep", bp(g_out), int(g_out is not None and g_out.dtype == torch.bfloat16), _lib.host_longs(lay.q), ptr(s["coef"]),
             ptr(s["g_loss"]) if has_loss else None, ptr(d_loss), ptr(s["o_e"]), ptr(s["delta_extra"]) if has_extra else None, rowmap, M,
             h, N, d, ptr(d_o), ptr(delta), stream())
        # dQ of the edit entry is written by the kernel straight into its slab of the full gradient tensor, in q's layout and dtype
        dq_is_bf16 = int(q_dtype == torch.bfloat16)
        ex_ld, ex_km = (s["ex_ld"], s["ex_key_major"]) if has_extra else (s["ld"], 0)
        if _lib.HAS_SM100 and Nk == N and N % 128 == 0 and d in (40, 80) and N >= 1024 and s["ld"] % 4 == 0:
            if has_extra and ex_km == 2:
                # flash-backward term by the tcgen05 kernel, removal term (inpaint rows only) by its own contraction on top of it
                call("gd_attn_bwd_sm100", bp(s["q_e"]), bp(s["k_e"]), bp(s["v_e"]), ptr(d_o), ptr(s["lse_e"]), ptr(delta), None, None, None, s["ld"], 0,
                     bp(lay.sl(dq, ce0)), h, N, d, float(spec.scale), lay.strides(), dq_is_bf16, 0, stream(), tag=(h, N, N, d))
                call("gd_removal_dq_rows", extra, bp(s["k_e"]), ptr(spec.cache.rows), ptr(d_loss), bp(lay.sl(dq, ce0)), h, M, N, Nk, d, float(spec.scale),
                     ex_ld, _lib.host_longs([lay.kv[0], lay.kv[1], lay.q[0], lay.q[1]]), dq_is_bf16, stream())
            else:
                call("gd_attn_bwd_sm100", bp(s["q_e"]), bp(s["k_e"]), bp(s["v_e"]), ptr(d_o), ptr(s["lse_e"]), ptr(delta), extra,
                     ptr(d_loss) if has_extra else None, rowmap, ex_ld, M, bp(lay.sl(dq, ce0)), h, N, d, float(spec.scale), lay.strides(), dq_is_bf16,
                     ex_km, stream(), tag=(h, N, N, d))
        else:
            assert ex_km == 0
            call("gd_attn_bwd", 0, bp(s["q_e"]), bp(s["k_e"]), bp(s["v_e"]), ptr(d_o), ptr(s["lse_e"]), ptr(delta), extra,
                 ptr(d_loss) if has_extra else None, rowmap, ex_ld, M, bp(lay.sl(dq, ce0)), h, N, Nk, d, float(spec.scale), lay.strides(),
                 dq_is_bf16, stream(), tag=(h, N, Nk, d))
        dk = None
        if spec.is_cross and spec.kind == "edit":
            dk = torch.zeros(k_shape, device=dev, dtype=k_dtype)
            # Nk = 77 -> two key tiles per head: split the query walk so the grid fills the 148 SMs (fixed-order partial sums)
            splits = max(1, min((N + 63) // 64, (2 * 148) // (h * ((Nk + 63) // 64))))
            ws = torch.empty(splits, h, Nk, d, device=dev, dtype=torch.float32) if splits > 1 else None
            call("gd_attn_bwd_dk_split", bp(s["q_e"]), bp(s["k_e"]), bp(s["v_e"]), ptr(d_o), ptr(s["lse_e"]), ptr(delta), extra,
                 ptr(d_loss) if has_extra else None, rowmap, s["ld"], M, bp(lay.sl(dk, ce0)), ptr(ws), splits, h, N, Nk, d, float(spec.scale),
                 lay.strides(out=lay.kv), int(k_dtype == torch.bfloat16), stream(), tag=(h, N, Nk, d))
        return dq, dk, None, None, None


def shared_attention_layer(q, k, v, spec):
    """-> (out, loss or None, terms or None).  q, k, v: heads-major (B*H, N, d) tensors (the reference's layout) or ProjViews; `out` comes
    back in the same layout (a ProjView input returns the plain (B, N, H*d) tensor).  Differentiable w.r.t. the edit sample's q (and k
    for cross layers) whenever autograd is recording; the loss exists only when spec.with_loss."""
    proj = isinstance(q, ProjView)
    if proj:
        q, k, v = q.t, k.t, v.t
    if spec.base_mode == "read":
        assert not spec.with_loss and not (torch.is_grad_enabled() and (q.requires_grad or k.requires_grad))
        return _forward_cached_impl(q, k, v, spec, proj), None, None
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad):
        out, terms = _SharedAttentionLayerFn.apply(q, k, v, spec, proj)
        if spec.with_loss:
            return out, terms[5], terms
        return out, None, None
    out, terms, _ = _forward_impl(q, k, v, spec, proj)
    return out, (terms[5] if terms is not None else None), terms


def attention_maps(q, k, scale, heads, entries=None):
    """softmax(scale q k^T) for (B*H, N, d) tensors or ProjViews, MATERIALISED: (len(entries)*H, N, Nk) fp32.  Only what the reference's
    AttentionStore keeps (maps of N <= 16^2 query tokens, attention_sharing.py:166-179) is ever asked for: the path itself never forms a map.
    Row sums come from a forward launch (LSE), the probabilities from gd_attn_probs."""
    proj = isinstance(q, ProjView)
    if proj:
        q, k = q.t, k.t
    lay = _Layout(q, k, heads, proj)
    B = q.shape[0] if proj else q.shape[0] // heads
    entries = list(range(B)) if entries is None else list(entries)
    qb, kb = _bf16(q), _bf16(k)
    N, Nk, d = lay.N, lay.Nk, lay.d
    ld = (Nk + 7) // 8 * 8
    qk_st = _lib.host_longs([lay.q[0], lay.q[1], lay.kv[0], lay.kv[1]])
    out = torch.empty(len(entries) * heads, N, Nk, device=q.device, dtype=torch.float32)
    for n, i in enumerate(entries):
        qi, ki = lay.sl(qb, i), lay.sl(kb, i)
        _, LSE = attention_forward([qi], [ki], [ki], scale, dims=(heads, N, Nk, d), strides=lay.strides())   # (values are irrelevant for the LSE)
        p = torch.empty(heads, N, ld, device=q.device, dtype=torch.bfloat16)
        call("gd_attn_probs", _lib.base_ptr(qi), _lib.base_ptr(ki), ptr(LSE[0]), None, N, heads, N, Nk, d, float(scale), ptr(p), ld, qk_st, stream())
        out[n * heads:(n + 1) * heads] = p[:, :, :Nk].float()
    return out


def plain_attention(q, k, v, scale, heads):
    """softmax(scale q k^T) v for (B*H, N, d) tensors or ProjViews: VanillaAttentionProcessor / outside the replace window
    (attention_processors.py:120-121, 646-647).  Forward only.  The kernel writes the result in q's dtype and layout."""
    proj = isinstance(q, ProjView)
    if proj:
        q, k, v = q.t, k.t, v.t
    lay = _Layout(q, k, heads, proj)
    B = q.shape[0] if proj else q.shape[0] // heads
    assert q.dtype in (torch.bfloat16, torch.float32)
    qb, kb, vb = _bf16(q), _bf16(k), _bf16(v)
    out = lay.new_batch(q, B, lay.N)
    for g0 in range(0, B, 8):
        g1 = min(B, g0 + 8)
        rng = range(g0, g1)
        attention_forward([lay.sl(qb, i) for i in rng], [lay.sl(kb, i) for i in rng], [lay.sl(vb, i) for i in rng], scale,
                          dims=(heads, lay.N, lay.Nk, lay.d), strides=lay.strides(), want32=[], os=[lay.sl(out, i) for i in rng],
                          os_is_bf16=q.dtype == torch.bfloat16)
    return out
