"""Fused warped shared-attention layer: the operator the controllers call once per UNet attention layer.

One call replaces, for one layer, the reference's sequence (attention_processors.py:633-664 with :513-624 / :384-508 / :748-928):
plain attention of the untouched batch entries, splat-warp of the base queries, attention of warped and edit queries against
the shared base K/V, the five attention-map loss terms, the output blend -- and, in backward, dQ (dK for cross layers) through
softmax and through the loss terms.  All arithmetic is in the C-ABI CUDA library (`_lib`); torch provides memory and autograd
bookkeeping only.  There is no CPU path.
"""
import math

import torch

from . import _lib, geometry
from ._lib import call, ptr, stream

TERM_NAMES = ("sim", "movement", "removal", "smoothness", "amodal")


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
        self.rows = rows.contiguous()
        self.M = int(rows.numel())
        rowmap = torch.full((self.N,), -1, device=dev, dtype=torch.int32)
        if self.M:
            rowmap[rows.long()] = torch.arange(self.M, device=dev, dtype=torch.int32)
        self.rowmap = rowmap
        self.sum_bg, self.sum_edit, self.sum_inp = float(self.m_bg.sum()), float(self.m_edit.sum()), float(self.m_inp.sum())
        self.idx = self.dist2 = None
        if coords_S is not None:
            idx, _, dist2 = geometry.splat_index(coords_S[None])
            self.idx, self.dist2 = keep("idx", idx), keep("dist2", dist2)
        self.knn_idx = self.knn_val = self.knn_w = None
        self.sum_w_am = 0.0
        if need_amodal:
            self.knn_idx = torch.empty(self.N, 4, device=dev, dtype=torch.int32)
            self.knn_val = torch.empty(self.N, 4, device=dev, dtype=torch.float32)
            self.knn_w = torch.empty(self.N, device=dev, dtype=torch.float32)
            call("gd_amodal_knn", ptr(self.m_edit), S, ptr(self.knn_idx), ptr(self.knn_val), ptr(self.knn_w), stream())
            self.sum_w_am = float((self.knn_w * self.m_am).sum())


def _bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


def attention_forward(qs, ks, vs, scale):
    """G query streams (H,N,d) bf16 against K/V (H,Nk,d) bf16 -> O (G,H,N,d) fp32, LSE (G,H,N) fp32."""
    G = len(qs)
    H, N, d = qs[0].shape
    Nk = ks[0].shape[1]
    O = torch.empty(G, H, N, d, device=qs[0].device, dtype=torch.float32)
    LSE = torch.empty(G, H, N, device=qs[0].device, dtype=torch.float32)
    entry = "gd_attn_fwd_generic"
    if _lib.HAS_SM100 and Nk == N and N % 128 == 0 and d in (40, 80) and N >= 1024:
        entry = "gd_attn_fwd_sm100"
    call(entry, _lib.ptr_array(qs), _lib.ptr_array(ks), _lib.ptr_array(vs), _lib.ptr_array([O[g] for g in range(G)]),
         _lib.ptr_array([LSE[g] for g in range(G)]), G, H, N, Nk, d, float(scale), stream(), tag=(G, H, N, Nk, d))
    return O, LSE


def warp_queries(q_base, cache):
    """q_base (H,N,d) bf16 -> q*(1-M) + M*splat(q)  (attention_processors.py:544), bf16"""
    return geometry.splat_composite(q_base, cache.idx, cache.dist2, channels_last=True, blend_mask=cache.m_edit, out_dtype=torch.bfloat16)


class LayerSpec:
    """Static description of one controller call."""
    __slots__ = ("kind", "is_cross", "heads", "cb", "ce", "scale", "blend", "with_loss", "weights", "cache", "log_accum", "w_rem_dev")
     # optional device scalar holding weights["removal"] (kept current by the controller; lets a captured pass follow the
                         # adaptive schedule)

    def __init__(self, **kw):
        self.w_rem_dev = None   # optional device scalar holding weights["removal"], kept current by the controller: lets a captured pass
        for k, v in kw.items():  # follow the adaptive schedule
            setattr(self, k, v)


def _forward_impl(q, k, v, spec):
    """Runs the fused layer.  Returns (out, terms6 or None, saved-for-backward dict or None)."""
    h, (cb0, cb1), (ce0, ce1) = spec.heads, spec.cb, spec.ce
    assert ce1 - ce0 == 1 and cb1 - cb0 == 1
    cache = spec.cache
    N, d = q.shape[1], q.shape[2]
    Nk = k.shape[1]
    qb, kb, vb = _bf16(q), _bf16(k), _bf16(v)
    sl = lambda t, i: t[i * h:(i + 1) * h]
    qs = [sl(qb, i) for i in range(cb1)]
    ks = [sl(kb, i) for i in range(cb1)]
    vs = [sl(vb, i) for i in range(cb1)]
    q_e = sl(qb, ce0)
    if spec.kind == "edit":
        q_w = warp_queries(sl(qb, cb0), cache)
        k_e = sl(kb, ce0) if spec.is_cross else sl(kb, cb0)
        qs += [q_w, q_e]
        ks += [sl(kb, cb0), k_e]
        vs += [sl(vb, cb0), sl(vb, cb0)]
        g_e = cb1 + 1
    else:
        k_e = sl(kb, cb0)
        qs += [q_e]
        ks += [k_e]
        vs += [sl(vb, cb0)]
        g_e = cb1
        if not spec.blend:
            qs += [q_e]
            ks += [sl(kb, ce0)]
            vs += [sl(vb, ce0)]
    O, LSE = attention_forward(qs, ks, vs, spec.scale)
    r = O[g_e]
    e = O[cb1] if spec.kind == "edit" else O[cb0]
    out = torch.empty(cb1 * h + h, N, d, device=q.device, dtype=q.dtype)
    out[:cb1 * h] = O[:cb1].reshape(cb1 * h, N, d)
    out_edit = out[cb1 * h:]
    is_bf16 = q.dtype == torch.bfloat16
    assert q.dtype in (torch.bfloat16, torch.float32)
    if spec.kind == "edit":
        if spec.blend:
            coef = cache.one_minus_m_edit
            call("gd_blend_rows", ptr(e), ptr(cache.m_edit), ptr(r), ptr(coef), h, N, d, ptr(out_edit), int(is_bf16), stream())
        else:
            coef = None
            call("gd_blend_rows", None, None, ptr(r), None, h, N, d, ptr(out_edit), int(is_bf16), stream())
    else:
        if spec.blend:
            coef = cache.m_inp_plus_bg
            call("gd_blend_rows", None, None, ptr(r), ptr(coef), h, N, d, ptr(out_edit), int(is_bf16), stream())
        else:
            coef = cache.m_bg
            call("gd_blend_rows", ptr(O[g_e + 1]), ptr(cache.m_inp), ptr(r), ptr(coef), h, N, d, ptr(out_edit), int(is_bf16), stream())
    saved = dict(q_e=q_e, k_e=k_e, v_e=sl(vb, cb0), o_e=r, lse_e=LSE[g_e], g_loss=None, extra=None, delta_extra=None, coef=coef,
                 ld=(Nk + 7) // 8 * 8, M=0)
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
        a_b = torch.empty(h, N, ld, device=dev, dtype=torch.bfloat16)
        call("gd_attn_probs", ptr(sl(qb, cb0)), ptr(sl(kb, cb0)), ptr(LSE[cb0]), None, N, h, N, Nk, d, float(spec.scale), ptr(a_b), ld, stream())
        a_e = torch.empty(h, M, ld, device=dev, dtype=torch.bfloat16)
        call("gd_attn_probs", ptr(q_e), ptr(k_e), ptr(LSE[g_e]), ptr(cache.rows), M, h, N, Nk, d, float(spec.scale), ptr(a_e), ld, stream())
        n_tiles = (N + 63) // 64
        partial = torch.empty(h, n_tiles, M, 4, device=dev, dtype=torch.float32)
        call("gd_corr_max_partial", ptr(a_e), ptr(a_b), h, M, N, Nk, ld, ptr(cache.m_inp), ptr(cache.m_bg), ptr(partial), stream())
        rem_terms = torch.empty(h * M, device=dev, dtype=torch.float32)
        g2 = torch.empty(h * M, 2, device=dev, dtype=torch.float32)
        j2 = torch.empty(h * M, 2, device=dev, dtype=torch.int32)
        delta_extra = torch.empty(h, M, device=dev, dtype=torch.float32)
        extra = torch.empty(h, M, ld, device=dev, dtype=torch.float32)
        wdev = spec.w_rem_dev
        call("gd_removal_finalize", ptr(partial), n_tiles, h, M, S, ptr(cache.rows), ptr(cache.m_inp), ptr(cache.m_bg),
             inv_rem if wdev is not None else w_rem * inv_rem, ptr(wdev), ptr(a_b), N, Nk, ld, ptr(rem_terms), ptr(g2), ptr(j2), ptr(delta_extra),
             ptr(extra), stream())
        del a_b, a_e, partial
    terms = torch.empty(6, device=dev, dtype=torch.float32)
    call("gd_loss_reduce", ptr(partials), n_part, ptr(rem_terms), h * M if M > 0 else 0,
         _lib.host_f32([inv_sim, inv_mov, inv_amo, inv_smh, inv_smh, inv_rem]), _lib.host_f32([w_sim, w_mov, w_amo, w_sm, w_rem]),
         ptr(spec.w_rem_dev), 1.0 if use_amodal else 0.0, ptr(terms), ptr(spec.log_accum), stream())
    saved.update(g_loss=g_loss, extra=extra, delta_extra=delta_extra, M=M)
    return out, terms, saved


class _SharedAttentionLayerFn(torch.autograd.Function):
    """(q, k, v) -> (out, terms[6]); terms[5] is the weighted layer loss (differentiable), terms[0:5] the logged terms."""

    @staticmethod
    def forward(ctx, q, k, v, spec):
        out, terms, saved = _forward_impl(q, k, v, spec)
        ctx.spec, ctx.saved_dict = spec, saved
        ctx.shapes = (q.shape, k.shape, q.dtype, k.dtype)
        if terms is None:
            terms = torch.zeros(6, device=q.device, dtype=torch.float32)
            ctx.mark_non_differentiable(terms)
        return out, terms

    @staticmethod
    def backward(ctx, d_out, d_terms):
        spec, s = ctx.spec, ctx.saved_dict
        h, (cb0, cb1), (ce0, ce1) = spec.heads, spec.cb, spec.ce
        q_shape, k_shape, q_dtype, k_dtype = ctx.shapes
        N, d = q_shape[1], q_shape[2]
        Nk = k_shape[1]
        dev = s["q_e"].device
        if spec.kind != "edit" and not spec.blend:
            raise NotImplementedError("gradient through the remover's identity branch is never requested by the reference loop "
                                      "(optimisation ends before obj_edit_step, editor.py:189)")
        g_out = None
        if d_out is not None:
            g_out = d_out[cb1 * h:].contiguous()
            if g_out.dtype not in (torch.bfloat16, torch.float32):
                g_out = g_out.float()
        has_loss = s["g_loss"] is not None and d_terms is not None
        d_loss = d_terms[5:6].to(torch.float32).contiguous() if has_loss else None
        M = s["M"] if has_loss else 0
        has_extra = M > 0
        rowmap = ptr(spec.cache.rowmap) if has_extra else None
        extra = ptr(s["extra"]) if has_extra else None
        dq = torch.zeros(q_shape, device=dev, dtype=q_dtype)
        if g_out is None and not has_loss:
            return dq, None, None, None
        d_o = torch.empty(h, N, d, device=dev, dtype=torch.bfloat16)
        delta = torch.empty(h, N, device=dev, dtype=torch.float32)
        call("gd_attn_bwd_prep", ptr(g_out), int(g_out is not None and g_out.dtype == torch.bfloat16), ptr(s["coef"]),
             ptr(s["g_loss"]) if has_loss else None, ptr(d_loss), ptr(s["o_e"]), ptr(s["delta_extra"]) if has_extra else None, rowmap, M,
             h, N, d, ptr(d_o), ptr(delta), stream())
        dq_e = torch.empty(h, N, d, device=dev, dtype=torch.float32)
        if _lib.HAS_SM100 and Nk == N and N % 128 == 0 and d in (40, 80) and N >= 1024 and s["ld"] % 4 == 0:
            call("gd_attn_bwd_sm100", ptr(s["q_e"]), ptr(s["k_e"]), ptr(s["v_e"]), ptr(d_o), ptr(s["lse_e"]), ptr(delta), extra,
                 ptr(d_loss) if has_extra else None, rowmap, s["ld"], M, ptr(dq_e), h, N, d, float(spec.scale), stream(), tag=(h, N, d))
        else:
            call("gd_attn_bwd", 0, ptr(s["q_e"]), ptr(s["k_e"]), ptr(s["v_e"]), ptr(d_o), ptr(s["lse_e"]), ptr(delta), extra,
                 ptr(d_loss) if has_extra else None, rowmap, s["ld"], M, ptr(dq_e), h, N, Nk, d, float(spec.scale), stream())
        dq[ce0 * h:ce1 * h] = dq_e.to(q_dtype)
        dk = None
        if spec.is_cross and spec.kind == "edit":
            dk_e = torch.empty(h, Nk, d, device=dev, dtype=torch.float32)
            # Nk = 77 -> two key tiles per head: split the query walk so the grid fills the 148 SMs (fixed-order partial sums)
            splits = max(1, min((N + 63) // 64, (2 * 148) // (h * ((Nk + 63) // 64))))
            ws = torch.empty(splits, h, Nk, d, device=dev, dtype=torch.float32) if splits > 1 else None
            call("gd_attn_bwd_dk_split", ptr(s["q_e"]), ptr(s["k_e"]), ptr(s["v_e"]), ptr(d_o), ptr(s["lse_e"]), ptr(delta), extra,
                 ptr(d_loss) if has_extra else None, rowmap, s["ld"], M, ptr(dk_e), ptr(ws), splits, h, N, Nk, d, float(spec.scale), stream())
            dk = torch.zeros(k_shape, device=dev, dtype=k_dtype)
            dk[ce0 * h:ce1 * h] = dk_e.to(k_dtype)
        return dq, dk, None, None


def shared_attention_layer(q, k, v, spec):
    """-> (out, loss or None, terms or None).  Differentiable w.r.t. the edit sample's q (and k for cross layers) whenever autograd
    is recording; the loss exists only when spec.with_loss."""
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad):
        out, terms = _SharedAttentionLayerFn.apply(q, k, v, spec)
        if spec.with_loss:
            return out, terms[5], terms
        return out, None, None
    out, terms, _ = _forward_impl(q, k, v, spec)
    return out, (terms[5] if terms is not None else None), terms


def plain_attention(q, k, v, scale, heads):
    """softmax(scale q k^T) v for (B*H, N, d) tensors: VanillaAttentionProcessor / outside the replace window
    (attention_processors.py:120-121, 646-647).  Forward only."""
    BH, N, d = q.shape
    B = BH // heads
    qb, kb, vb = _bf16(q), _bf16(k), _bf16(v)
    outs = []
    for g0 in range(0, B, 8):
        g1 = min(B, g0 + 8)
        sl = lambda t, i: t[i * heads:(i + 1) * heads]
        O, _ = attention_forward([sl(qb, i) for i in range(g0, g1)], [sl(kb, i) for i in range(g0, g1)], [sl(vb, i) for i in range(g0, g1)], scale)
        outs.append(O.reshape((g1 - g0) * heads, N, d))
    O = outs[0] if len(outs) == 1 else torch.cat(outs)
    return O.to(q.dtype)
