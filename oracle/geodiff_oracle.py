"""oracle/geodiff_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (fp32, torch-on-CPU for the floating-point contractions, C for the integer /
bit-exact geometry in oracle/geom_cpu.c + oracle/pt3d_cpu.c) of GeoDiffuser's geometry-warped
shared-attention hot path.  Every function cites the reference lines it follows (paths relative
to /root/reference/GeoDiffuser/utils/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker / the timed CPU baseline.  The product package
`geodiffuser_b200` never imports it.

Pinning: oracle/make_golden.py runs the *real* reference modules (imported from /root/reference
with the stubs of oracle/ref_import.py) on the same seeded inputs and asserts equality with this
restatement before writing tests/golden/*.npz.  The pytorch3d boundary (rasterize_points /
alpha_composite / rasterize_meshes) has no source under /root/reference: PARITY UNPINNED there
(see pt3d_cpu.c).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import pt3d_shim

FOCAL = 550.0  # vis_utils.py:404
SPLAT_RADIUS, SPLAT_K, SPLAT_TAU = 1.3, 15, 1.0  # warp_utils.py:50-58 (effective for the whole run)


# --------------------------------------------------------------------------------------------
# A1  correspondence field
# --------------------------------------------------------------------------------------------
def normalise_depth(depth: np.ndarray) -> np.ndarray:
    """vis_utils.py:408-418 (float64 numpy, as in the reference)."""
    depth = np.array(depth, dtype=np.float64)
    if np.sum(depth) == 0.5 * (depth.shape[0] * depth.shape[1]):
        return np.ones_like(depth) * 0.5
    depth = depth / (depth.max() + 1e-8)
    depth[depth > 0.95] = 1.0
    return depth


def camera_K(h, w, focal=FOCAL):
    """vis_utils.py:79-88,406: K = [[f,0,w/2],[0,f,h/2],[0,0,1]]"""
    return np.array([[focal, 0, w / 2.0], [0, focal, h / 2.0], [0, 0, 1]], dtype=np.float64)


def centred_transform(T: torch.Tensor, centre: torch.Tensor) -> torch.Tensor:
    """warp_utils.py:431-437: T' = C^-1 @ T @ C with C = translate(-centre); fp32 torch ops on the host,
    exactly as the reference spells them (4x4 LAPACK inverse + two 4x4 matmuls)."""
    C = torch.eye(4, dtype=torch.float32)
    C[:3, 3] += -centre
    C = C[None]
    return (C.inverse() @ T[None].float() @ C).float()[0]


def corr_build(depth, obj_mask, T, focal=FOCAL):
    """vis_utils.py:404-479 -> warp_utils.py:407-444.  Returns dict(coords (H,W,3) f32, cam (3,H,W),
    centre (3,), Tc (4,4), mask (H,W) f32, depth (H,W) f32, valid (H,W) bool)."""
    L = pt3d_shim.lib()
    import ctypes

    fp = ctypes.POINTER(ctypes.c_float)
    d64 = normalise_depth(depth)
    m = (d64 < 0.95) * 1.0
    if obj_mask is not None:
        m = np.asarray(obj_mask, dtype=np.float64) * m
    mask = ((torch.tensor(m)[None, None] >= 0.5) * 1.0)[0, 0].float().numpy().copy()  # vis_utils.py:425
    H, W = d64.shape
    d32 = torch.from_numpy(d64).float().numpy().copy()  # depth.float()  warp_utils.py:410
    K = torch.from_numpy(camera_K(H, W, focal))[None].float()  # .type_as(depth)
    Kinv = K.inverse()[0].contiguous().numpy()
    Kf = K[0].contiguous().numpy()
    cam = np.empty((3, H, W), np.float32)
    L.geo_pixel2cam.argtypes = [fp, ctypes.c_int, ctypes.c_int, fp, fp]
    L.geo_pixel2cam(d32.ctypes.data_as(fp), H, W, Kinv.ctypes.data_as(fp), cam.ctypes.data_as(fp))
    centre = np.empty(3, np.float32)
    L.geo_centroid.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, fp]
    L.geo_centroid.restype = ctypes.c_long
    L.geo_centroid(cam.ctypes.data_as(fp), mask.ctypes.data_as(fp), H, W, centre.ctypes.data_as(fp))
    Tc = centred_transform(T, torch.from_numpy(centre))
    Rt = Tc[:3, :].contiguous().numpy()
    coords = np.empty((H, W, 3), np.float32)
    L.geo_project.argtypes = [fp, ctypes.c_int, ctypes.c_int, fp, fp, fp]
    L.geo_project(cam.ctypes.data_as(fp), H, W, Rt.ctypes.data_as(fp), Kf.ctypes.data_as(fp), coords.ctypes.data_as(fp))
    valid = np.abs(coords[..., :2]).max(-1) <= 1  # warp_utils.py:473
    return dict(coords=coords, cam=cam, centre=centre, Tc=Tc.numpy(), Kinv=Kinv, K=Kf, mask=mask, depth=d32,
                valid=valid)


# --------------------------------------------------------------------------------------------
# A4  per-resolution resize of coords / masks
# --------------------------------------------------------------------------------------------
def resize_bilinear(x: np.ndarray, S: int) -> np.ndarray:
    """(C,Hin,Win) -> (C,S,S).  generic_torch.py:185,205 (T.Resize BILINEAR antialias=False)."""
    import ctypes

    L = pt3d_shim.lib()
    fp = ctypes.POINTER(ctypes.c_float)
    x = np.ascontiguousarray(x, dtype=np.float32)
    C, Hin, Win = x.shape
    out = np.empty((C, S, S), np.float32)
    L.geo_resize_bilinear.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
    L.geo_resize_bilinear(x.ctypes.data_as(fp), C, Hin, Win, S, S, out.ctypes.data_as(fp))
    return out


def resize_coords(coords: np.ndarray, S: int) -> np.ndarray:
    """(H,W,3) -> (S,S,3).  generic_torch.py:156-186."""
    return np.ascontiguousarray(resize_bilinear(np.transpose(coords, (2, 0, 1)), S).transpose(1, 2, 0))


def binarize(t, thresh=0.5):
    """generic_torch.py:122"""
    return ((t > thresh) * 1.0).astype(np.float32) if isinstance(t, np.ndarray) else (t > thresh) * 1.0


def build_masks(image_mask, mask_new_warped, amodal_mask, S):
    """attention_processors.py:338-360.  image_mask (512,512) {0,1}; mask_new_warped (512,512) binarised
    warped mask; amodal_mask (512,512).  Returns dict of (S,S) fp32 arrays."""
    m_src = resize_bilinear(binarize(np.asarray(image_mask, np.float32))[None], S)[0]  # mask_warp
    m_warp = resize_bilinear(np.asarray(mask_new_warped, np.float32)[None], S)[0]  # mask_new_warped (soft)
    am = resize_bilinear(np.asarray(amodal_mask, np.float32)[None], S)[0]
    m_amodal = binarize(am - m_warp)
    m_inter = binarize((m_warp + m_amodal) * m_src)
    m_inpaint = binarize(m_src - m_inter)  # mask_1_empty
    m_bg = binarize(np.ones_like(m_warp) - (m_inpaint + m_warp))  # mask_wo_edit
    return dict(mask_new_warped=m_warp, mask_warp=m_src, amodal_mask=m_amodal, mask_intersection=m_inter,
                mask_1_empty=m_inpaint, mask_wo_edit=m_bg)


# --------------------------------------------------------------------------------------------
# A3  forward splat (pytorch3d point rasteriser + alpha compositor)
# --------------------------------------------------------------------------------------------
def splat_radius_ndc(S, radius_px=SPLAT_RADIUS):
    """warp_utils.py:94 (python double)"""
    return float(radius_px) / float(S) * 2.0


def splat_index(coords, radius_px=SPLAT_RADIUS, K=SPLAT_K):
    """coords (B,S,S,3) fp32 -> idx int32 (B,S,S,K) [packed index b*S*S+p], zbuf, dist2.
    warp_utils.py:80-113 (x,y negated :90-91)."""
    coords = np.asarray(coords, np.float32)
    B, S = coords.shape[0], coords.shape[1]
    pts = coords.reshape(B, S * S, 3).copy()
    pts[:, :, 0] = -pts[:, :, 0]
    pts[:, :, 1] = -pts[:, :, 1]
    return pt3d_shim.rasterize_points_np(pts, S, splat_radius_ndc(S, radius_px), K)


def splat_alpha(dist2, S, radius_px=SPLAT_RADIUS, tau=SPLAT_TAU):
    """warp_utils.py:131-140: alpha = (1 - clamp(d2 / r^2, 1e-3, 1)^0.5)^tau ; r^2 in python double,
    divided as fp32 scalar."""
    r2 = np.float32(pow(splat_radius_ndc(S, radius_px), 2))
    a = np.float32(1.0) - np.sqrt(np.clip(dist2 / r2, np.float32(1e-3), np.float32(1.0)))
    if tau != 1.0:
        a = np.power(a, np.float32(tau))
    return a.astype(np.float32)


def splat_composite(src, idx, dist2, radius_px=SPLAT_RADIUS, tau=SPLAT_TAU):
    """src (B,C,S,S) -> (B,C,S,S) fp32 holding fp16-rounded values (warp_utils.py:155-176)."""
    src = np.asarray(src, np.float32)
    B, C, S, _ = src.shape
    alpha = splat_alpha(dist2, S, radius_px, tau)
    feat = np.ascontiguousarray(src.reshape(B, C, S * S).transpose(1, 0, 2).reshape(C, B * S * S))
    out = pt3d_shim.alpha_composite_np(np.ascontiguousarray(idx.transpose(0, 3, 1, 2)),
                                       np.ascontiguousarray(alpha.transpose(0, 3, 1, 2)), feat)
    return out.astype(np.float16).astype(np.float32)  # .to(torch.half)


def warp_grid_edit(src, coords, radius_px=SPLAT_RADIUS, K=SPLAT_K, tau=SPLAT_TAU):
    """warp_utils.py:798-837.  src (B,C,S,S), coords (B,S,S,3)."""
    idx, _, d2 = splat_index(coords, radius_px, K)
    return splat_composite(src, idx, d2, radius_px, tau)


def mesh_mask(coords, mask):
    """A2: warp_utils.py:364-399 + 235-298.  coords (H,W,3), mask (H,W) -> (H,W) coverage in {0,1}.
    Triangles (tl,tr,bl) and (bl,tr,br) for every 2x2 quad fully inside the mask."""
    H, W = mask.shape
    inside = mask >= 0.5
    ids = -np.ones((H, W), np.int64)
    ids[inside] = np.arange(int(inside.sum()))
    verts = coords[inside].astype(np.float32).copy()
    verts[:, :2] = -verts[:, :2]
    tl, tr, bl, br = ids[:-1, :-1], ids[:-1, 1:], ids[1:, :-1], ids[1:, 1:]
    f1 = np.stack([tl, tr, bl], 0).reshape(3, -1)
    f2 = np.stack([bl, tr, br], 0).reshape(3, -1)
    faces = np.concatenate([f1, f2], -1)
    faces = faces[:, faces.min(0) > -1].T
    return pt3d_shim.mesh_coverage_np(verts, faces.astype(np.int32), H, float(1e-6) / float(2 * H))


def erode3(a):
    """generic_torch.py:210-221 (3x3 all-ones conv == 9)"""
    t = torch.from_numpy(np.asarray(a, np.float32))[None, None]
    k = torch.ones(1, 1, 3, 3)
    return ((F.conv2d(t, k, padding=1) == 9.0) * 1.0)[0, 0].numpy().astype(np.float32)


def dilate(a, kernel=3):
    """generic_torch.py:223-235"""
    t = torch.from_numpy(np.asarray(a, np.float32))[None, None]
    k = torch.ones(1, 1, kernel, kernel)
    return ((F.conv2d(t, k, padding=kernel // 2) >= 1) * 1.0)[0, 0].numpy().astype(np.float32)


# --------------------------------------------------------------------------------------------
# A6-A10  attention + loss math (fp32 torch on CPU)
# --------------------------------------------------------------------------------------------
def attention(q, k, v, scale):
    """attention_sharing.py:30-47 + torch.bmm at the call sites.  q (H,N,d), k/v (H,Nk,d).
    Returns (P, O).  The fg/bg score masking is a no-op in the reference."""
    s = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype), q, k.transpose(1, 2),
                      beta=0, alpha=scale)
    p = F.softmax(s, dim=-1)
    return p, torch.bmm(p, v)


def distance_grid(S):
    """generic_torch.py:132-140"""
    grid = F.affine_grid(torch.eye(3)[:2][None], (1, 1, S, S), align_corners=None)
    d = grid.reshape(1, -1, 2)
    return torch.sqrt(torch.sum(torch.square(d[:, :, None] - d[:, None]), -1) + 1e-12)


def gaussian_kernel5():
    """generic_torch.py:27-54 with kernel_size=5, sigma=5//2*2/6"""
    size, std = 5, (5 // 2 * 2 / 6.0)
    g = torch.arange(size, dtype=torch.float32)
    mean = (size - 1) / 2
    k1 = 1 / (std * math.sqrt(2 * math.pi)) * torch.exp(-(((g - mean) / (2 * std)) ** 2))
    k = k1[:, None] * k1[None, :]
    return k / k.sum()


def loss_sim(e, r, m_bg, eps=1e-8):
    """attention_processors.py:231-246.  e,r (1,H,N,d); m (1,1,N,1)"""
    return torch.sum(torch.sum(torch.abs(e - r), -1)[..., None] * m_bg) / (torch.sum(m_bg.expand_as(r)) + eps)


def loss_move(e, r, m_edit, eps=1e-8):
    """attention_processors.py:283-287"""
    return torch.sum(torch.abs(e - r) * m_edit) / (torch.sum(m_edit.expand_as(r)) + eps)


def loss_removal(A_e, A_b, m_inp, m_bg, dgrid, H):
    """attention_processors.py:248-280.  A_e (H,N,Nk) with grad, A_b (H,Nb,Nk) no grad; m (1,1,N,1);
    dgrid (1,N,N)."""
    rows = m_inp[0, 0, :, 0] > 0.5
    corr = torch.bmm(A_e[:, rows], A_b.transpose(1, 2))
    c_in = corr * m_inp[..., 0]
    c_bg = corr * m_bg[..., 0]
    mi, mb = torch.max(c_in, -1), torch.max(c_bg, -1)
    d_bg = dgrid[:, rows, mb.indices]
    w = torch.exp(-d_bg)
    return torch.sum(w * (-torch.log(mb.values + 1e-4) + torch.log(mi.values + 1e-4))) / (torch.sum(m_inp) * H + 1e-8), \
        dict(p_in=mi.values, p_bg=mb.values, j_in=mi.indices, j_bg=mb.indices, w=w, rows=rows)


def amodal_interp(e, m_edit, dgrid):
    """attention_sharing.py:68-105.  e (1,H,N,d) -> (interp (1,H,N,d), weights (1,H,N))."""
    fg = (m_edit[:1, :1, :, 0] > 0.5) * 1.0
    dist = dgrid * 512 / 2.0 + 100000 * (1.0 - fg)
    inv = 1.0 / (dist + 1e-4)
    tk = torch.topk(inv, k=4, dim=-1, largest=True, sorted=False)
    idx = tk.indices[0]  # (N,4)
    val = tk.values[0]
    sel = e[:, :, idx]  # (1,H,N,4,d)
    interp = torch.sum(sel * val[None, None, :, :, None], -2) / (torch.sum(val, -1)[None, None, :, None] + 1e-12)
    w = torch.exp(-(1 / torch.max(val, -1).values) / 5)
    return interp, w[None, None].expand(e.shape[0], e.shape[1], -1), idx, val


def smooth5(x):
    """generic_torch.py:145-154.  x (1,H,N,d)"""
    b, h, n, D = x.shape
    S = int(np.sqrt(n))
    xi = x.permute(0, 1, 3, 2).reshape(-1, 1, S, S)
    out = F.conv2d(xi, gaussian_kernel5()[None, None], padding=2)
    return out.reshape(b, h, D, n).permute(0, 1, 3, 2)


def amodal_target(e, m_edit, dgrid):
    """attention_processors.py:291-293"""
    interp, w, _, _ = amodal_interp(e, m_edit, dgrid)
    fg = m_edit[0, 0, :, 0] > 0.5
    interp = interp.clone()
    interp[:, :, fg] = e[:, :, fg]
    return smooth5(interp), w


def loss_amodal(e, r, m_edit, dgrid, m_am, eps=1e-8):
    """attention_processors.py:289-305"""
    tgt, w = amodal_target(e, m_edit, dgrid)
    return torch.sum(torch.abs(tgt - r) * w[..., None] * m_am) / (torch.sum(w[..., None] * m_am.expand_as(r)) + eps)


def loss_smooth(r):
    """loss.py:22-41.  r (1,H,N,d)"""
    b, f, hw, d = r.shape
    S = int(np.sqrt(hw))
    x = r.reshape(b, f, S, S, d)
    return (x[:, :, 1:, :] - x[:, :, :-1, :]).abs().mean() + (x[:, :, :, 1:] - x[:, :, :, :-1]).abs().mean()


EDIT_WEIGHTS = {"self": {"sim": 110, "movement": 13.5, "removal": 1.67, "smoothness": 35.0, "amodal": 80.5},
                "cross": {"sim": 60, "movement": 6.34, "removal": 1.6, "smoothness": 20.0, "amodal": 3.5}}
REMOVER_WEIGHTS = {"self": {"sim": 110.0, "removal": 3.6, "smoothness": 35.0},
                   "cross": {"sim": 60.0, "removal": 3.6, "smoothness": 20.0}}


def edit_layer(q, k, v, is_cross, scale, heads, coords_base, coords_edit, masks, coords_S, use_cfg, blend,
               weights=None):
    """One AttentionGeometryEdit layer inside the replace window: attention_processors.py:633-664 with
    :513-624 (self) / :384-508 (cross).
    q,k,v (B'*H, N|Nk, d) fp32 (q, k may require grad); masks: dict of (S,S) arrays from build_masks;
    coords_S (S,S,3).  blend: cur_step < int(num_steps*obj_edit_step).
    Returns dict(out, terms{...} or None, loss or None, edit_out, replace_out, A_e, q_w)."""
    h = heads
    cb0, cb1 = coords_base
    ce0, ce1 = coords_edit
    N, d = q.shape[1], q.shape[2]
    S = int(round(math.sqrt(N)))
    P_plain, O_plain = attention(q[: cb1 * h], k[: cb1 * h], v[: cb1 * h], scale)
    qb, kb, vb = (t[cb0 * h: cb1 * h].detach() for t in (q, k, v))
    qe, ke = q[ce0 * h: ce1 * h], k[ce0 * h: ce1 * h]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    m_warp = t(masks["mask_new_warped"])
    with torch.no_grad():
        q_img = qb.permute(0, 2, 1).reshape(h, d, S, S)
        warped = t(warp_grid_edit(q_img.numpy(), np.broadcast_to(coords_S[None], (h, S, S, 3))))
        q_w = (q_img * (1.0 - m_warp) + m_warp * warped).reshape(h, d, N).permute(0, 2, 1)
        _, edit_out = attention(q_w, kb, vb, scale)
        e = edit_out[None]
    if is_cross:
        A_e, rep = attention(qe, ke, vb, scale)  # :432-433 (own text keys, base values)
    else:
        A_e, rep = attention(qe, kb, vb, scale)  # :555-557
    r = rep[None]
    flat = lambda a: t(a).reshape(-1)[None, None, :, None]
    m_inp, m_bg, m_edit, m_am = flat(masks["mask_1_empty"]), flat(masks["mask_wo_edit"]), flat(
        masks["mask_new_warped"]), flat(masks["amodal_mask"])
    terms, loss = None, None
    if N >= 32 ** 2 and not use_cfg:
        dg = distance_grid(S)
        A_b = P_plain[cb0 * h: cb1 * h].detach()
        rem, _ = loss_removal(A_e, A_b, m_inp, m_bg, dg, h)
        sim = loss_sim(e, r, m_bg)
        mov = loss_move(e, r, m_edit)
        amo = loss_amodal(e, r, m_edit, dg, m_am)
        if N <= 32 ** 2:
            amo = 0.0 * mov
        smo = loss_smooth(r)
        lw = (weights or EDIT_WEIGHTS)["cross" if is_cross else "self"]
        loss = lw["sim"] * sim + lw["movement"] * mov + lw["removal"] * rem + lw["smoothness"] * smo + lw["amodal"] * amo
        terms = dict(sim=sim, movement=mov, removal=rem, smoothness=smo, amodal=amo)
    out_edit = e.detach() * m_edit + r * (1.0 - m_edit) if blend else r
    out = torch.cat([O_plain[: cb1 * h], out_edit[0]])
    return dict(out=out, terms=terms, loss=loss, edit_out=edit_out, replace_out=rep, A_e=A_e, q_w=q_w,
                O_plain=O_plain)


def remover_layer(q, k, v, is_cross, scale, heads, coords_base, coords_edit, image_mask_dilated, use_cfg, blend,
                  weights=None):
    """One AttentionGeometryRemover layer: attention_processors.py:931-959 with :842-928 / :748-837.
    image_mask_dilated (512,512): torch_dilate(mask,5) from the ctor (:986)."""
    h = heads
    cb0, cb1 = coords_base
    ce0, ce1 = coords_edit
    N = q.shape[1]
    S = int(round(math.sqrt(N)))
    P_plain, O_plain = attention(q[: cb1 * h], k[: cb1 * h], v[: cb1 * h], scale)
    kb, vb = k[cb0 * h: cb1 * h].detach(), v[cb0 * h: cb1 * h].detach()
    qe, ke, ve = q[ce0 * h: ce1 * h], k[ce0 * h: ce1 * h], v[ce0 * h: ce1 * h]
    m_src = resize_bilinear(binarize(np.asarray(image_mask_dilated, np.float32))[None], S)[0]
    m_inp_a = binarize(m_src)
    m_bg_a = binarize(np.ones_like(m_src) - m_inp_a)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    flat = lambda a: t(a).reshape(-1)[None, None, :, None]
    m_inp, m_bg = flat(m_inp_a), flat(m_bg_a)
    A_b = P_plain[cb0 * h: cb1 * h].detach()
    e = O_plain[cb0 * h: cb1 * h][None].detach()
    A_e, rep = attention(qe, kb, vb, scale)  # note: v_base NOT detached at :883 but base has no grad path
    r = rep[None]
    ident = None
    if not blend:
        _, ident = attention(qe, ke, ve, scale)
        ident = ident[None]
    terms, loss = None, None
    if N >= 32 ** 2 and not use_cfg:
        dg = distance_grid(S)
        sim = loss_sim(e, r, m_bg)
        rem, _ = loss_removal(A_e, A_b, m_inp, m_bg, dg, h)
        smo = loss_smooth(r)
        lw = (weights or REMOVER_WEIGHTS)["cross" if is_cross else "self"]
        loss = (lw["sim"] * sim + lw["removal"] * rem) + lw["smoothness"] * smo
        terms = dict(sim=sim, removal=rem, smoothness=smo)
    out_edit = r * m_inp + r * m_bg if blend else ident * m_inp + r * m_bg
    out = torch.cat([O_plain[: cb1 * h], out_edit[0]])
    return dict(out=out, terms=terms, loss=loss, replace_out=rep, A_e=A_e, O_plain=O_plain)


# --------------------------------------------------------------------------------------------
# A15-A17  DDIM step and latent update
# --------------------------------------------------------------------------------------------
def ddim_alphas(num_train=1000, beta_start=0.00085, beta_end=0.012):
    """diffusion.py:110 DDIMScheduler(scaled_linear, clip_sample=False, set_alpha_to_one=False)"""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(num_steps=50, num_train=1000):
    """diffusers 0.25 DDIMScheduler.set_timesteps, 'leading' spacing, steps_offset=0 -> 980, 960, ..., 0"""
    ratio = num_train // num_steps
    return (np.arange(0, num_steps) * ratio).round()[::-1].copy().astype(np.int64)


def ddim_step(x, eps, t, alphas, num_steps=50, num_train=1000):
    """diffusion.py:55 scheduler.step(eta=0) (formula restated in-tree at inversion.py:47-55)"""
    prev_t = t - num_train // num_steps
    a_t = alphas[t]
    a_prev = alphas[prev_t] if prev_t >= 0 else alphas[0]  # set_alpha_to_one=False -> final_alpha_cumprod = alphas[0]
    x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps


def cfg_combine(eps_u, eps_c, g):
    """diffusion.py:46"""
    return eps_u + g * (eps_c - eps_u)


def update_latent(latents, grad_lat, step, mask512, context=None, grad_ctx=None):
    """optimization.py:213-253 (optimizer=None branch).  latents (2,4,64,64); mask512 (512,512) = mask_new_warped[0,0]"""
    g = torch.nan_to_num(grad_lat, posinf=0.0, neginf=0.0, nan=0.0)
    m = torch.from_numpy(resize_bilinear(np.asarray(mask512, np.float32)[None], latents.shape[-1]))[None]
    last = latents[-1:] - 2.0 * m * step * g[-1:]
    last = last - (1.0 - m) * step * g[-1:]
    new_lat = torch.cat([latents[:-1], last], 0)
    new_ctx = None
    if context is not None:
        gc = torch.nan_to_num(grad_ctx, posinf=0.0, neginf=0.0, nan=0.0)
        new_ctx = torch.cat([context[:-1], context[-1:] - step * gc[-1:]], 0)
    return new_lat, new_ctx


def norm_tensor(a, eps=1e-12):
    """generic_torch.py:87"""
    return torch.sqrt(torch.sum(a * a) + eps)
