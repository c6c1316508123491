"""Host-side mirror of the reference's geometry helpers for the hot path, backed by the sm_100a kernels in
csrc/geometry.cu (no torch / pytorch3d arithmetic on the data path; torch only allocates device memory).

Mirrors (same names, argument meaning and return conventions):
  vis_utils.get_transform_coordinates        /root/reference/GeoDiffuser/utils/vis_utils.py:404-479
  warp_utils.warp_grid_edit                  warp_utils.py:798-837 (RasterizePointsXYsBlending.forward :72-176)
  generic_torch.reshape_transform_coords     generic_torch.py:156-186
  generic_torch.reshape_attention_mask       generic_torch.py:189-207
  generic_torch.binarize_tensor / torch_erode / torch_dilate   generic_torch.py:122, 210-235
  attention_processors.process_and_cache_masks (the mask algebra)  attention_processors.py:338-360
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream

FOCAL_LENGTH = 550.0  # vis_utils.py:404


class SplatSettings:
    """Module-level splat parameters, as the reference's shared `SPLATTER` object (warp_utils.py:50-58,179)."""
    radius = 1.3
    points_per_pixel = 15
    tau = 1.0


SPLATTER = SplatSettings()


def _dev(device=None):
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def _f32(t, device=None):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    return t.to(device=_dev(device) if not t.is_cuda else t.device, dtype=torch.float32).contiguous()


def camera_matrix(fx, fy, cx, cy):
    """vis_utils.py:79-88"""
    return np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=np.float64)


def binarize_tensor(t, thresh=0.5):
    return (t > thresh) * 1.0


def normalise_depth(depth):
    """vis_utils.py:408-418 (numpy float64 on the host, exactly as the reference)"""
    depth = np.array(depth, dtype=np.float64)
    if np.sum(depth) == 0.5 * (depth.shape[0] * depth.shape[1]):
        return np.ones_like(depth) * 0.5
    depth = depth / (depth.max() + 1e-8)
    depth[depth > 0.95] = 1.0
    return depth


def stage_depth_mask(depth, obj_mask, device=None):
    """Host part of A1 (vis_utils.py:408-425: float64 numpy depth normalisation and mask threshold, as the reference) followed by
    the host->device copy of the two (H,W) fp32 planes.  Returns (depth_dev, mask_dev)."""
    dev = _dev(device)
    d64 = normalise_depth(depth)
    m = (d64 < 0.95) * 1.0
    if obj_mask is not None:
        m = np.asarray(obj_mask, dtype=np.float64) * m
    mask = ((torch.tensor(m)[None, None] >= 0.5) * 1.0)[0, 0].float()
    return torch.from_numpy(d64).float().to(dev).contiguous(), mask.to(dev).contiguous()


def correspondence_field_device(d_dev, mask_dev, transform_in, focal_length=FOCAL_LENGTH):
    """A1 on device-resident inputs: d_dev, mask_dev (H,W) fp32 CUDA -> dict with coords (H,W,3) f32 = (x_norm, y_norm, Z), cam (3,H,W),
    centre (3,), Tc (4,4 host).  The 4x4 conjugation T' = C^-1 T C is done with torch on the host like the reference
    (warp_utils.py:431-437, 12-byte D2H of the centroid); everything per-pixel runs in csrc/geometry.cu."""
    dev = d_dev.device
    H, W = d_dev.shape
    K = torch.from_numpy(camera_matrix(focal_length, focal_length, W / 2.0, H / 2.0))[None].float()
    Kinv = K.inverse()[0].contiguous()
    cam = torch.empty(3, H, W, device=dev, dtype=torch.float32)
    centre4 = torch.empty(4, device=dev, dtype=torch.float32)
    call("gd_corr_pixel2cam", ptr(d_dev), ptr(mask_dev), H, W, _lib.host_f32(Kinv.reshape(-1).tolist()), ptr(cam), ptr(centre4), stream())
    centre = centre4[:3].cpu()  # 12-byte D2H, once per edit
    C = torch.eye(4, dtype=torch.float32)
    C[:3, 3] += -centre
    C = C[None]
    T = transform_in if torch.is_tensor(transform_in) else torch.tensor(np.asarray(transform_in))
    Tc = (C.inverse() @ T.cpu()[None].float() @ C).float()[0]
    coords = torch.empty(H, W, 3, device=dev, dtype=torch.float32)
    call("gd_corr_project", ptr(cam), H, W, _lib.host_f32(Tc[:3, :].reshape(-1).tolist()), _lib.host_f32(K[0].reshape(-1).tolist()),
         ptr(coords), stream())
    return dict(coords=coords, mask=mask_dev, cam=cam, centre=centre, Tc=Tc, depth=d_dev)


def correspondence_field(depth, obj_mask, transform_in, focal_length=FOCAL_LENGTH, device=None):
    """A1 from host inputs: depth (H,W) numpy, obj_mask (H,W) or None, transform_in (4,4)."""
    d_dev, mask_dev = stage_depth_mask(depth, obj_mask, device)
    return correspondence_field_device(d_dev, mask_dev, transform_in, focal_length)


def mesh_mask(coords, mask):
    """A2: amodal projected mask of the object's depth mesh (warp_utils.py:364-399 + 235-298). coords (H,W,3), mask (H,W)."""
    H, W = mask.shape
    out = torch.empty(H, W, device=coords.device, dtype=torch.float32)
    call("gd_mesh_mask", ptr(coords), ptr(mask), H, W, float(np.float32(float(1e-6) / float(2 * H))), ptr(out), stream())
    return out


def _morph(A, kernel, mode):
    A = _f32(A)
    shp = A.shape
    H, W = shp[-2:]
    B = int(A.numel() // (H * W))
    out = torch.empty_like(A)
    call("gd_morph", ptr(A), B, H, W, int(kernel), mode, ptr(out), stream())
    return out.reshape(shp)


def torch_erode(A, kernel=3):
    return _morph(A, kernel, 0)


def torch_dilate(A, kernel=3):
    return _morph(A, kernel, 1)


def resize_bilinear(x, size, channels_last=False):
    """T.Resize(BILINEAR, antialias=False): x (...,H,W) planes, or (...,H,W,C) when channels_last."""
    x = _f32(x)
    if channels_last:
        lead, (H, W, C) = x.shape[:-3], x.shape[-3:]
        outs = [torch.empty(size, size, C, device=x.device, dtype=torch.float32) for _ in range(max(1, int(np.prod(lead))))]
        xs = x.reshape(-1, H, W, C)
        for i, o in enumerate(outs):
            call("gd_resize_bilinear", ptr(xs[i]), C, H, W, 1, ptr(o), size, size, stream())
        return torch.stack(outs).reshape(*lead, size, size, C)
    H, W = x.shape[-2:]
    C = int(x.numel() // (H * W))
    out = torch.empty(*x.shape[:-2], size, size, device=x.device, dtype=torch.float32)
    call("gd_resize_bilinear", ptr(x), C, H, W, 0, ptr(out), size, size, stream())
    return out


def reshape_transform_coords(transform_coords, in_mat=None, in_mat_shape=None):
    """(B,H,W,3) -> (B,s,s,3); generic_torch.py:156-186"""
    s = in_mat.shape[-1] if in_mat is not None else in_mat_shape[-1]
    return resize_bilinear(_f32(transform_coords), int(s), channels_last=True)


def reshape_attention_mask(mask, in_mat=None, in_mat_shape=None):
    """(B,C,H,W) -> (B,C,s,s); generic_torch.py:189-207"""
    s = in_mat.shape[-1] if in_mat is not None else in_mat_shape[-1]
    return resize_bilinear(_f32(mask), int(s))


def build_masks(image_mask, mask_new_warped, amodal_mask, S):
    """attention_processors.py:338-360 in one kernel.  Inputs are (Hin,Hin) fp32 CUDA planes (mask_new_warped / amodal_mask may be
    None == zeros: the AttentionGeometryRemover case, attention_processors.py:856-866).  Returns dict of (S,S) planes."""
    image_mask = _f32(image_mask)
    Hin = image_mask.shape[-1]
    out = torch.empty(6, S, S, device=image_mask.device, dtype=torch.float32)
    call("gd_masks_build", ptr(image_mask), ptr(mask_new_warped), ptr(amodal_mask), Hin, S, ptr(out), stream())
    names = ("mask_new_warped", "mask_warp", "amodal_mask", "mask_intersection", "mask_1_empty", "mask_wo_edit")
    return {n: out[i] for i, n in enumerate(names)}


def splat_radius_ndc(S, radius_px=None):
    """warp_utils.py:94 (python double)"""
    return float(SPLATTER.radius if radius_px is None else radius_px) / float(S) * 2.0


def splat_index(coords, radius_px=None, points_per_pixel=None):
    """A3 integer part.  coords (B,S,S,3) fp32 CUDA -> idx int32 (B,S,S,K) [packed b*S*S+p], zbuf, dist2 (pytorch3d
    rasterize_points semantics, ties on z broken by ascending packed index)."""
    coords = _f32(coords)
    B, S = coords.shape[0], coords.shape[1]
    K = int(SPLATTER.points_per_pixel if points_per_pixel is None else points_per_pixel)
    idx = torch.empty(B, S, S, K, device=coords.device, dtype=torch.int32)
    zbuf = torch.empty(B, S, S, K, device=coords.device, dtype=torch.float32)
    dist2 = torch.empty(B, S, S, K, device=coords.device, dtype=torch.float32)
    call("gd_splat_index", ptr(coords), B, S, float(np.float32(splat_radius_ndc(S, radius_px))), K, ptr(idx), ptr(zbuf), ptr(dist2), stream())
    return idx, zbuf, dist2


def splat_composite(src, idx, dist2, *, channels_last=False, blend_mask=None, binarize=False, out_dtype=torch.float32,
                    radius_px=None, tau=None):
    """A3 float part: front-to-back alpha composite of `src` through idx/dist2, result rounded through fp16 as the reference's
    `.to(torch.half)` (warp_utils.py:176).  src (B,C,S,S) [or (B,S*S,C) when channels_last]; idx/dist2 (Bi,S,S,K), Bi in {1,B}.
    blend_mask (S*S): out = src*(1-m) + m*warped (attention_processors.py:544)."""
    assert src.dtype in (torch.float32, torch.bfloat16)
    src = src.contiguous()
    Bi, S, _, K = idx.shape
    P_ = S * S
    if channels_last:
        B, _, C = src.shape
    else:
        B, C = src.shape[:2]
    out = torch.empty(src.shape, device=src.device, dtype=out_dtype)
    r2 = float(np.float32(pow(splat_radius_ndc(S, radius_px), 2)))
    call("gd_splat_composite", ptr(src), 0 if src.dtype == torch.float32 else 1, 1 if channels_last else 0, ptr(idx), ptr(dist2), Bi, B,
         P_, C, K, r2, float(SPLATTER.tau if tau is None else tau), ptr(blend_mask), 1 if binarize else 0, ptr(out),
         0 if out_dtype == torch.float32 else 1, stream())
    return out


def warp_grid_edit(src, t_coords, padding_mode="zeros", mode="bilinear", align_corners=False, depth=None, use_softsplat=True,
                   splatting_radius=None, splatting_tau=None, splatting_points_per_pixel=None):
    """warp_utils.py:798-837: forward-splat `src` (B,C,S,S) through t_coords (B,S,S,3).  Returns fp32 holding fp16-rounded values."""
    if splatting_radius is not None:
        SPLATTER.radius = splatting_radius
    if splatting_tau is not None:
        SPLATTER.tau = splatting_tau
    if splatting_points_per_pixel is not None:
        SPLATTER.points_per_pixel = splatting_points_per_pixel
    src = _f32(src)
    t_coords = _f32(t_coords, src.device)
    assert src.shape[0] == t_coords.shape[0] and src.shape[-1] == t_coords.shape[1]
    idx, _, d2 = splat_index(t_coords)
    return splat_composite(src, idx, d2)


def get_transform_coordinates(image, depth, obj_mask=None, transform_in=torch.eye(4), use_softsplat=True, focal_length=FOCAL_LENGTH,
                              return_mesh=False, device=None):
    """vis_utils.py:404-479.  image (H,W,3) float in [0,1]; returns (t_coords (H,W,3) numpy, projected image numpy[, amodal mask numpy
    (1,1,H,W)]) like the reference; the CUDA tensors are kept on `get_transform_coordinates.last` for callers that stay on device."""
    g = correspondence_field(depth, obj_mask, transform_in, focal_length, device)
    coords = g["coords"]
    H, W = coords.shape[:2]
    idx, _, d2 = splat_index(coords[None])
    img = _f32(np.ascontiguousarray(np.transpose(np.asarray(image, dtype=np.float32), (2, 0, 1)))[None], coords.device)
    proj = splat_composite(img, idx, d2)
    valid = (coords[..., :2].abs().amax(-1) <= 1)
    np_image = np.clip((proj[0] * valid[None]).permute(1, 2, 0).cpu().numpy(), 0, 1)
    g["splat_idx"], g["splat_dist2"] = idx, d2
    get_transform_coordinates.last = g
    if return_mesh:
        mm = mesh_mask(coords, g["mask"])
        g["mesh_mask"] = mm
        return coords.cpu().numpy(), np_image, mm[None, None].cpu().numpy()
    return coords.cpu().numpy(), np_image


get_transform_coordinates.last = None
