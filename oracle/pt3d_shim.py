"""oracle/pt3d_shim.py -- TEST INFRASTRUCTURE ONLY.

Minimal stand-ins for the pytorch3d names warp_utils.py imports (warp_utils.py:5-18), backed by
oracle/pt3d_cpu.c.  Only the call shapes GeoDiffuser uses are supported.  PARITY UNPINNED (see
pt3d_cpu.c header): pytorch3d@89653419 is not vendored in /root/reference and cannot be installed.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libpt3d_cpu.so")
        if not os.path.exists(so):
            subprocess.check_call(["sh", os.path.join(_HERE, "build_oracle.sh")])
        L = ctypes.CDLL(so)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.pt3d_rasterize_points.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                            ctypes.c_int, ip, fp, fp]
        L.pt3d_rasterize_points.restype = None
        L.pt3d_alpha_composite.argtypes = [ip, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_long, fp]
        L.pt3d_alpha_composite.restype = None
        L.pt3d_mesh_coverage.argtypes = [fp, ctypes.c_long, ip, ctypes.c_long, ctypes.c_int, ctypes.c_float, fp]
        L.pt3d_mesh_coverage.restype = None
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def rasterize_points_np(pts, S, radius, K):
    """pts (B,P,3) fp32 in pytorch3d convention -> idx int32, zbuf, dist2 each (B,S,S,K)."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    B, P, _ = pts.shape
    idx = np.empty((B, S, S, K), np.int32)
    zbuf = np.empty((B, S, S, K), np.float32)
    d2 = np.empty((B, S, S, K), np.float32)
    lib().pt3d_rasterize_points(_fp(pts), B, P, S, np.float32(radius), K, _ip(idx), _fp(zbuf), _fp(d2))
    return idx, zbuf, d2


def alpha_composite_np(idx_bkss, alpha_bkss, feat_cp):
    idx_bkss = np.ascontiguousarray(idx_bkss, dtype=np.int32)
    alpha_bkss = np.ascontiguousarray(alpha_bkss, dtype=np.float32)
    feat_cp = np.ascontiguousarray(feat_cp, dtype=np.float32)
    B, K, S, _ = idx_bkss.shape
    C, Ptot = feat_cp.shape
    out = np.empty((B, C, S, S), np.float32)
    lib().pt3d_alpha_composite(_ip(idx_bkss), _fp(alpha_bkss), _fp(feat_cp), B, K, S, C, Ptot, _fp(out))
    return out


def mesh_coverage_np(verts, faces, S, blur):
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    out = np.empty((S, S), np.float32)
    lib().pt3d_mesh_coverage(_fp(verts), verts.shape[0], _ip(faces), faces.shape[0], S, np.float32(blur), _fp(out))
    return out


# ---- the pytorch3d-shaped API ------------------------------------------------------------
class Pointclouds:
    def __init__(self, points, features=None):
        self._points = points  # (B,P,3)
        self._features = features  # (B,P,C)

    def features_packed(self):
        return self._features.reshape(-1, self._features.shape[-1])


def rasterize_points(pointclouds, image_size, radius, points_per_pixel):
    pts = pointclouds._points.detach().cpu().numpy()
    idx, zbuf, d2 = rasterize_points_np(pts, int(image_size), float(radius), int(points_per_pixel))
    return torch.from_numpy(idx), torch.from_numpy(zbuf), torch.from_numpy(d2)


class compositing:
    @staticmethod
    def alpha_composite(pointsidx, alphas, pt_clds):
        out = alpha_composite_np(pointsidx.detach().cpu().numpy(), alphas.detach().cpu().numpy(),
                                 pt_clds.detach().cpu().numpy())
        return torch.from_numpy(out)


class TexturesVertex:
    def __init__(self, verts_features):
        self.verts_features = verts_features


class Fragments:
    def __init__(self, pix_to_face, zbuf, bary_coords, dists):
        self.pix_to_face, self.zbuf, self.bary_coords, self.dists = pix_to_face, zbuf, bary_coords, dists


class Meshes:
    def __init__(self, verts, faces, textures=None):
        self.verts, self.faces, self.textures = verts, faces, textures

    def sample_textures(self, frags):
        # all-ones vertex texture (warp_utils.py:385): value 1 where a face was hit (k = 0 slot), else 0
        cov = frags.pix_to_face  # (1,S,S,K) with coverage in slot 0
        return cov[..., None].float()


def rasterize_meshes(mesh, image_size, blur_radius, faces_per_pixel, perspective_correct=True):
    v = mesh.verts[0].detach().cpu().numpy()
    f = mesh.faces[0].detach().cpu().numpy()
    cov = mesh_coverage_np(v, f, int(image_size), float(blur_radius))
    p2f = torch.zeros(1, image_size, image_size, faces_per_pixel)
    p2f[0, :, :, 0] = torch.from_numpy(cov)
    return p2f, None, None, None
