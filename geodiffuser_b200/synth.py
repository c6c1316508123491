"""Seeded synthetic inputs for the BASELINE.json configs.

There is no network for datasets or checkpoints, so every test / bench input is generated here:
image = smooth 2-D sinusoid + uniform noise, mask = axis-aligned rectangle rows/cols 176..336 (scaled with the
image size), depth = constant 0.5 (2-D edits, reference `depth_predictor.get_constant_depth`:321) or a hemisphere
bump inside the mask with 1.0 outside (3-D edits), transform = translate(0.1,0,0) / rotate 20 deg about y / identity.
"""
import math

import numpy as np
import torch

SEED = 1234  # reference editor.py:47

EDIT_KINDS = ("translate2d", "rotate3d", "remove")


def translate_matrix(x, y, z):
    """reference vis_utils.py:68-76 (fp32 4x4)"""
    m = torch.eye(4)
    m[0, 3] += x
    m[1, 3] += y
    m[2, 3] += z
    return m


def rotate_axis(degrees, axis):
    """reference vis_utils.py:26-66 (fp64 4x4)"""
    r = math.radians(degrees)
    c, s = math.cos(r), math.sin(r)
    if axis == 2:
        m = [[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    elif axis == 1:
        m = [[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]]
    else:
        m = [[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]]
    return torch.tensor(m, dtype=torch.float64)


def edit_inputs(kind: str, size: int = 512, seed: int = SEED):
    """-> image (size,size,3) uint8, depth (size,size) f64, mask (size,size) f64 {0,1}, T (4,4) tensor"""
    assert kind in EDIT_KINDS, kind
    rs = np.random.RandomState(seed)
    v, u = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    base = 0.5 + 0.25 * np.sin(2 * np.pi * u / size * 3)[..., None] * np.cos(2 * np.pi * v / size * 2)[..., None]
    image = np.clip(base + 0.1 * (rs.rand(size, size, 3) - 0.5), 0, 1)
    image = (image * 255).astype(np.uint8)
    lo, hi = int(round(size * 176 / 512)), int(round(size * 336 / 512))
    mask = np.zeros((size, size), np.float64)
    mask[lo:hi, lo:hi] = 1.0
    if kind == "rotate3d":
        c = (lo + hi - 1) / 2.0
        R = (hi - lo) / 2.0 * math.sqrt(2.0) * 1.01
        rr = ((u - c) / R) ** 2 + ((v - c) / R) ** 2
        hemi = 1.0 - np.sqrt(np.clip(1.0 - rr, 0.0, 1.0))
        depth = np.where(mask > 0.5, 0.35 + 0.25 * hemi, 1.0)
        T = rotate_axis(20.0, 1)
    elif kind == "translate2d":
        depth = np.ones((size, size), np.float64) * 0.5
        T = translate_matrix(0.1, 0.0, 0.0)
    else:
        depth = np.ones((size, size), np.float64) * 0.5
        T = torch.eye(4)
    return image, depth, mask, T


def qkv(seed, B, H, N, Nk, d, scale=1.0):
    """q (B*H,N,d), k/v (B*H,Nk,d) fp32, head-major within batch (diffusers head_to_batch_dim)."""
    rs = np.random.RandomState(seed)
    q = (rs.randn(B * H, N, d) * scale).astype(np.float32)
    k = (rs.randn(B * H, Nk, d) * scale).astype(np.float32)
    v = rs.randn(B * H, Nk, d).astype(np.float32)
    return q, k, v
