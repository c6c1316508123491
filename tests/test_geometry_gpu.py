"""GPU parity of subsystem (1) -- correspondence field, masks, splat index/composite, amodal mesh mask -- through the C ABI
against (a) the committed golden vectors made by oracle/make_golden.py from the reference itself and (b) the CPU oracle on the
same seeded inputs.  Integer / index / binarised artefacts: bit-exact.  Float warps: 1e-3 (one fp16 ulp, see make_golden)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, relerr
from geodiffuser_b200 import synth

pytestmark = pytest.mark.gpu

KINDS = ("translate2d", "rotate3d", "remove")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def geo():
    from geodiffuser_b200 import geometry as G

    out = {}
    for kind in KINDS:
        image, depth, mask, T = synth.edit_inputs(kind)
        g = G.correspondence_field(depth.copy(), mask.copy(), T)
        g["mesh"] = G.mesh_mask(g["coords"], g["mask"])
        g["amodal"] = G.torch_erode(g["mesh"][None, None])[0, 0]
        g["idx512"], _, g["d2_512"] = G.splat_index(g["coords"][None])
        m = torch.from_numpy(mask.astype(np.float32)).cuda()
        g["mnw"] = G.splat_composite(m[None, None].contiguous(), g["idx512"], g["d2_512"], binarize=True)[0, 0]
        g["obj_mask"] = mask
        out[kind] = g
    return out


@pytest.mark.parametrize("kind", KINDS)
def test_coords512_bit_exact_vs_golden(geo, kind):
    z = np.load(os.path.join(GOLDEN, f"geometry_{kind}.npz"))
    assert sha(geo[kind]["coords"].cpu().numpy()) == str(z["coords512_sha"])
    np.testing.assert_array_equal(geo[kind]["centre"].numpy(), z["centre"])


@pytest.mark.parametrize("kind", KINDS)
def test_coords512_bit_exact_vs_oracle(geo, kind):
    from oracle import geodiff_oracle as O

    image, depth, mask, T = synth.edit_inputs(kind)
    ref = O.corr_build(depth.copy(), mask.copy(), T)
    np.testing.assert_array_equal(geo[kind]["coords"].cpu().numpy(), ref["coords"])
    np.testing.assert_array_equal(geo[kind]["cam"].cpu().numpy(), ref["cam"])


@pytest.mark.parametrize("kind", KINDS)
def test_mask_warp_and_amodal_bit_exact(geo, kind):
    z = np.load(os.path.join(GOLDEN, f"geometry_{kind}.npz"))
    g = geo[kind]
    if int(z["coords_mismatch_vs_reference"]) == 0:   # golden idx512 was built on the reference's own coords
        assert sha(g["idx512"].cpu().numpy()) == str(z["idx512_sha"])
        assert sha(g["mnw"].cpu().numpy()) == str(z["mask_new_warped512_sha"])
    # ... and on the canonical coords (rotate3d: the reference's fp32 torch.mean centroid moves the coords by a few ulp; the golden records how
    # many of the 512^2 x 15 index entries / warped-mask pixels that changes)
    assert sha(g["idx512"].cpu().numpy()) == str(z["idx512_canonical_sha"])
    assert sha(g["mnw"].cpu().numpy()) == str(z["mask_new_warped512_canonical_sha"])
    assert int(z["mask_new_warped512_mismatch_canonical_vs_reference_coords"]) == 0
    assert float(g["mnw"].sum()) == float(z["mask_new_warped512_sum"])
    assert sha(g["amodal"].cpu().numpy()) == str(z["amodal512_sha"])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("S", (64, 32, 16, 8))
def test_per_resolution_fields_bit_exact(geo, kind, S):
    from geodiffuser_b200 import geometry as G

    z = np.load(os.path.join(GOLDEN, f"geometry_{kind}.npz"))
    g = geo[kind]
    cS = G.reshape_transform_coords(g["coords"][None], in_mat_shape=(1, 1, S, S))[0]
    if int(z["coords_mismatch_vs_reference"]) == 0:
        np.testing.assert_array_equal(cS.cpu().numpy(), z[f"coords{S}"])
    idx, zb, d2 = G.splat_index(cS[None])
    np.testing.assert_array_equal(idx[0].cpu().numpy(), z[f"idx{S}"])          # THE bit-exact artefact
    if int(z["coords_mismatch_vs_reference"]) == 0:
        np.testing.assert_array_equal(d2[0].cpu().numpy(), z[f"dist2_{S}"])
    mk = G.build_masks(torch.from_numpy(g["obj_mask"].astype(np.float32)).cuda(), g["mnw"], g["amodal"], S)
    for name in ("mask_new_warped", "mask_warp", "amodal_mask", "mask_intersection", "mask_1_empty", "mask_wo_edit"):
        np.testing.assert_array_equal(mk[name].cpu().numpy(), z[f"{name}{S}"], err_msg=name)
    # float part: feature warp
    rs = np.random.RandomState(7 + S)
    feat = rs.randn(2, 3, S, S).astype(np.float32)
    cS_g = torch.from_numpy(z[f"coords{S}"]).cuda()
    w = G.warp_grid_edit(torch.from_numpy(feat).cuda(), cS_g[None].tile(2, 1, 1, 1)).cpu().numpy()
    assert relerr(w, z[f"warp_feat{S}"]) <= 1e-3
    assert (w != z[f"warp_feat{S}"]).sum() == 0     # in fact identical: IEEE sqrt/div in both


def test_splat_index_edge_cases():
    """empty pixels (-1 fill), points behind the camera, NaNs, more than K candidates per pixel, ties on z"""
    from geodiffuser_b200 import geometry as G
    from oracle import geodiff_oracle as O

    rs = np.random.RandomState(0)
    S = 16
    coords = np.zeros((2, S, S, 3), np.float32)
    coords[..., :2] = rs.uniform(-0.2, 0.2, (2, S, S, 2))      # everything piles into few pixels -> > K candidates
    coords[..., 2] = 0.5                                        # all z tie
    coords[0, 0, 0, 2] = -1.0                                   # behind camera
    coords[0, 0, 1, 0] = np.nan
    coords[1, :, :, :2] = rs.uniform(-3, 3, (S, S, 2))          # mostly off-screen -> empty pixels
    coords[1, :, :, 2] = rs.uniform(0.1, 2.0, (S, S))
    i_ref, z_ref, d_ref = O.splat_index(coords)
    idx, zb, d2 = G.splat_index(torch.from_numpy(coords).cuda())
    np.testing.assert_array_equal(idx.cpu().numpy(), i_ref)
    np.testing.assert_array_equal(zb.cpu().numpy(), z_ref)
    np.testing.assert_array_equal(d2.cpu().numpy(), d_ref)
    assert (i_ref == -1).any() and (i_ref[0, S // 2, S // 2] >= 0).all()


def test_splat_index_768_property():
    """BASELINE config 4 size (96^2 tokens): sortedness + radius property instead of an oracle run"""
    from geodiffuser_b200 import geometry as G

    S = 96
    ys, xs = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    coords = np.stack([2 * xs / (S - 1) - 1 + 0.013, 2 * ys / (S - 1) - 1 - 0.007, 0.3 + 0.001 * (xs + ys)], -1).astype(np.float32)
    idx, zb, d2 = G.splat_index(torch.from_numpy(coords[None]).cuda())
    zb, d2, idx = zb.cpu().numpy()[0], d2.cpu().numpy()[0], idx.cpu().numpy()[0]
    r = np.float32(G.splat_radius_ndc(S))
    live = idx >= 0
    assert (d2[live] < r * r).all()
    both = live[..., 1:] & live[..., :-1]
    assert (zb[..., 1:][both] >= zb[..., :-1][both]).all()      # ascending z
    assert not (live[..., 1:] & ~live[..., :-1]).any()          # empties last
    assert live[S // 2, S // 2].sum() >= 4


def test_morph_matches_conv():
    from geodiffuser_b200 import geometry as G

    rs = np.random.RandomState(3)
    a = (rs.rand(2, 1, 40, 40) > 0.6).astype(np.float32)
    t = torch.from_numpy(a)
    k3, k5 = torch.ones(1, 1, 3, 3), torch.ones(1, 1, 5, 5)
    er = (torch.nn.functional.conv2d(t, k3, padding=1) == 9.0) * 1.0
    di = (torch.nn.functional.conv2d(t, k5, padding=2) >= 1) * 1.0
    np.testing.assert_array_equal(G.torch_erode(t.cuda(), 3).cpu().numpy(), er.numpy())
    np.testing.assert_array_equal(G.torch_dilate(t.cuda(), 5).cpu().numpy(), di.numpy())


@pytest.mark.parametrize("kind,S,B,C,dt", [("rotate3d", 64, 8, 40, torch.bfloat16), ("translate2d", 32, 8, 80, torch.bfloat16),
                                           ("rotate3d", 16, 3, 18, torch.float32), ("translate2d", 64, 8, 40, torch.float32)])
def test_query_warp_rows_kernel_bit_exact_vs_pixel_kernel_and_oracle(geo, kind, S, B, C, dt):
    """the per-layer query warp (one warp per pixel, blend weights formed once per pixel) == the one-thread-per-element composite on the
    same data, bit for bit, with and without the M_edit blend; and == the CPU oracle's composite"""
    from geodiffuser_b200 import geometry as G
    from oracle import geodiff_oracle as O

    cS = G.reshape_transform_coords(geo[kind]["coords"][None], in_mat_shape=(1, 1, S, S))
    idx, _, d2 = G.splat_index(cS)
    g = torch.Generator(device="cuda").manual_seed(S + C)
    src = torch.randn(B, S * S, C, device="cuda", generator=g).to(dt)
    blend = torch.rand(S * S, device="cuda", generator=g)
    nchw = src.permute(0, 2, 1).reshape(B, C, S, S).contiguous()
    for bm in (None, blend):
        rows = G.splat_composite(src, idx, d2, channels_last=True, blend_mask=bm, out_dtype=dt)
        pix = G.splat_composite(nchw, idx, d2, channels_last=False, blend_mask=bm, out_dtype=dt)
        assert torch.equal(rows, pix.reshape(B, C, S * S).permute(0, 2, 1))
    ref = O.splat_composite(nchw[:1].float().cpu().numpy(), idx.cpu().numpy(), d2.cpu().numpy())
    got = G.splat_composite(src, idx, d2, channels_last=True, out_dtype=torch.float32)
    assert relerr(got[:1].permute(0, 2, 1).reshape(1, C, S, S).cpu().numpy(), ref) <= 1e-3   # one fp16 ulp (see make_golden)


@pytest.mark.parametrize("S", [64, 96])
def test_amodal_knn_table_is_the_serial_scan(S):
    """gd_amodal_knn (one warp per pixel, lane-local top 4 + warp merge) must give the table a serial scan over all candidates gives: the 4 largest
    inverse grid distances to foreground pixels per pixel, ties by ascending pixel index (attention_sharing.py:79-83).  Checked against torch on the
    values (sorted) and on the selected distances; the weights w = exp(-(1 / max inv) / 5)."""
    from geodiffuser_b200._lib import call, ptr, stream
    from oracle import geodiff_oracle as O

    N = S * S
    m = torch.zeros(S, S, device="cuda")
    m[S // 5:S // 2, S // 4:S // 2 + 3] = 1.0
    m[S - 7, 3] = 1.0
    m = m.reshape(-1).contiguous()
    idx = torch.empty(N, 4, device="cuda", dtype=torch.int32)
    val = torch.empty(N, 4, device="cuda")
    w = torch.empty(N, device="cuda")
    call("gd_amodal_knn", ptr(m), S, ptr(idx), ptr(val), ptr(w), stream())
    dgrid = O.distance_grid(S).cuda()                                                # (1, N, N)
    fg = (m > 0.5).float()
    inv = 1.0 / (dgrid[0] * 512 / 2.0 + 100000 * (1.0 - fg)[None, :] + 1e-4)
    ref = torch.topk(inv, k=4, dim=-1, largest=True, sorted=True)
    # (the kernel forms the grid coordinates as (2i+1)/S - 1, torch's affine_grid as a linspace: distances agree to ~1e-7 absolute, i.e. ~1e-5
    #  relative between neighbouring pixels; neighbouring CANDIDATES differ by per cent)
    assert torch.allclose(val, ref.values, rtol=1e-4, atol=0)
    picked = torch.gather(inv, 1, idx.long())
    assert torch.allclose(picked, ref.values, rtol=1e-4, atol=0)                    # the indices point at those candidates
    assert (fg[idx.long()] == 1).all() and (idx[:, 0] != idx[:, 1]).all()
    # equal distances: lower pixel index first
    same = val[:, :-1] == val[:, 1:]
    assert (idx[:, :-1][same] < idx[:, 1:][same]).all()
    assert torch.allclose(w, torch.exp(-(1 / ref.values[:, 0]) / 5), rtol=1e-4)
