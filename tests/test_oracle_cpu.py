"""The CPU oracle (test infrastructure) against the golden vectors that oracle/make_golden.py produced from the reference itself.
Sized to run in well under a minute."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, relerr
from geodiffuser_b200 import synth
from oracle import geodiff_oracle as O


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("kind", ["translate2d", "rotate3d", "remove"])
def test_oracle_geometry_vs_golden(kind):
    z = np.load(os.path.join(GOLDEN, f"geometry_{kind}.npz"))
    image, depth, mask, T = synth.edit_inputs(kind)
    g = O.corr_build(depth.copy(), mask.copy(), T)
    assert sha(g["coords"]) == str(z["coords512_sha"])
    for S in (16, 8):
        cS = z[f"coords{S}"]
        idx, _, d2 = O.splat_index(cS[None])
        np.testing.assert_array_equal(idx[0], z[f"idx{S}"])
        np.testing.assert_array_equal(d2[0], z[f"dist2_{S}"])
    amodal = O.erode3(O.mesh_mask(g["coords"], g["mask"]))
    assert sha(amodal) == str(z["amodal512_sha"])


def test_oracle_attention_layer_vs_golden():
    """edit_cross_S16_cfg: forward-only case small enough for the CPU suite"""
    z = np.load(os.path.join(GOLDEN, "attn_edit_cross_S16_cfg.npz"))
    zg = np.load(os.path.join(GOLDEN, "geometry_translate2d.npz"))
    S, H, d = 16, 2, 32
    masks = {k: zg[f"{k}{S}"] for k in ("mask_new_warped", "mask_warp", "amodal_mask", "mask_intersection", "mask_1_empty", "mask_wo_edit")}
    q, k, v = (torch.from_numpy(a) for a in synth.qkv(106, 4, H, S * S, 77, d))
    with torch.no_grad():
        res = O.edit_layer(q, k, v, True, d ** -0.5, H, (2, 3), (3, 4), masks, zg[f"coords{S}"], True, True)
    assert relerr(res["out"].numpy(), z["out"]) <= 2e-5


def test_oracle_cross_layer_at_product_shape_vs_golden():
    """edit_cross_S32_H8d80_opt (H = 8, head_dim 80, N = 1024 x 77 keys: cheap on the CPU): the oracle's layer with losses and its gradients against
    the golden the REFERENCE's AttentionGeometryEdit produced (oracle/make_golden.py:cross_cases) -- output, loss, every logged term, dQ and dK"""
    z = np.load(os.path.join(GOLDEN, "attn_edit_cross_S32_H8d80_opt.npz"))
    S, H, d = 32, 8, 80
    image, depth, mask, T = synth.edit_inputs("translate2d")
    g = O.corr_build(depth.copy(), mask.copy(), T)
    amodal = O.erode3(O.mesh_mask(g["coords"], g["mask"]))
    idx512, _, d2 = O.splat_index(g["coords"][None])
    mnw = O.binarize(O.splat_composite(mask.astype(np.float32)[None, None], idx512, d2))[0, 0]
    masks = O.build_masks(mask.astype(np.float32), mnw, amodal, S)
    q, k, v = (torch.from_numpy(a).requires_grad_(True) for a in synth.qkv(209, 2, H, S * S, 77, d))
    with torch.enable_grad():
        res = O.edit_layer(q, k, v, True, d ** -0.5, H, (0, 1), (1, 2), masks, O.resize_coords(g["coords"], S), False, True)
        gq, gk = torch.autograd.grad(res["loss"] + 0.37 * res["out"].sum(), [q, k])
    rows = z["rows"]
    assert relerr(res["out"].detach().numpy()[:, rows], z["out"]) <= 2e-5
    assert abs(float(res["loss"]) - float(z["loss"])) <= 2e-5 * abs(float(z["loss"]))
    for key in ("sim", "movement", "removal", "smoothness"):
        assert abs(float(res["terms"][key]) - float(z["term_" + key])) <= 5e-4 * max(abs(float(z["term_" + key])), 1e-6), key
    assert relerr(gq.numpy()[:, rows], z["dq"]) <= 1e-4
    assert relerr(gk.numpy(), z["dk"]) <= 1e-4


def test_oracle_elementwise_vs_golden():
    z = np.load(os.path.join(GOLDEN, "elementwise.npz"))
    lat, ctx = torch.from_numpy(z["lat"]), torch.from_numpy(z["ctx"])
    nl, nc = O.update_latent(lat, torch.from_numpy(z["w1"]), 0.3, z["mask512"], ctx, 0.5 * torch.from_numpy(z["w2"]))
    assert relerr(nl.numpy(), z["new_lat"]) <= 1e-6 and relerr(nc.numpy(), z["new_ctx"]) <= 1e-6
    al = O.ddim_alphas()
    np.testing.assert_allclose(al.numpy(), z["alphas"], rtol=1e-6)
    assert relerr(O.ddim_step(lat, torch.from_numpy(z["eps"]), 980, al).numpy(), z["ddim_980"]) <= 1e-6


def test_loop_golden_is_pinned_to_reference():
    for kind in ("translate2d", "rotate3d", "remove"):
        z = np.load(os.path.join(GOLDEN, f"loop_{kind}_tiny.npz"))
        assert float(z["pin_err_vs_reference"]) <= 1e-2
        assert np.isfinite(z["latents"]).all()


def test_postprocess_oracle_vs_reference_golden():
    """oracle/postprocess_oracle.py (numpy restatement of image_processing.py:24-77) reproduces, bit for bit, the float64 arrays the REAL
    reference function returned for the seeded inputs (tests/golden/postprocess.npz, made by oracle/make_golden_post.py)"""
    from oracle import postprocess_oracle as PO

    z = np.load(os.path.join(GOLDEN, "postprocess.npz"))
    for seed in (1, 2):
        src, tmpl, mask, mask_source = PO.synthetic_case(seed)
        out = PO.masked_histogram_matching(src, tmpl, mask, mask_source if bool(z[f"uses_mask_source{seed}"]) else None)
        assert out.dtype == np.float64
        np.testing.assert_array_equal(out, z[f"out{seed}"])
