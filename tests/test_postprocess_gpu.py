"""SURVEY 8(f) row N3 -- masked histogram matching on the device (csrc/postprocess.cu) against the golden arrays made by the reference's own
image_processing.masked_histogram_matching and against the CPU oracle at image size: float64 results, bit-exact (integer histograms, and
np.interp restated with explicit IEEE double operations)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_histogram_matching_bit_exact_vs_reference_golden():
    from geodiffuser_b200 import image_processing as IP
    from oracle import postprocess_oracle as PO

    z = np.load(os.path.join(GOLDEN, "postprocess.npz"))
    for seed in (1, 2):
        src, tmpl, mask, mask_source = PO.synthetic_case(seed)
        out = IP.masked_histogram_matching(src, tmpl, mask, mask_source if bool(z[f"uses_mask_source{seed}"]) else None)
        assert out.dtype == np.float64 and out.shape == src.shape
        np.testing.assert_array_equal(out, z[f"out{seed}"])


@pytest.mark.parametrize("H,W", [(512, 512), (768, 768), (37, 53)])
def test_histogram_matching_bit_exact_vs_oracle_at_image_size(H, W):
    from geodiffuser_b200 import image_processing as IP
    from oracle import postprocess_oracle as PO

    src, tmpl, mask, mask_source = PO.synthetic_case(H + W, H, W)
    ref = PO.masked_histogram_matching(src, tmpl, mask, mask_source)
    out = IP.masked_histogram_matching(src, tmpl, mask, mask_source)
    np.testing.assert_array_equal(out, ref)
    # identity mask (the reference's default) and a tensor input staying on the device
    ref1 = PO.masked_histogram_matching(src, tmpl)
    out1 = IP.masked_histogram_matching(torch.from_numpy(src).cuda(), torch.from_numpy(tmpl).cuda())
    assert out1.is_cuda
    np.testing.assert_array_equal(out1.cpu().numpy(), ref1)
    # property: the remap is monotone in the source value (matching an image to itself is NOT the identity when histogram bins are empty:
    # np.interp then returns the last of the tied quantiles -- reference behaviour, reproduced)
    for c in range(3):
        order = np.argsort(src[..., c].reshape(-1), kind="stable")
        assert np.all(np.diff(out[..., c].reshape(-1)[order]) >= 0)


def test_empty_mask_is_an_error():
    from geodiffuser_b200 import image_processing as IP
    from oracle import postprocess_oracle as PO

    src, tmpl, mask, _ = PO.synthetic_case(3)
    with pytest.raises(ValueError):
        IP.masked_histogram_matching(src, tmpl, np.zeros_like(mask))


def test_postprocess_edited_image_follows_the_reference_tail():
    """editor.py:659-690 on synthetic inputs: image warp through the splat, mask algebra, histogram matching -- against the same steps on the
    CPU oracle"""
    from geodiffuser_b200 import geometry as G, image_processing as IP, synth
    from oracle import geodiff_oracle as O, postprocess_oracle as PO

    image, depth, mask, T = synth.edit_inputs("rotate3d")
    g = G.correspondence_field(depth.copy(), mask.copy(), T)
    idx, _, d2 = G.splat_index(g["coords"][None])
    m = torch.from_numpy(mask.astype(np.float32)).cuda()
    mnw = G.splat_composite(m[None, None].contiguous(), idx, d2, binarize=True)[0, 0]
    img_u8 = (np.clip(np.asarray(image, dtype=np.float64), 0, 1) * 255).astype(np.uint8) if np.asarray(image).max() <= 1.0 else np.asarray(image).astype(np.uint8)
    rng = np.random.default_rng(0)
    edited = np.clip(img_u8.astype(np.float64) * 0.8 + rng.normal(0, 12, img_u8.shape) + 20, 0, 255).astype(np.uint8)
    out = IP.postprocess_edited_image(edited, img_u8, g["coords"], mnw, mask, "geometry_editor")
    # oracle: same tail in numpy
    coords = g["coords"].cpu().numpy()
    warped = O.warp_grid_edit((img_u8.transpose(2, 0, 1)[None] / 255.0).astype(np.float32), coords[None])
    p_image = (warped[0].transpose(1, 2, 0) * 255.0).astype("uint8")
    mask_edit = mnw.cpu().numpy().astype(np.float64)
    mask_changed = ((mask_edit + mask) > 0.5) * 1.0
    mask_wo = ((np.ones_like(mask_changed) - mask_changed) > 0.5) * 1.0
    p_new = (mask_wo[..., None] * img_u8 + mask_edit[..., None] * p_image).astype("uint8")
    msrc = ((mask_edit + mask_wo) > 0.5) * 1.0
    ref = PO.masked_histogram_matching(edited, p_new, msrc, msrc)
    np.testing.assert_array_equal(out, ref)
    out_r = IP.postprocess_edited_image(edited, img_u8, None, None, mask, "geometry_remover")
    np.testing.assert_array_equal(out_r, PO.masked_histogram_matching(edited, img_u8, 1.0 - mask))
