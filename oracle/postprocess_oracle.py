"""oracle/postprocess_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU restatement; never imported by geodiffuser_b200/).

Masked histogram matching, following /root/reference/GeoDiffuser/utils/image_processing.py:24-77 line by line in numpy.
Pinned: oracle/make_golden_post.py runs the REAL reference function (imported from /root/reference on CPU) on the seeded inputs and
asserts this restatement returns the same float64 array bit for bit before writing tests/golden/postprocess.npz."""
import numpy as np


def match_cumulative_cdf(source, template, mask=None, mask_source=None):
    """image_processing.py:24-65 (source, template: one uint8 channel)"""
    if mask is None:
        mask = np.ones_like(source)
    if mask_source is None:
        mask_source = mask
    src_lookup = source[mask_source > 0.5].reshape(-1)
    src_counts = np.bincount(src_lookup, minlength=256)
    tmpl_counts = np.bincount(template[mask > 0.5].reshape(-1), minlength=256)
    tmpl_values = np.linspace(0, 255, 256).astype("uint8")
    src_quantiles = np.cumsum(src_counts) / source[mask_source > 0.5].size
    tmpl_quantiles = np.cumsum(tmpl_counts) / template[mask > 0.5].size
    interp_a_values = np.interp(src_quantiles, tmpl_quantiles, tmpl_values)
    return interp_a_values[source.reshape(-1)].reshape(source.shape)


def masked_histogram_matching(source, template, mask=None, mask_source=None):
    """image_processing.py:68-77"""
    return np.stack([match_cumulative_cdf(source[..., i], template[..., i], mask, mask_source) for i in range(source.shape[-1])], -1)


def synthetic_case(seed, H=96, W=96):
    """seeded uint8 images with different tone curves, flat regions (empty histogram bins -> duplicate quantiles) and two masks"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = 127.5 + 100.0 * np.sin(xx / 9.0 + seed) * np.cos(yy / 13.0)
    src = np.clip(base[..., None] + rng.normal(0, 20, (H, W, 3)), 0, 255).astype(np.uint8)
    tmpl = np.clip(0.6 * base[..., None] + 40 + rng.normal(0, 10, (H, W, 3)), 0, 255).astype(np.uint8)
    tmpl[..., 2] = (tmpl[..., 2] // 16) * 16                      # posterised channel: many empty bins
    mask = ((xx - W / 2) ** 2 + (yy - H / 2) ** 2 > (H / 4) ** 2).astype(np.float64)
    mask_source = (xx > W // 5).astype(np.float64)
    return src, tmpl, mask, mask_source
