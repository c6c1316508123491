"""oracle/make_golden_post.py -- TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):
    python -m oracle.make_golden_post
Runs the REAL reference `masked_histogram_matching` (GeoDiffuser/utils/image_processing.py, imported on CPU; scikit-image is absent and only
imported at module top, so it is stubbed) on seeded synthetic images, asserts oracle/postprocess_oracle.py returns the same float64 array bit
for bit, and writes tests/golden/postprocess.npz."""
import importlib
import os
import sys
import types

import numpy as np

from . import postprocess_oracle as PO
from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    ref_import.install_stubs()
    for n in ("skimage", "skimage.exposure"):
        try:
            importlib.import_module(n)
        except Exception:
            m = types.ModuleType(n)
            m.__path__ = []
            m.match_histograms = None
            sys.modules[n] = m
    sys.path.insert(0, ref_import.REF_ROOT)
    ip = importlib.import_module("GeoDiffuser.utils.image_processing")
    rec = {}
    for seed, use_src_mask in ((1, True), (2, False)):
        src, tmpl, mask, mask_source = PO.synthetic_case(seed)
        ms = mask_source if use_src_mask else None
        ref = ip.masked_histogram_matching(src, tmpl, mask, ms)
        mine = PO.masked_histogram_matching(src, tmpl, mask, ms)
        assert ref.dtype == np.float64 and np.array_equal(ref, mine), "restatement differs from the reference"
        rec[f"out{seed}"] = ref.astype(np.float64)
        rec[f"uses_mask_source{seed}"] = np.array(use_src_mask)
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **rec)
    print("wrote postprocess.npz", {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
