"""The C-ABI library loads on a machine without a GPU and exports every symbol include/geodiffuser_b200.h declares, with the
argument counts the Python binding uses; and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "geodiffuser_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(gd_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(2)] = n
    return out


def test_library_exports_every_declared_symbol():
    from geodiffuser_b200 import _lib

    fns = header_functions()
    assert len(fns) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in fns:
        assert hasattr(L, name), f"{name} declared in the header but not exported"


def test_binding_table_matches_header():
    from geodiffuser_b200 import _lib

    fns = header_functions()
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in fns, f"{name} bound in _lib.py but missing from the header"
        assert len(argtypes) == fns[name], f"{name}: binding has {len(argtypes)} args, header {fns[name]}"
    for name in fns:
        assert name in _lib.SIGNATURES or name in ("gd_last_error", "gd_version"), name
    assert _lib.lib().gd_version() >= 100


def test_invalid_arguments_return_status_not_crash():
    from geodiffuser_b200 import _lib

    L = _lib.lib()
    rc = L.gd_splat_index(None, 1, 8, ctypes.c_float(0.1), 15, None, None, None, None)
    assert rc == 1 and b"gd" not in b"" and len(L.gd_last_error()) > 0   # GD_ERR_INVALID, message set
    with pytest.raises(_lib.GeoDiffuserB200Error):
        _lib.call("gd_morph", None, 1, 4, 4, 3, 0, None, None)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from geodiffuser_b200 import _lib, functional as Fn

    q = torch.zeros(2, 16, 8)
    with pytest.raises((_lib.GeoDiffuserB200Error, RuntimeError, AssertionError)):
        Fn.plain_attention(q, q, q, 1.0, 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "geodiffuser_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            assert "from oracle" not in src and "import oracle" not in src, fn


def test_stride_and_shape_validation_happens_before_any_launch():
    """argument checks of the strided entry points return a status (no CUDA call is made, so this runs without a GPU): slab strides must be
    positive multiples of 8 elements (16-byte rows for TMA / cp.async); the caller-side norms reject channel counts they do not serve"""
    from geodiffuser_b200 import _lib

    L = _lib.lib()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    arr = (ctypes.c_void_p * 1)(p.value)
    bad = (ctypes.c_long * 6)(40, 12, 40, 8, 40, 8)          # head stride 12: not a multiple of 8
    rc = L.gd_attn_fwd_generic(arr, arr, arr, arr, arr, None, 1, 1, 8, 8, 8, ctypes.c_float(1.0), bad, 0, None)
    assert rc == 1 and b"strides_ok" in L.gd_last_error()
    rc = L.gd_attn_bwd(0, p, p, p, p, p, p, None, None, None, 8, 0, p, 1, 8, 8, 8, ctypes.c_float(1.0), bad, 0, None)
    assert rc == 1
    rc = L.gd_attn_fwd_sm100(arr, arr, arr, arr, arr, None, 1, 1, 100, 100, 40, ctypes.c_float(1.0), None, 0, None)
    assert rc == 3 and b"N % 128" in L.gd_last_error()        # GD_ERR_UNSUPPORTED: shape outside the tcgen05 kernel's range
    rc = L.gd_group_norm_nhwc_fwd(p, None, p, p, 1, 1, 16, 36, 4, ctypes.c_float(1e-5), 1, p, 1 << 20, p, p, p, None)
    assert rc == 3 and b"multiple of 8" in L.gd_last_error()
    rc = L.gd_layer_norm_fwd(p, p, p, 4, 2048, ctypes.c_float(1e-5), p, p, p, None)
    assert rc == 3 and b"<= 1280" in L.gd_last_error()
    rc = L.gd_splat_composite_rows(p, 1, None, p, p, 2, 16, 12, 15, ctypes.c_float(0.1), ctypes.c_float(1.0), None, 0, p, 1, None, None)
    assert rc == 1                                            # C % 8 != 0
    rc = L.gd_masked_histogram_match(p, p, p, p, 0, 3, p, p, p, None)
    assert rc == 1                                            # npix == 0
    # removal rows of dQ: head dims / key counts the kernel was not built for, unaligned K slabs, odd leading dimensions
    rc = L.gd_removal_dq_rows(p, p, p, None, p, 1, 4, 1024, 1024, 64, ctypes.c_float(1.0), 1024, None, 0, None)
    assert rc == 3 and b"d in {40, 80}" in L.gd_last_error()
    rc = L.gd_removal_dq_rows(p, p, p, None, p, 1, 4, 1000, 1000, 40, ctypes.c_float(1.0), 1000, None, 0, None)
    assert rc == 3
    bad4 = (ctypes.c_long * 4)(40, 12, 40, 8)
    rc = L.gd_removal_dq_rows(p, p, p, None, p, 1, 4, 1024, 1024, 40, ctypes.c_float(1.0), 1024, bad4, 0, None)
    assert rc == 1 and b"16-byte" in L.gd_last_error()
    rc = L.gd_removal_weighted_rows(p, p, p, 1, 4, 1024, 1023, p, None)
    assert rc == 1                                            # ld < Nk
    rc = L.gd_attn_sm100_config(2, 96)
    assert rc == 1                                            # keys per step: 0 / 64 / 128 only
    assert L.gd_attn_sm100_config(2, 0) == 0
