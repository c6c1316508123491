"""ctypes binding of the C-ABI CUDA library (include/geodiffuser_b200.h).

There is NO fallback: if libgeodiffuser_b200.so is missing or a call returns a non-zero status this module raises.
PyTorch is used by the callers only for device memory, streams and autograd plumbing.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GD_LIB_PATH") or os.path.join(_HERE, "libgeodiffuser_b200.so")   # (GD_LIB_PATH: A/B builds of kernel variants)

P, I, F, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_long

# name -> argument ctypes (every entry point returns int status); mirrors include/geodiffuser_b200.h
SIGNATURES = {
    "gd_corr_pixel2cam": [P, P, I, I, P, P, P, P],
    "gd_corr_project": [P, I, I, P, P, P, P],
    "gd_resize_bilinear": [P, I, I, I, I, P, I, I, P],
    "gd_masks_build": [P, P, P, I, I, P, P],
    "gd_splat_index": [P, I, I, F, I, P, P, P, P],
    "gd_splat_composite": [P, I, I, P, P, I, I, I, I, I, F, F, P, I, P, I, P],
    "gd_mesh_mask": [P, P, I, I, F, P, P],
    "gd_morph": [P, I, I, I, I, I, P, P],
    "gd_attn_fwd_generic": [P, P, P, P, P, P, I, I, I, I, I, F, P, I, P],
    "gd_attn_fwd_sm100": [P, P, P, P, P, P, I, I, I, I, I, F, P, I, P],
    "gd_attn_sm100_config": [I, I],
    "gd_attn_bwd_prep": [P, I, P, P, P, P, P, P, P, I, I, I, I, P, P, P],
    "gd_attn_bwd": [I, P, P, P, P, P, P, P, P, P, I, I, P, I, I, I, I, F, P, I, P],
    "gd_attn_bwd_dk_split": [P, P, P, P, P, P, P, P, P, I, I, P, P, I, I, I, I, I, F, P, I, P],
    "gd_attn_bwd_sm100": [P, P, P, P, P, P, P, P, P, I, I, P, I, I, I, F, P, I, I, P],
    "gd_cast_f32_to_bf16": [P, P, L, P],
    "gd_attn_probs": [P, P, P, P, I, I, I, I, I, F, P, I, P, P],
    "gd_corr_max_partial": [P, P, I, I, I, I, I, P, P, P, P],
    "gd_removal_corr_sm100": [P, P, P, P, I, I, I, I, F, I, P, P, P, P, P],
    "gd_attn_probs_rows2": [P, P, P, P, I, I, I, I, I, F, P, I, P, P],
    "gd_removal_extra_rows": [P, P, I, I, I, I, P, I, P],
    "gd_removal_weighted_rows": [P, P, P, I, I, I, I, P, P],
    "gd_removal_dq_rows": [P, P, P, P, P, I, I, I, I, I, F, I, P, I, P],
    "gd_attn_l1_losses": [P, P, P, P, P, P, P, F, F, F, F, F, I, I, I, P, P, I, P],
    "gd_removal_finalize": [P, I, I, I, I, P, P, P, F, P, P, I, I, I, P, P, P, P, P, P],
    "gd_loss_reduce": [P, I, P, I, P, P, P, F, P, P, P],
    "gd_amodal_knn": [P, I, P, P, P, P],
    "gd_amodal_target": [P, P, P, P, P, I, I, I, P, P, P],
    "gd_blend_rows": [P, P, P, P, I, I, I, P, I, P, P],
    "gd_splat_composite_rows": [P, I, P, P, P, I, I, I, I, F, F, P, I, P, I, P, P],
    "gd_ddim_step": [P, P, P, I, F, F, F, F, F, L, P, P, P],
    "gd_latent_update": [P, P, P, I, F, L, P, P],
    "gd_norm_rescale": [P, L, F, P, P, P],
    "gd_latent_blend": [P, P, P, I, I, L, P, P],
    "gd_group_norm_nhwc_fwd": [P, P, P, P, I, I, I, I, I, F, I, P, L, P, P, P, P],
    "gd_group_norm_nhwc_bwd": [P, P, P, P, P, I, P, I, I, I, I, I, P, L, P, P, P],
    "gd_masked_histogram_match": [P, P, P, P, L, I, P, P, P, P],
    "gd_layer_norm_fwd": [P, P, P, L, I, F, P, P, P, P],
    "gd_geglu_fwd": [P, L, I, P, P],
    "gd_geglu_bwd": [P, P, L, I, P, P],
    "gd_add_bias_residual": [P, P, P, L, I, P, P],
    "gd_group_norm_nhwc_workspace": [I, I, I, I],
    "gd_group_norm_config": [I],
}

_LIB = None
HAS_SM100 = "gd_attn_fwd_sm100" in SIGNATURES


class _Counters:
    """per host thread (runner.EditWorkers drives several edits of one GPU from several threads): kernels launched through the C ABI and their
    algorithmic attention-path FLOP.  `_lib.LAUNCHES` / `_lib.FLOPS` read the sum over all threads (bench.py: gpu_launches, edit-level roofline)."""
    __slots__ = ("launches", "flops")

    def __init__(self):
        self.launches, self.flops = 0, 0.0


_TLS = threading.local()
_ALL_COUNTERS = []


def counters():
    c = getattr(_TLS, "override", None) or getattr(_TLS, "c", None)
    if c is None:
        c = _TLS.c = _Counters()
        _ALL_COUNTERS.append(c)
    return c


class count_into:
    """`with count_into(c):` -- launches made by this thread are booked on the counters `c` of another thread.  The backward of the fused layer
    runs on autograd's device thread; it books its launches on the thread that ran the forward (the edit lane), which is the one that records and
    replays the optimisation-pass graph (graphs.GraphedGradPass)."""

    def __init__(self, c):
        self.c = c

    def __enter__(self):
        self.prev = getattr(_TLS, "override", None)
        _TLS.override = self.c

    def __exit__(self, *exc):
        _TLS.override = self.prev


def __getattr__(name):      # module attribute access: the process-wide totals
    if name == "LAUNCHES":
        return sum(c.launches for c in _ALL_COUNTERS)
    if name == "FLOPS":
        return sum(c.flops for c in _ALL_COUNTERS)
    raise AttributeError(name)


KERNELS_PER_CALL = {"gd_attn_sm100_config": 0, "gd_corr_pixel2cam": 2, "gd_removal_finalize": 1, "gd_amodal_target": 2, "gd_attn_bwd_dk_split": 2,
                    "gd_group_norm_nhwc_fwd": 1, "gd_group_norm_nhwc_bwd": 2, "gd_group_norm_nhwc_workspace": 0, "gd_group_norm_config": 0, "gd_masked_histogram_match": 3}  # every other entry point launches one


class GeoDiffuserB200Error(RuntimeError):
    pass


def lib():
    """Loads the shared library once.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise GeoDiffuserB200Error(
                f"{LIB_PATH} not found: the sm_100a CUDA extension is not built and there is no CPU fallback. "
                "Run `sh geodiffuser_b200/csrc/build.sh`.")
        L_ = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L_, name)  # AttributeError here == header / library mismatch
            fn.argtypes = args
            fn.restype = ctypes.c_int
        L_.gd_last_error.restype = ctypes.c_char_p
        L_.gd_version.restype = ctypes.c_int
        _LIB = L_
    return _LIB


PROFILE = None  # bench.py sets this to {} to time selected entry points with CUDA events on the launching (current) stream


def profile_begin(names):
    """start collecting (start, end) CUDA-event pairs for the named entry points; meta() may tag each call (e.g. with its FLOPs)"""
    global PROFILE
    PROFILE = {"names": set(names), "events": []}


def profile_end():
    """-> {name: [(milliseconds, tag), ...]} ; call after torch.cuda.synchronize()"""
    global PROFILE
    out = {}
    for name, tag, e0, e1 in PROFILE["events"]:
        out.setdefault(name, []).append((e0.elapsed_time(e1), tag))
    PROFILE = None
    return out


def algorithmic_flops(name, tag):
    """forward 4*G*H*N*Nk*d (QK^T + PV per stream), backward 6*H*N*Nk*d (recompute S, dP, dQ or dK), removal correlation
    2*H*M*N*Nk.  `tag` = the shape tuple the caller passes with the launch."""
    if tag is None:
        return 0.0
    if name in ("gd_attn_fwd_sm100", "gd_attn_fwd_generic"):
        G, H, N, Nk, d = tag
        return 4.0 * G * H * N * Nk * d
    if name in ("gd_attn_bwd_sm100", "gd_attn_bwd", "gd_attn_bwd_dk_split"):
        H, N, Nk, d = tag
        return 6.0 * H * N * Nk * d
    if name in ("gd_corr_max_partial", "gd_removal_corr_sm100"):
        H, M, N, Nk = tag
        return 2.0 * H * M * N * Nk
    return 0.0      # (gd_attn_probs re-materialises map rows for the correlation: implementation work, not in SURVEY 8(d)'s count)


def call(name, *args, tag=None):
    L_ = lib()
    cnt = counters()
    if tag is not None:
        cnt.flops += algorithmic_flops(name, tag)
    if PROFILE is not None and name in PROFILE["names"] and not torch.cuda.is_current_stream_capturing():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(L_, name)(*args)
        e1.record()
        PROFILE["events"].append((name, tag, e0, e1))
    else:
        rc = getattr(L_, name)(*args)
    cnt.launches += KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise GeoDiffuserB200Error(f"{name} failed with status {rc}: {L_.gd_last_error().decode()}")


def ptr(t):
    """device pointer of a contiguous CUDA tensor (None -> NULL)"""
    if t is None:
        return None
    if not t.is_cuda:
        raise GeoDiffuserB200Error("expected a CUDA tensor: geodiffuser_b200 has no CPU path")
    if not t.is_contiguous():
        raise GeoDiffuserB200Error("expected a contiguous tensor")
    return ctypes.c_void_p(t.data_ptr())


def host_f32(values):
    """host float array for the few entry points that take small host-side matrices"""
    arr = (ctypes.c_float * len(values))(*[float(v) for v in values])
    return arr


def ptr_array(tensors):
    """host array of device pointers (None -> NULL)"""
    arr = (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])
    return arr


def base_ptr(t):
    """device pointer of the first element of a (possibly strided) CUDA tensor view: for the slab arguments, whose strides travel separately"""
    if t is None:
        return None
    if not t.is_cuda:
        raise GeoDiffuserB200Error("expected a CUDA tensor: geodiffuser_b200 has no CPU path")
    return ctypes.c_void_p(t.data_ptr())


def host_longs(values):
    """host long array (element strides of slab arguments); None -> NULL (contiguous)"""
    if values is None:
        return None
    return (ctypes.c_long * len(values))(*[int(v) for v in values])


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
