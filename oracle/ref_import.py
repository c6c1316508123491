"""oracle/ref_import.py -- TEST INFRASTRUCTURE ONLY.

Imports the *real* GeoDiffuser hot-path modules from /root/reference on CPU so that
`oracle/make_golden.py` can (a) pin the restatement in `oracle/geodiff_oracle.py` against
the reference's own code and (b) write golden vectors under tests/golden/.

/root/reference exists only in the build container; nothing under tests/ -m gpu, smoke()
or bench.py imports this file.

What is stubbed (packages absent from this image; none of them is on the arithmetic path
except pytorch3d):
  matplotlib, mpl_toolkits, IPython, cupy, cv2-free paths untouched, diffusers
  (only `USE_PEFT_BACKEND=False` is read, attention_processors.py:10), pytorch3d.*.
What is monkey-patched:
  * `DISTANCE_CLASS.get_coord_distance` default device "cuda" -> "cpu" (generic_torch.py:132)
  * `warp_utils.SPLATTER` / `warp_grid_edit` forced-to-"cuda" move (warp_utils.py:809-812)
    replaced by the same Python body running on CPU with pytorch3d's two operators provided
    by oracle/pt3d_cpu.c (PARITY UNPINNED at that boundary, see that file's header).
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("GEODIFFUSER_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "GeoDiffuser", "utils"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave as a package
    sys.modules[name] = m
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, n):
        return _Anything()


def install_stubs():
    import torch  # noqa: F401

    def need(name):
        try:
            importlib.import_module(name)
            return False
        except Exception:
            return True

    if need("matplotlib"):
        _stub("matplotlib", pyplot=_Anything())
        _stub("matplotlib.pyplot", imshow=_Anything(), show=_Anything(), imsave=_Anything(), figure=_Anything())
    if need("mpl_toolkits"):
        _stub("mpl_toolkits")
        _stub("mpl_toolkits.mplot3d", Axes3D=_Anything)
    if need("IPython"):
        _stub("IPython")
        _stub("IPython.display", display=_Anything())
    if need("cupy"):
        _stub("cupy", memoize=lambda **k: (lambda f: f), cuda=_Anything(), RawKernel=_Anything)
    if need("diffusers"):
        _stub("diffusers")
        _stub("diffusers.models")
        _stub("diffusers.models.attention_processor", USE_PEFT_BACKEND=False)
    if need("pytorch3d"):
        from . import pt3d_shim

        _stub("pytorch3d")
        _stub("pytorch3d.structures", Pointclouds=pt3d_shim.Pointclouds, Meshes=pt3d_shim.Meshes)
        _stub(
            "pytorch3d.renderer",
            compositing=pt3d_shim.compositing,
            TexturesVertex=pt3d_shim.TexturesVertex,
            TexturesUV=_Anything,
            MeshRenderer=_Anything,
            MeshRasterizer=_Anything,
        )
        _stub("pytorch3d.renderer.points", rasterize_points=pt3d_shim.rasterize_points)
        _stub("pytorch3d.renderer.mesh", rasterize_meshes=pt3d_shim.rasterize_meshes)
        _stub("pytorch3d.renderer.mesh.rasterizer", Fragments=pt3d_shim.Fragments)
    try:
        import tqdm.notebook  # noqa: F401
    except Exception:
        import tqdm

        _stub("tqdm.notebook", tqdm=tqdm.tqdm)


_REF = None


def load_reference():
    """Returns a namespace with the reference's hot-path modules, patched to run on CPU."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError("reference tree not present (only available in the build container)")
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import torch

    ap = importlib.import_module("GeoDiffuser.utils.attention_processors")
    wu = importlib.import_module("GeoDiffuser.utils.warp_utils")
    gt = importlib.import_module("GeoDiffuser.utils.generic_torch")
    ash = importlib.import_module("GeoDiffuser.utils.attention_sharing")
    opt = importlib.import_module("GeoDiffuser.utils.optimization")
    loss = importlib.import_module("GeoDiffuser.utils.loss")
    vis = importlib.import_module("GeoDiffuser.utils.vis_utils")
    gen = importlib.import_module("GeoDiffuser.utils.generic")

    # --- patch 1: distance grid on CPU (generic_torch.py:132 default device="cuda")
    _orig_gcd = gt.CoordinateDistances.get_coord_distance

    def _gcd_cpu(self, size, device="cpu"):
        return _orig_gcd(self, size, device="cpu")

    gt.CoordinateDistances.get_coord_distance = _gcd_cpu

    # --- patch 2: warp_grid_edit without the forced .to("cuda") (warp_utils.py:809-812);
    #     body otherwise identical in effect: reshape, SPLATTER(coords, src).
    def _warp_grid_edit_cpu(src, t_coords, padding_mode=None, mode=None, align_corners=False, depth=None,
                            use_softsplat=True, splatting_radius=None, splatting_tau=None,
                            splatting_points_per_pixel=None):
        assert use_softsplat
        if splatting_radius is not None:
            wu.SPLATTER.radius = splatting_radius
        if splatting_tau is not None:
            wu.SPLATTER.tau = splatting_tau
        if splatting_points_per_pixel is not None:
            wu.SPLATTER.points_per_pixel = splatting_points_per_pixel
        b, f, h, w = src.shape
        # RasterizePointsXYsBlending.forward is decorated @torch.autocast("cuda") which is inert on CPU
        # tensors; it negates pts in place on its own fp32 copy (.to(float32) of an fp32 tensor aliases!).
        return wu.SPLATTER(t_coords.reshape(b, h * w, -1).clone(), src.reshape(b, f, h * w))

    wu.warp_grid_edit = _warp_grid_edit_cpu
    ap.warp_grid_edit = _warp_grid_edit_cpu

    # get_transform_coordinates forces .to("cuda") (vis_utils.py:464): provide a CPU twin
    def _get_transform_coordinates_cpu(image, depth, obj_mask=None, transform_in=torch.eye(4), focal_length=550,
                                       return_mesh=False):
        import numpy as np

        K = vis.camera_matrix(focal_length, focal_length, image.shape[1] / 2.0, image.shape[0] / 2.0)
        if np.sum(depth) == 0.5 * (depth.shape[0] * depth.shape[1]):
            depth = np.ones_like(depth) * 0.5
        else:
            depth = depth / (depth.max() + 1e-8)
            depth[depth > 0.95] = 1.0
        mask = (depth < 0.95) * 1.0
        if obj_mask is not None:
            mask = obj_mask * mask
        mask_torch = (torch.tensor(mask)[None, None] >= 0.5) * 1.0
        image_t = torch.from_numpy(image)[None].permute(0, 3, 1, 2)
        with torch.no_grad():
            out = wu.forward_splatting_pytorch3d_warp(
                image_t, torch.from_numpy(depth)[None][None], torch.from_numpy(K)[None], transform_in[None],
                return_coordinates=True, obj_mask=mask_torch, return_mesh=return_mesh)
        return out, depth, mask_torch

    _REF = types.SimpleNamespace(ap=ap, wu=wu, gt=gt, ash=ash, opt=opt, loss=loss, vis=vis, gen=gen,
                                 get_transform_coordinates_cpu=_get_transform_coordinates_cpu)
    return _REF
