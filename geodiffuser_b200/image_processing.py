"""Post-processing tail of an edit, mirror of /root/reference/GeoDiffuser/utils/image_processing.py:24-77 and of the
image branch of editor.py:659-690: masked histogram matching of the edited image against the (forward-warped) input image.  The reference
does this in numpy on the host; here the histograms, the np.interp look-up table and the remap run on the device (csrc/postprocess.cu),
bit-identical to the reference's float64 result.  VAE decoding (diffusion.py:62-68) needs pretrained weights and stays out of scope."""
import numpy as np
import torch

from . import geometry
from ._lib import call, ptr, stream


def _dev_u8(a, device):
    if torch.is_tensor(a):
        return a.to(device=device, dtype=torch.uint8).contiguous()
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise TypeError("masked_histogram_matching expects uint8 images (as np.bincount in the reference does)")
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def _dev_f32(a, device):
    if torch.is_tensor(a):
        return a.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(device)


def masked_histogram_matching(source, template, mask=None, mask_source=None, device="cuda"):
    """image_processing.py:68-77.  source, template (H,W,C) uint8; mask / mask_source (H,W), > 0.5 selects (mask_source defaults to mask,
    mask to all ones).  Returns (H,W,C) float64: numpy in -> numpy out, tensor in -> CUDA tensor out."""
    as_numpy = not torch.is_tensor(source)
    s, t = _dev_u8(source, device), _dev_u8(template, device)
    if s.shape != t.shape or s.dim() != 3:
        raise ValueError("source and template must be (H, W, C) images of the same shape")
    H, W, C = s.shape
    m = torch.ones(H, W, device=s.device) if mask is None else _dev_f32(mask, s.device)
    ms = m if mask_source is None else _dev_f32(mask_source, s.device)
    if float((m > 0.5).sum()) == 0 or float((ms > 0.5).sum()) == 0:
        raise ValueError("empty mask: the reference divides by the masked pixel count")
    counts = torch.empty(C, 2, 256, device=s.device, dtype=torch.int32)
    lut = torch.empty(C, 256, device=s.device, dtype=torch.float64)
    out = torch.empty(H, W, C, device=s.device, dtype=torch.float64)
    call("gd_masked_histogram_match", ptr(s), ptr(t), ptr(m), ptr(ms), H * W, C, ptr(counts), ptr(lut), ptr(out), stream())
    return out.cpu().numpy() if as_numpy else out


def postprocess_edited_image(edited_image, image, t_coords, mask_new_warped, image_mask, edit_type="geometry_editor", device="cuda"):
    """editor.py:659-690.  edited_image, image (H,W,3) uint8; t_coords (H,W,3) correspondence field; mask_new_warped (H,W) (the controller's
    binarised warped object mask, [0,0] slice); image_mask (H,W).  Returns the histogram-matched edited image, float64 (H,W,3)."""
    img = _dev_u8(image, device)
    if edit_type == "geometry_editor":
        tc = t_coords if torch.is_tensor(t_coords) else torch.from_numpy(np.asarray(t_coords, dtype=np.float32))
        src = (img.permute(2, 0, 1)[None].float() / 255.0)
        warped = geometry.warp_grid_edit(src, tc.to(img.device)[None].float())              # editor.py:663
        p_image = (warped[0].permute(1, 2, 0) * 255.0).to(torch.uint8)                       # .astype("uint8"): truncation
        mask_edit = _dev_f32(mask_new_warped, img.device)
        mask_im = _dev_f32(image_mask, img.device)
        mask_changed = ((mask_edit + mask_im) > 0.5).float()
        mask_wo_edit = ((1.0 - mask_changed) > 0.5).float()
        p_new = (mask_wo_edit[..., None].double() * img.double() + mask_edit[..., None].double() * p_image.double()).to(torch.uint8)
        mask_source = ((mask_edit + mask_wo_edit) > 0.5).float()
        out = masked_histogram_matching(_dev_u8(edited_image, device), p_new, mask_source, mask_source)
    elif edit_type == "geometry_stitch":
        out = masked_histogram_matching(_dev_u8(edited_image, device), img, 1.0 - _dev_f32(mask_new_warped, img.device))
    elif edit_type == "geometry_remover":
        out = masked_histogram_matching(_dev_u8(edited_image, device), img, 1.0 - _dev_f32(image_mask, img.device))
    else:
        raise ValueError(edit_type)
    return out.cpu().numpy() if not torch.is_tensor(edited_image) else out
