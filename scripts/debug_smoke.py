"""diagnostic for __graft_entry__.smoke(): where does the dQ difference against the oracle sit?"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import _lib, geometry as G, synth
from geodiffuser_b200.attention_processors import AttentionGeometryEdit
from oracle import geodiff_oracle as O

image, depth, mask, T = synth.edit_inputs("translate2d")
g = G.correspondence_field(depth.copy(), mask.copy(), T)
ref = O.corr_build(depth.copy(), mask.copy(), T)
S, H, d = 32, 2, 16
amodal = G.torch_erode(G.mesh_mask(g["coords"], g["mask"])[None, None])
idx512, _, dd = O.splat_index(ref["coords"][None])
mnw = O.binarize(O.splat_composite(mask.astype(np.float32)[None, None], idx512, dd))[0, 0]
masks = O.build_masks(mask, mnw, O.erode3(O.mesh_mask(ref["coords"], ref["mask"])), S)
for seed in (101, 102, 103):
    c = AttentionGeometryEdit(["", ""], 50, cross_replace_steps={"default_": 0.95}, self_replace_steps=0.95, image_mask=mask.astype(np.float32),
                              empty_scale=0.0, use_all=False, obj_edit_step=0.9, device="cuda")
    c.num_att_layers, c.amodal_mask, c.use_cfg, c.coords_base, c.coords_edit = 32, amodal, False, (0, 1), (1, 2)
    q, k, v = synth.qkv(seed, 2, H, S * S, S * S, d)
    qc, kc, vc = (torch.from_numpy(a).cuda().requires_grad_(True) for a in (q, k, v))
    out = c(qc, kc, vc, False, "down", transform_coords=g["coords"][None], scale=d ** -0.5)
    (gq,) = torch.autograd.grad(c.loss, [qc])
    qo, ko, vo = (torch.from_numpy(a).requires_grad_(True) for a in (q, k, v))
    res = O.edit_layer(qo, ko, vo, False, d ** -0.5, H, (0, 1), (1, 2), masks, O.resize_coords(ref["coords"], S), False, True)
    (gq_ref,) = torch.autograd.grad(res["loss"], [qo])
    a, b = gq.cpu()[H:], gq_ref[H:]
    err_rows = (a - b).abs().amax(-1) / b.abs().max()
    bad = (err_rows > 2e-2)
    print(f"seed {seed}: dq relerr {float(err_rows.max()):.3e}; rows over 2e-2: {int(bad.sum())} of {bad.numel()}; median row err {float(err_rows.median()):.2e}; "
          f"terms ours {[round(float(x), 5) for x in c.loss_log_dict['self'].values()]}", flush=True)
    hh, rr = torch.nonzero(bad, as_tuple=True)
    inp = torch.from_numpy(np.asarray(masks["mask_1_empty"]).reshape(-1) > 0.5)
    print("   bad rows in inpaint set:", int(inp[rr].sum()), "of", int(bad.sum()), "; first bad rows", rr[:8].tolist(), flush=True)
