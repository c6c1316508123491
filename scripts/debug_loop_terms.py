"""debug (GPU box): first optimisation pass of the tiny loop -- per layer, our loss terms vs the CPU oracle layer on the SAME q,k,v."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from geodiffuser_b200 import unet_sd15, editor, functional as Fn
from oracle import geodiff_oracle as O

kind = sys.argv[1] if len(sys.argv) > 1 else "translate2d"
model = unet_sd15.build_model("cuda", tiny=True)
rec = []
orig = Fn.shared_attention_layer

def hook(q, k, v, spec):
    out, loss, terms = orig(q, k, v, spec)
    if spec.with_loss and len(rec) < 12:
        rec.append((q.detach().float().cpu(), k.detach().float().cpu(), v.detach().float().cpu(), spec, terms.detach().cpu().clone(), out.detach().float().cpu()))
    return out, loss, terms

Fn.shared_attention_layer = hook
import geodiffuser_b200.attention_processors as AP
AP.Fn.shared_attention_layer = hook
try:
    editor.perform_synthetic_edit(model, kind, num_ddim_steps=10, return_log=True, optimize_steps=0.05)
except Exception as e:
    print("edit raised", e)
for (q, k, v, spec, terms, out) in rec:
    c = spec.cache
    masks = {kk: vv.cpu().numpy() for kk, vv in c.masks.items()}
    N = q.shape[1]; S = c.S
    if spec.kind == "edit":
        idx_coords = None
        # coords at S from the controller cache are not kept in the ResolutionCache: rebuild through the oracle from the synthetic inputs
        from geodiffuser_b200 import synth
        image, depth, mask, T = synth.edit_inputs(kind)
        ref = O.corr_build(depth.copy(), mask.copy(), T)
        coords_S = O.resize_coords(ref["coords"], S)
        res = O.edit_layer(q, k, v, spec.is_cross, spec.scale, spec.heads, spec.cb, spec.ce, masks, coords_S, False, spec.blend, weights={"self": spec.weights, "cross": spec.weights})
        t = res["terms"]
        print(f"S={S} cross={spec.is_cross} d={q.shape[2]} ours sim {terms[0]:.6f} mov {terms[1]:.6f} rem {terms[2]:.6f} smo {terms[3]:.6f} amo {terms[4]:.6f} | "
              f"oracle sim {float(t['sim']):.6f} mov {float(t['movement']):.6f} rem {float(t['removal']):.6f} smo {float(t['smoothness']):.6f} amo {float(t['amodal']):.6f}")
        h = spec.heads
        e_o, r_o = res["edit_out"], res["replace_out"]
        d_or = (e_o - r_o).abs().sum(-1).mean(0).reshape(S, S)
        # q identical?
        print("   q_edit==q_base:", bool(torch.equal(q[:h], q[h:2*h])), " oracle |e-r| rowsum: max %.4f at %s ; bg-rows mean %.6f" % (float(d_or.max()), np.unravel_index(int(d_or.argmax()), (S, S)), float((d_or * torch.from_numpy(masks['mask_wo_edit'])).sum() / masks['mask_wo_edit'].sum())))
        print("   out relerr", float((out - res["out"]).abs().max() / res["out"].abs().max()))
