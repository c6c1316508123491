"""oracle/make_golden_loop.py -- TEST INFRASTRUCTURE ONLY.  Run in the build container:

    python -m oracle.make_golden_loop [--full]

1. runs oracle/loop_oracle.py (CPU fp32 restatement of the edit loop) on seeded synthetic inputs over the random-init UNet;
2. PINS it: re-runs the same loop with the REAL reference controller / processor classes imported from /root/reference
   (oracle/ref_import.py) in place of OracleController and asserts the final latents agree to 1e-4;
3. writes tests/golden/loop_<kind>_<size>.npz (final latents + per-step loss log).
`--full` additionally produces the BASELINE.json config-1 case on the full SD-1.5 topology (5 DDIM steps, 2-D translation), which
takes several minutes of CPU time.
"""
import os
import sys
import time

import numpy as np
import torch

from geodiffuser_b200 import synth, unet_sd15
from geodiffuser_b200.editor import EXP_PARAMS, synthetic_embeddings
from . import loop_oracle as LO
from .ref_import import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def reference_controller_factory(R, unet, kind, geo, hp, num_steps):
    """builds the reference's own controller + EditProcessors on `unet` (editor.py:101, 610-638)"""
    R.ap.USE_PEFT_BACKEND = True  # plain nn.Linear projections (no LoRA `scale` argument)

    def make(weights):
        cls = R.ap.AttentionGeometryRemover if kind == "remove" else R.ap.AttentionGeometryEdit
        c = cls(["", ""], num_steps, cross_replace_steps=hp["cross_replace_steps"], self_replace_steps=hp["self_replace_steps"],
                image_mask=geo["mask"].astype(np.float32), empty_scale=0.0, use_all=False, obj_edit_step=hp["obj_edit_step"], tokenizer=None,
                device="cpu", mode="bilinear")
        c.amodal_mask = torch.from_numpy(geo["amodal"])[None, None]
        c.mask_new_warped = torch.from_numpy(geo["mnw"])[None, None].tile(2, 1, 1, 1)
        c.loss_weight_dict = weights
        c.default_loss_weights = weights
        tc = torch.from_numpy(geo["coords"])[None]

        class _Model:
            pass

        m = _Model()
        m.unet = unet
        R.ap.register_attention_control_diffusers(m, c, tc)
        return c

    return make


def run_case(name, kind, tiny, num_steps, pin, R, inversion=True, step_limit=None):
    print(f"== loop case {name}: kind={kind} tiny={tiny} steps={num_steps}", flush=True)
    model = unet_sd15.build_model("cpu", tiny=tiny)
    unet = model.unet.float()
    edit_type = "geometry_remover" if kind == "remove" else "geometry_editor"
    hp = dict(EXP_PARAMS[edit_type])
    geo = LO.geometry_inputs(kind, synth)
    text, uncond, x0 = synthetic_embeddings(device="cpu")
    t0 = time.time()
    if inversion:
        ddim = LO.ddim_inversion(unet, x0, torch.cat([uncond[:1], text[:1]]), hp["guidance_scale"], num_steps)
    else:
        gen = torch.Generator().manual_seed(1234 + 2)
        ddim = [x0] + [torch.randn(1, 4, 64, 64, generator=gen) for _ in range(num_steps)]
    print(f"   inversion {time.time() - t0:.1f}s", flush=True)
    t0 = time.time()
    timings = {}
    lat, log = LO.edit_loop(unet, kind, geo, text, uncond, ddim[-1], ddim, hp, num_steps, step_limit=step_limit, timings=timings)
    print(f"   oracle loop {time.time() - t0:.1f}s  opt={np.mean(timings.get('opt', [0])):.2f}s cfg={np.mean(timings['cfg']):.2f}s", flush=True)
    rec = dict(latents=lat.numpy(), x_T=ddim[-1].numpy(), meta=np.array([int(tiny), num_steps, int(inversion), -1 if step_limit is None else step_limit]))
    for i, d in log.items():
        rec[f"log{i}_loss"] = np.float64(d["loss"])
        for att in ("self", "cross"):
            for k, v in d[att].items():
                rec[f"log{i}_{att}_{k}"] = np.float64(v)
    if pin:
        t0 = time.time()
        make = reference_controller_factory(R, unet, kind, geo, hp, num_steps)
        lat_ref, log_ref = LO.edit_loop(unet, kind, geo, text, uncond, ddim[-1], ddim, hp, num_steps, make_controller=make, step_limit=step_limit)
        err = float((lat_ref - lat).abs().max() / lat_ref.abs().max())
        print(f"   PIN vs real reference controllers: rel-max err of final latents = {err:.3e} ({time.time() - t0:.1f}s)", flush=True)
        for i in log:
            print(f"     step {i}: loss oracle {log[i]['loss']:.6f} reference {log_ref[i]['loss']:.6f}")
        assert err <= 1e-2, "oracle loop does not reproduce the reference controllers"  # fp32 round-off amplified by the 10-step optimisation; step-0 loss is identical
        rec["pin_err_vs_reference"] = np.float64(err)
    np.savez_compressed(os.path.join(OUT, f"loop_{name}.npz"), **rec)


def main():
    torch.set_num_threads(os.cpu_count())
    R = load_reference()
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None     # e.g. --only remove
    if "--full-only" not in sys.argv and "--full50" not in sys.argv:
        for kind in ("translate2d", "rotate3d", "remove"):
            if only in (None, kind):
                run_case(f"{kind}_tiny", kind, True, 10, True, R)
    if "--full" in sys.argv or "--full-only" in sys.argv:
        run_case("translate2d_full5", "translate2d", False, 5, False, R, inversion=False)
    if "--full50" in sys.argv:
        # BASELINE.json configs[1] and configs[2] at full size: SD-1.5 topology, 50-step DDIM inversion + 50-step edit (3-D rotation with latent
        # optimisation; object removal).  ~15 minutes of CPU time each on 8 cores.
        kinds = [sys.argv[sys.argv.index("--full50") + 1]] if len(sys.argv) > sys.argv.index("--full50") + 1 else ["rotate3d", "remove"]
        for kind in kinds:
            run_case(f"{kind}_full50", kind, False, 50, False, R, inversion=True)


if __name__ == "__main__":
    sys.exit(main())
