#!/usr/bin/env python
"""bench.py -- edits/sec of the geometry-warped shared-attention edit path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one complete edit of BASELINE.json configs[1]: SD-1.5 topology (random init), 512x512, full 50-step DDIM inversion +
3-D rotation edit with latent optimisation (17 optimisation passes + 50 CFG passes), synthetic image/depth/mask.
N > 1: request-level data parallelism, one independent edit stream per rank (weak scaling, no collective on the data path).
`--impl reference`: the CPU oracle (oracle/loop_oracle.py, the reference's algorithm in fp32 with materialised attention maps) on
the host cores, each step a bounded sample (1 inversion eval + 1 optimisation pass + 1 CFG pass) extrapolated to a whole edit.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "edits/sec at 512^2 50-step"
WORKLOAD = "configs[1]: SD-1.5 random-init, 512x512, 50-step DDIM inversion + 3D rotation edit with latent optimisation (17 opt + 50 CFG passes)"
N_OPT, N_CFG, N_INV = 17, 50, 50   # UNet passes per edit with the perform_exp hyper-parameters (large_scale_editor.py:290-299)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                sm.append(float(f[0])); mx.append(float(f[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def kernel_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")) as f:
            return json.load(f)["attn_fwd_sm100_kernel<40> G=3 H=8 N=4096"]["dram_bytes_per_launch"]
    except Exception:
        return None


def attn_flops(tag):
    G, H, N, Nk, d = tag
    return 4.0 * G * H * N * Nk * d


def cpu_reference_sample(threads):
    """One bounded sample of the workload on the host: 1 DDIM-inversion UNet eval + 1 optimisation pass + 1 CFG pass of configs[1]
    through the CPU oracle.  Returns (seconds for the three parts, extrapolated seconds per edit)."""
    from geodiffuser_b200 import synth, unet_sd15
    from geodiffuser_b200.editor import EXP_PARAMS, synthetic_embeddings
    from oracle import loop_oracle as LO

    torch.set_num_threads(threads)
    if not hasattr(cpu_reference_sample, "state"):
        model = unet_sd15.build_model("cpu")
        geo = LO.geometry_inputs("rotate3d", synth)
        text, uncond, x0 = synthetic_embeddings(device="cpu")
        cpu_reference_sample.state = (model.unet.float(), geo, text, uncond, x0)
    unet, geo, text, uncond, x0 = cpu_reference_sample.state
    hp = dict(EXP_PARAMS["geometry_editor"])
    t0 = time.perf_counter()
    with torch.no_grad():
        unet.set_attn_processor(LO.OracleVanillaProcessor())
        unet(torch.cat([x0] * 2), 0, encoder_hidden_states=torch.cat([uncond[:1], text[:1]]))
    t_inv = time.perf_counter() - t0
    gen = torch.Generator().manual_seed(7)
    ddim = [x0] + [torch.randn(1, 4, 64, 64, generator=gen) for _ in range(50)]
    timings = {}
    LO.edit_loop(unet, "rotate3d", geo, text, uncond, ddim[-1], ddim, hp, 50, step_limit=1, timings=timings)
    t_opt, t_cfg = timings["opt"][0], timings["cfg"][0]
    return (t_inv, t_opt, t_cfg), N_INV * t_inv + N_OPT * t_opt + N_CFG * t_cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count()
    per_edit = []
    for i in range(args.warmup + args.steps):
        parts, t_edit = cpu_reference_sample(threads)
        if i >= args.warmup:
            per_edit.append(t_edit)
        if i == 0 and t_edit / 100.0 * (args.warmup + args.steps) > 400:   # keep the whole run within minutes on small hosts
            per_edit = [t_edit]
            break
    t = float(np.mean(per_edit))
    value = 1.0 / t
    sample = "1 DDIM-inversion UNet eval + 1 optimisation pass (fwd+bwd) + 1 CFG pass of configs[1], extrapolated x(50, 17, 50)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "edits/s", "n_gpus": args.gpus, "steps": len(per_edit),
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "edits/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "edits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: geodiffuser_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from geodiffuser_b200 import _lib, editor, unet_sd15

    _lib.lib()  # fail loudly if the extension is missing
    dev = torch.device("cuda", local)
    model = unet_sd15.build_model(dev)
    req = editor.synthetic_request("rotate3d", seed=1234 + rank)
    api = lambda: editor.perform_geometric_edit(model, req["depth"], req["image_mask"], req["transform_in"], req["text_embeddings"],
                                                req["uncond_embeddings"], req["x0"], req["edit_type"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        out, h2d, d2h = api()
    assert torch.isfinite(out).all()

    # (1) device-resident inputs: staging happens before the timed region
    staged, _ = editor.stage_inputs(req["depth"], req["image_mask"], req["text_embeddings"], req["uncond_embeddings"], req["x0"], dev)
    resident = lambda: editor.run_edit(model, staged, req["transform_in"], req["edit_type"])
    sampler = ClockSampler(local)
    sampler.start()
    _lib.profile_begin(["gd_attn_fwd_sm100", "gd_attn_fwd_generic", "gd_attn_bwd", "gd_attn_bwd_sm100", "gd_attn_probs", "gd_corr_max_partial"])
    l0 = _lib.LAUNCHES
    ms_value = timed(resident, args.steps)
    launches = _lib.LAUNCHES - l0
    prof = _lib.profile_end()
    clocks = sampler.stop()
    # (2) end to end through the public API with host buffers (H2D of the request + D2H of the result inside the timed region)
    ms_e2e = timed(api, args.steps)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        value = world * args.steps / (ms_value / 1e3)
        e2e = world * args.steps / (ms_e2e / 1e3)
        # roofline of the dominant kernel: the tcgen05 forward at the 64^2 level (N = 4096, d = 40), timed live above
        by_kernel = {k: sum(ms for ms, _ in v) for k, v in prof.items()}
        dom = [(ms, tag) for ms, tag in prof.get("gd_attn_fwd_sm100", []) if tag and tag[2] == 4096]
        roof = None
        if dom:
            fl = sum(attn_flops(t) for _, t in dom)
            ms = sum(m for m, _ in dom)
            ach = fl / (ms * 1e-3) / 1e12
            pk = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
            roof = {"kernel": "attn_fwd_sm100_kernel<40> (N=4096, d=40)", "bound": "tensor", "achieved": ach, "peak": pk, "unit": "TFLOP/s",
                    "frac": ach / pk, "traffic": kernel_traffic(), "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
                    "launches": len(dom), "avg_launch_ms": ms / len(dom), "algorithmic_flops_per_launch": fl / len(dom),
                    "share_of_step_ms": {k: round(v / args.steps, 3) for k, v in by_kernel.items()}}
            bw = [(ms, tag) for ms, tag in prof.get("gd_attn_bwd_sm100", []) if tag and tag[1] == 4096]
            if bw:   # the matching backward (dQ) at the same level, same peak: 6*H*N^2*d algorithmic FLOP per launch
                flb = sum(6.0 * t[0] * t[1] * t[1] * t[2] for _, t in bw)
                msb = sum(m for m, _ in bw)
                roof["backward"] = {"kernel": "attn_bwd_sm100_kernel<40> (N=4096, d=40)", "achieved": flb / (msb * 1e-3) / 1e12, "unit": "TFLOP/s",
                                    "frac": flb / (msb * 1e-3) / 1e12 / pk, "launches": len(bw), "avg_launch_ms": msb / len(bw)}
        line = {"metric": METRIC, "value": value, "unit": "edits/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "parallelism": f"request-level dp{world} (independent edits, no collective)",
                           "l2": "no flush needed: every UNet pass streams 1.7 GB of weights + activations (> 126 MB L2) between repeats"},
                "e2e": {"value": e2e, "unit": "edits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count()
            parts, t_edit = cpu_reference_sample(threads)
            line["cpu_baseline"] = {"value": 1.0 / t_edit, "unit": "edits/s", "cores": threads, "kind": "port",
                                    "sample": "1 DDIM-inversion UNet eval (%.1fs) + 1 optimisation pass (%.1fs) + 1 CFG pass (%.1fs) of "
                                              "configs[1] on the CPU oracle, extrapolated x(50, 17, 50)" % parts}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
