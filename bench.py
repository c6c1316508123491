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


def bench_config(n):
    """the same dict in both arms (the driver compares them)"""
    return {"workload": WORKLOAD, "parallelism": f"request-level dp{n} (independent edits, no collective; 4 edit lanes per GPU in the GPU arm)",
            "l2": "no flush needed: every UNet pass streams 1.7 GB of weights + activations (> 126 MB L2) between repeats"}


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


def kernel_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed `ncu --set full` capture of this round"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")) as f:
            return json.load(f)[key]["dram_bytes_per_launch"]
    except Exception:
        return None


def kernel_rooflines(dev, M64, M32, peaks, iters=60):
    """The dominant kernels of the path, timed in this process at the shapes the edit launches them with (the edit itself replays CUDA graphs,
    inside which single launches cannot carry events).  `iters` back-to-back launches, cycling through enough distinct operand sets that
    consecutive launches never find their inputs in the 126 MB L2, are recorded into ONE CUDA graph; the graph is replayed once to warm up and
    once between two CUDA events on the launching stream.  (Launched one by one from Python the host needs 30-45 us per call -- ctypes plus
    3 G tensor-map encodes -- which is longer than the 32^2-token kernels themselves and used to be counted as kernel time.)
    Algorithmic work per launch: SURVEY 8(d) (forward 4*G*H*N^2*d, backward 6*H*N^2*d, correlation 2*H*M*N^2, removal dQ rows 2*H*M*N*d).
    Peak: the BURST bf16 figure of MEASURED_PEAKS.json (kernel timed alone); `frac_sustained` uses the sustained one."""
    from geodiffuser_b200 import _lib
    from geodiffuser_b200._lib import call, ptr, stream

    H = 8
    burst, sust = peaks["bf16_tflops"], peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    gen = torch.Generator(device=dev).manual_seed(4321)
    mk = lambda *shape, s=1.5: (torch.randn(*shape, device=dev, generator=gen) * s).bfloat16()

    def timed(fns):
        for i in range(max(3, len(fns))):
            fns[i % len(fns)]()
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                fns[i % len(fns)]()
        g.replay()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / iters

    def entry(name, kernel, flops, ms, **extra):
        ach = flops / (ms * 1e-3) / 1e12
        return dict(kernel=kernel, entry=name, bound="tensor", achieved=ach, peak=burst, unit="TFLOP/s", frac=ach / burst, frac_sustained=ach / sust,
                    avg_launch_ms=ms, launches=iters, algorithmic_flops_per_launch=flops, **extra)

    out = []
    for G, N, d in ((3, 4096, 40), (4, 4096, 40), (2, 4096, 40), (3, 1024, 80), (4, 1024, 80)):
        nsets = max(2, int(140e6 // ((G + 2) * H * N * d * 2)) + 1)
        fns = []
        for _ in range(nsets):
            qs, k, v = [mk(H, N, d) for _ in range(G)], mk(H, N, d), mk(H, N, d)
            O = torch.empty(G, H, N, d, device=dev, dtype=torch.float32)
            L = torch.empty(G, H, N, device=dev, dtype=torch.float32)
            a = (_lib.ptr_array(qs), _lib.ptr_array([k] * G), _lib.ptr_array([v] * G), _lib.ptr_array([O[i] for i in range(G)]),
                 _lib.ptr_array([L[i] for i in range(G)]))
            fns.append(lambda a=a, keep=(qs, k, v, O, L), G=G, N=N, d=d: call("gd_attn_fwd_sm100", *a, None, G, H, N, N, d, float(d ** -0.5), None, 0, stream()))
        ms = timed(fns)
        out.append(entry("gd_attn_fwd_sm100", f"attn_fwd_sm100_kernel<{d}> G={G} H={H} N={N}", 4.0 * G * H * N * N * d, ms, operand_sets=nsets,
                         traffic=kernel_traffic(f"attn_fwd_sm100_kernel<{d}> G={G} H={H} N={N}")))
        del fns
    # backward: the flash term by the tcgen05 kernel; the removal term of the M inpaint rows by its own contraction (gd_removal_dq_rows)
    for N, d in ((4096, 40), (1024, 80)):
        nsets = max(2, int(140e6 // (4 * H * N * d * 2)) + 1)
        fns = []
        ld = (N + 7) // 8 * 8
        for _ in range(nsets):
            q, k, v, do = mk(H, N, d), mk(H, N, d), mk(H, N, d), mk(H, N, d, s=1.0)
            L = (torch.randn(H, N, device=dev, generator=gen) * 0.1 + 8.0).float()
            delta = torch.randn(H, N, device=dev, generator=gen).float() * 0.01
            dq = torch.empty(H, N, d, device=dev, dtype=torch.float32)
            fns.append(lambda q=q, k=k, v=v, do=do, L=L, delta=delta, dq=dq, N=N, d=d, ld=ld:
                       call("gd_attn_bwd_sm100", ptr(q), ptr(k), ptr(v), ptr(do), ptr(L), ptr(delta), None, None, None, ld, 0, ptr(dq), H, N, d,
                            float(d ** -0.5), None, 0, 0, stream()))
        ms = timed(fns)
        out.append(entry("gd_attn_bwd_sm100", f"attn_bwd64_sm100_kernel<{d}> H={H} N={N}", 6.0 * H * N * N * d, ms, operand_sets=nsets,
                         traffic=kernel_traffic(f"attn_bwd64_sm100_kernel<{d}> H={H} N={N}")))
        del fns
    # removal loss: correlation of attention maps (base map recomputed in TMEM) and the removal rows of dQ, at this edit's inpaint-row counts
    for N, d, M in ((4096, 40, M64), (4096, 40, 410), (1024, 80, M32)):
        if M <= 0:
            continue
        ld = (N + 7) // 8 * 8
        nsets = max(2, int(140e6 // (2 * H * N * d * 2 + H * M * ld * 2)) + 1)
        corr, rowsk = [], []
        for _ in range(nsets):
            qb, kb = mk(H, N, d), mk(H, N, d)
            lse = (torch.randn(H, N, device=dev, generator=gen) * 0.1 + 8.0).float()
            a_e = (torch.rand(H, M, ld, device=dev, generator=gen) * (2.0 / N)).bfloat16()
            rows = (torch.arange(M, device=dev) + N // 3).int()
            m_in = torch.zeros(N, device=dev)
            m_in[rows.long()] = 1.0
            m_bg = (1.0 - m_in).contiguous()
            part = torch.empty(H, N // 32, M, 4, device=dev)
            dq = torch.zeros(H, N, d, device=dev)
            gs = torch.ones(1, device=dev)
            corr.append(lambda qb=qb, kb=kb, lse=lse, a_e=a_e, m_in=m_in, m_bg=m_bg, part=part, N=N, d=d, M=M, ld=ld:
                        call("gd_removal_corr_sm100", ptr(qb), ptr(kb), ptr(lse), ptr(a_e), H, M, N, d, float(d ** -0.5), ld, None, ptr(m_in), ptr(m_bg),
                             ptr(part), stream()))
            rowsk.append(lambda a_e=a_e, kb=kb, rows=rows, gs=gs, dq=dq, N=N, d=d, M=M, ld=ld:
                         call("gd_removal_dq_rows", ptr(a_e), ptr(kb), ptr(rows), ptr(gs), ptr(dq), H, M, N, N, d, float(d ** -0.5), ld, None, 0, stream()))
        out.append(entry("gd_removal_corr_sm100", f"removal_corr_sm100_kernel<{d}> H={H} N={N} M={M}", 2.0 * H * M * N * N, timed(corr), operand_sets=nsets,
                         traffic=kernel_traffic(f"removal_corr_sm100_kernel<{d}> H={H} N={N} M={M}")))
        out.append(entry("gd_removal_dq_rows", f"removal_dq_rows_kernel<{d}> H={H} N={N} M={M}", 2.0 * H * M * N * d, timed(rowsk), operand_sets=nsets, traffic=None,
                         note="mma.sync; latency-bound by design (0.2 GFLOP): reported for completeness"))
        del corr, rowsk
    return out


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
            "data": "synthetic", "config": bench_config(args.gpus),
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
    from geodiffuser_b200 import _lib, editor, runner, unet_sd15

    _lib.lib()  # fail loudly if the extension is missing
    dev = torch.device("cuda", local)
    model = unet_sd15.build_model(dev)
    # `lanes` independent edits in flight per GPU (runner.EditWorkers: one thread + stream + model replica over shared weights per lane)
    workers = runner.EditWorkers(model, args.lanes)
    req = editor.synthetic_request("rotate3d", seed=1234 + rank)
    api = lambda m: editor.perform_geometric_edit(m, req["depth"], req["image_mask"], req["transform_in"], req["text_embeddings"],
                                                  req["uncond_embeddings"], req["x0"], req["edit_type"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        """K edits (steps), each taken by the first free lane; device time between two events on the launching stream, which waits for every lane"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = workers.map(lambda m, _: fn(m), list(range(K)))
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), res

    if args.workload == "mixed64":
        return run_mixed64(args, world, rank, dev, workers, barrier)
    res = workers.map_every_lane(lambda m, _: api(m), list(range(max(args.warmup, 3))))     # every lane warms up (graphs, cuDNN autotune)
    out, h2d, d2h = res[-1]
    assert torch.isfinite(out).all()

    # (1) device-resident inputs: staging happens before the timed region
    staged, _ = editor.stage_inputs(req["depth"], req["image_mask"], req["text_embeddings"], req["uncond_embeddings"], req["x0"], dev)
    resident = lambda m: editor.run_edit(m, staged, req["transform_in"], req["edit_type"])
    sampler = ClockSampler(local)
    sampler.start()
    l0, f0 = _lib.LAUNCHES, _lib.FLOPS
    ms_value, _ = timed(resident, args.steps)
    launches = _lib.LAUNCHES - l0
    flops_per_edit = (_lib.FLOPS - f0) / args.steps     # algorithmic attention-path FLOP (graph replays re-count what they captured)
    clocks = sampler.stop()
    # (2) end to end through the public API with host buffers (H2D of the request + D2H of the result inside the timed region)
    ms_e2e, _ = timed(api, args.steps)

    # (3) roofline of the path's dominant kernels at the shapes of this edit (eager, outside the timed regions)
    roofs = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        ctrl, tc = editor.make_controller(model, staged, req["transform_in"], req["edit_type"], dict(editor.EXP_PARAMS[req["edit_type"]]))
        M64, M32 = (ctrl._get_cache(S, tc, dev).M for S in (64, 32))
        del ctrl
        roofs = kernel_rooflines(dev, M64, M32, peaks)
        if not roofs:
            raise SystemExit("bench.py: the roofline leg produced nothing")

    if rank == 0:
        value = world * args.steps / (ms_value / 1e3)
        e2e = world * args.steps / (ms_e2e / 1e3)
        roof = dict(roofs[0])
        roof["peak_source"] = peak_src + ": burst bf16 (kernel timed alone, back to back); frac_sustained = against the sustained figure"
        roof["other_kernels"] = roofs[1:]
        pk_s = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        # the whole edit against the attention roofline (north_star): algorithmic attention-path FLOP of one edit / step time / peak
        roof["edit_level"] = {"attention_path_flops_per_edit": flops_per_edit,
                              "attention_roofline_frac_of_edit": flops_per_edit / (ms_value / args.steps * 1e-3) / 1e12 / pk_s,
                              "peak": pk_s, "note": "whole-edit time (UNet body included) against the sustained bf16 peak; per GPU, identical at every N (weak scaling)"}
        line = {"metric": METRIC, "value": value, "unit": "edits/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": bench_config(world),
                "e2e": {"value": e2e, "unit": "edits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "lanes_per_gpu": args.lanes}
        assert line["roofline"] is not None and line["roofline"]["frac"] > 0
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count()
            parts, t_edit = cpu_reference_sample(threads)
            line["cpu_baseline"] = {"value": 1.0 / t_edit, "unit": "edits/s", "cores": threads, "kind": "port",
                                    "sample": "1 DDIM-inversion UNet eval (%.1fs) + 1 optimisation pass (%.1fs) + 1 CFG pass (%.1fs) of "
                                              "configs[1] on the CPU oracle, extrapolated x(50, 17, 50)" % parts}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_mixed64(args, world, rank, dev, workers, barrier):
    """BASELINE.json configs[4]: 64 independent mixed edits -- 16 each of 2-D translation, 3-D rotation, object removal and 3-D rotation at
    768 x 768 -- in the seed-1234 shuffle of SURVEY 8(d), dealt round-robin to the ranks (runner.shard_round_robin); inside a rank every request is taken by the first free
    edit lane, the 768^2 requests first.  Every request goes through the public API with host buffers (H2D + D2H inside the timed region).  One JSON line (rank 0)."""
    import torch.distributed as dist
    from geodiffuser_b200 import editor, runner

    kinds = ["translate2d", "rotate3d", "remove", "rotate3d@768"] * 16
    np.random.RandomState(1234).shuffle(kinds)
    make = lambda i, k: editor.synthetic_request(k.split("@")[0], seed=1234 + i, image_size=768 if k.endswith("@768") else 512)
    mine = runner.shard_round_robin(len(kinds), rank, world)
    reqs = [make(i, kinds[i]) for i in mine]
    api = lambda m, r: editor.perform_geometric_edit(m, r["depth"], r["image_mask"], r["transform_in"], r["text_embeddings"], r["uncond_embeddings"],
                                                     r["x0"], r["edit_type"])
    # warm-up: every lane sees every kind twice (inversion / CFG graphs, cuDNN autotune per lane thread, first-edit eager optimisation pass)
    warm = [make(1000 + j, k) for k in ("translate2d", "rotate3d", "remove", "rotate3d@768") for j in range(2)]
    workers.map_every_lane(api, warm)
    # longest first: a 768^2 edit costs ~2.3 x a 512^2 one, and the lanes of a rank take requests as they become free (runner.EditWorkers)
    reqs = [reqs[i] for i in runner.longest_first([2.3 if kinds[m].endswith("@768") else 1.0 for m in mine])]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = workers.map(api, reqs)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert all(torch.isfinite(r[0]).all() for r in res)
    if rank == 0:
        t = float(ms) / 1e3
        line = {"metric": METRIC, "value": len(kinds) / t, "unit": "edits/s", "n_gpus": world, "steps": len(kinds), "warmup": len(warm) * args.lanes, "ms_per_step": t * 1e3 / len(kinds),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "configs[4]: 64 independent mixed edits (16 x 2-D translation, 16 x 3-D rotation, 16 x removal, 16 x 3-D rotation at 768^2), "
                                       "seed-1234 shuffle, sharded round-robin over the GPUs", "parallelism": f"request-level dp{world}, {args.lanes} edit lanes per GPU",
                           "edits_per_rank": len(mine)},
                "e2e": {"value": len(kinds) / t, "unit": "edits/s", "h2d_bytes_per_step": int(np.mean([r[1] for r in res])), "d2h_bytes_per_step": int(np.mean([r[2] for r in res]))},
                "lanes_per_gpu": args.lanes, "seconds": t}
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
    ap.add_argument("--workload", default="config1", choices=["config1", "mixed64"],
                    help="config1 = BASELINE configs[1] (the metric's configuration); mixed64 = configs[4], the 64-edit mixed sweep")
    ap.add_argument("--lanes", type=int, default=4, help="independent edits in flight per GPU (1 = one edit at a time)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
