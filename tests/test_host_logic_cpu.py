"""Host-side mirror of the reference interface: controller state machine, loss-weight aliasing, adaptive schedule, scheduler tables,
processor registration.  No CUDA needed."""
import numpy as np
import pytest
import torch

from geodiffuser_b200 import attention_processors as AP
from geodiffuser_b200 import diffusion, optimization, unet_sd15


def make(kind="edit"):
    cls = AP.AttentionGeometryEdit if kind == "edit" else AP.AttentionGeometryRemover
    return cls(["", ""], 50, cross_replace_steps={"default_": 0.95}, self_replace_steps=0.95, image_mask=np.zeros((512, 512), np.float32),
               empty_scale=0.0, use_all=False, obj_edit_step=0.9, device="cpu")


def test_controller_attributes_match_reference_contract():
    c = make()
    for attr in ("loss", "loss_log_dict", "loss_weight_dict", "mask_new_warped", "cur_step", "amodal_mask", "image_mask", "store_attention_maps",
                 "attention_store", "coords_base", "coords_edit", "use_cfg", "num_self_replace", "masks_cache_dict", "batch_size"):
        assert hasattr(c, attr), attr
    assert c.num_self_replace == (0, 47) and c.image_mask.shape == (2, 512, 512)
    assert set(c.loss_log_dict["self"]) == {"sim", "movement", "removal", "smoothness"}       # no `amodal` key (reference :626-630)
    assert set(make("remove").loss_log_dict["cross"]) == {"sim", "removal", "smoothness"}
    assert c.default_loss_weights["self"]["removal"] == 1.67 and make("remove").default_loss_weights["self"]["removal"] == 3.6


def test_loss_weight_aliasing_quirk():
    """initialize_default_loss_weights() re-binds the SAME dict, so the adaptive schedule's reset is a no-op"""
    c = make()
    c.loss_weight_dict["self"]["removal"] *= 1.3
    c.initialize_default_loss_weights()
    assert abs(c.loss_weight_dict["self"]["removal"] - 1.67 * 1.3) < 1e-9


def test_adaptive_schedule():
    c = make()
    w0 = c.loss_weight_dict["self"]["removal"]
    log = {"self": {"removal": 0.0}}
    optimization.adaptive_optimization_step_editing(c, 0, 2, log, num_ddim_steps=50, removal_loss_value_in=-1.5)   # expected < current -> x1.3
    assert abs(c.loss_weight_dict["self"]["removal"] - w0 * 1.3) < 1e-9
    log = {"self": {"removal": -5.0}}
    optimization.adaptive_optimization_step_editing(c, 0, 2, log, num_ddim_steps=50, removal_loss_value_in=-1.5)   # far below -> /2
    assert abs(c.loss_weight_dict["self"]["removal"] - w0 * 1.3 / 2.0) < 1e-9
    r = make("remove")
    w0 = r.loss_weight_dict["self"]["removal"]
    optimization.adaptive_optimization_step_remover(r, 0, 2, log, num_ddim_steps=50, removal_loss_value_in=-1.5)
    assert abs(r.loss_weight_dict["self"]["removal"] - w0 / 2.5) < 1e-9
    optimization.adaptive_optimization_step_remover(r, 25, 2, {"self": {"removal": 0.0}}, num_ddim_steps=50)       # 0.4 < frac < 0.8 -> x2
    assert abs(r.loss_weight_dict["self"]["removal"] - w0 / 2.5 * 2.0) < 1e-9


def test_scheduler_tables():
    s = diffusion.DDIMScheduler()
    s.set_timesteps(50)
    assert s.timesteps.tolist()[:3] == [980, 960, 940] and s.timesteps.tolist()[-1] == 0
    c1, c2, c3, c4 = s._coefficients(0)
    a0 = float(s.alphas_cumprod[0])
    assert abs(c3 - a0 ** 0.5) < 1e-6            # set_alpha_to_one=False: final alpha = alphas_cumprod[0]
    i1, i2, i3, i4 = s.inverse_coefficients(0)
    assert i1 == 0.0 and i2 == 1.0               # DDIMInverseScheduler: alpha before the first step is 1


def test_registration_and_mode_switch():
    model = unet_sd15.build_model("cpu", tiny=True)
    c = make()
    tc = torch.zeros(1, 512, 512, 3)
    AP.register_attention_control_diffusers(model, c, tc)
    procs = model.unet.attn_processors
    assert len(procs) == 32 and c.num_att_layers == 32
    assert all(isinstance(p, AP.EditProcessor) for p in procs.values())
    places = {k.split(".")[0]: p.place_in_unet for k, p in procs.items()}
    assert places == {"down_blocks": "down", "mid_block": "mid", "up_blocks": "up"}
    AP.set_attn_processor_for_edit(model, coords_base=(0, 1), coords_edit=(1, 2), use_cfg=False)
    assert (c.coords_base, c.coords_edit, c.use_cfg) == ((0, 1), (1, 2), False)
    model.unet.set_attn_processor(AP.VanillaAttentionProcessor())
    assert all(isinstance(p, AP.VanillaAttentionProcessor) for p in model.unet.attn_processors.values())
    with pytest.raises(ValueError):
        model.unet.set_attn_processor({"x": None})


def test_unet_topology_is_sd15():
    import torch.nn as nn

    with torch.device("meta"):
        u = unet_sd15.UNet2DConditionModel()
    assert sum(p.numel() for p in u.parameters()) == 859520964    # SD-1.x UNet parameter count
    heads = {m.to_q.in_features // m.heads for _, m in u._attention_modules()}
    assert heads == {40, 80, 160}


def test_experiment_folder_format_round_trip(tmp_path):
    """ui_utils.save_exp / read_exp layout: category folders decide the controller, two categories are skipped, the request read
    back from a folder equals the one written, and the round-robin shards cover every folder exactly once."""
    import numpy as np
    from geodiffuser_b200 import runner, synth

    root = tmp_path / "exp_root"
    written = {}
    for cat, kind in (("Translation_2D", "translate2d"), ("Rotation_3D", "rotate3d"), ("Removal", "remove"), ("Scaling", "translate2d")):
        for n in (1, 2):
            image, depth, mask, T = synth.edit_inputs(kind)
            folder = str(root / cat / str(n))
            runner.save_exp(folder, image, depth, mask, T.numpy())
            written[folder + "/"] = (image, depth, mask, T.numpy())
    folders = runner.list_exp_folders(str(root))
    assert len(folders) == 6 and all("Scaling" not in f for f, _ in folders)
    assert {t for f, t in folders if "Removal" in f} == {"geometry_remover"}
    assert {t for f, t in folders if "Removal" not in f} == {"geometry_editor"}
    f, t = folders[0]
    exp = runner.read_exp(f)
    image, depth, mask, T = written[f]
    assert np.array_equal(exp["input_image_png"], image) and exp["background_image_png"] is None
    assert np.array_equal(exp["image_shape_npy"], [512, 512])
    req = runner.request_from_exp(exp, t)
    assert np.array_equal(req["image_mask"], mask) and np.array_equal(req["depth"], depth)
    assert np.allclose(req["transform_in"].numpy(), T) and req["x0"].shape == (1, 4, 64, 64)
    shards = [runner.shard_round_robin(len(folders), r, 4) for r in range(4)]
    assert sorted(i for s in shards for i in s) == list(range(6))


def test_projection_view_and_slab_strides():
    """functional.ProjView presents the projection output (B, N, H*d) with the reference's (B*H, N, d) shape, and _Layout hands the kernels the
    element strides of one (H, N, d) slab in either layout -- pure host logic, checked against explicit indexing"""
    from geodiffuser_b200 import functional as Fn

    B, N, Nk, H, d = 3, 10, 7, 4, 8
    q = torch.arange(B * N * H * d, dtype=torch.float32).reshape(B, N, H * d)
    k = torch.arange(B * Nk * H * d, dtype=torch.float32).reshape(B, Nk, H * d)
    pv = Fn.ProjView(q, H)
    assert tuple(pv.shape) == (B * H, N, d) and pv.dtype == q.dtype and pv.device == q.device and not pv.requires_grad
    lay = Fn._Layout(q, k, H, True)
    assert (lay.N, lay.Nk, lay.d) == (N, Nk, d) and lay.q == (H * d, d) and lay.kv == (H * d, d)
    slab = lay.sl(q, 2)                                   # batch entry 2
    flat = q.reshape(-1)
    base = slab.storage_offset()
    for h, n, c in ((0, 0, 0), (1, 3, 5), (3, 9, 7)):
        assert flat[base + h * lay.q[1] + n * lay.q[0] + c] == q[2, n, h * d + c]
    # the reference's head_to_batch_dim layout through the same descriptor
    qh = q.reshape(B, N, H, d).permute(0, 2, 1, 3).reshape(B * H, N, d).contiguous()
    kh = k.reshape(B, Nk, H, d).permute(0, 2, 1, 3).reshape(B * H, Nk, d).contiguous()
    lay_h = Fn._Layout(qh, kh, H, False)
    assert lay_h.q == (d, N * d) and lay_h.kv == (d, Nk * d)
    slab_h = lay_h.sl(qh, 2)
    base_h = slab_h.storage_offset()
    flat_h = qh.reshape(-1)
    for h, n, c in ((0, 0, 0), (1, 3, 5), (3, 9, 7)):
        assert flat_h[base_h + h * lay_h.q[1] + n * lay_h.q[0] + c] == q[2, n, h * d + c]
    assert tuple(lay.new_batch(q, 2, N).shape) == (2, N, H * d) and tuple(lay_h.new_batch(qh, 2, N).shape) == (2 * H, N, d)
    st = lay.strides(out=lay.kv)
    assert list(st) == [H * d, d, H * d, d, H * d, d]


def test_algorithmic_flop_formulas_match_survey_8d():
    """forward 3 x 4 H N^2 d = 64.4 GF, backward 6 H N^2 d = 32.2 GF at the 64^2 level; correlation 2 H M N^2 = 110 GF at M = 410"""
    from geodiffuser_b200 import _lib

    assert abs(_lib.algorithmic_flops("gd_attn_fwd_sm100", (3, 8, 4096, 4096, 40)) / 1e9 - 64.4) < 0.05
    assert abs(_lib.algorithmic_flops("gd_attn_bwd_sm100", (8, 4096, 4096, 40)) / 1e9 - 32.2) < 0.05
    assert abs(_lib.algorithmic_flops("gd_removal_corr_sm100", (8, 410, 4096, 4096)) / 1e9 - 110.0) < 0.1
    assert _lib.algorithmic_flops("gd_attn_probs", (8, 410, 4096, 40)) == 0.0           # implementation work, not in the survey's count
    assert _lib.algorithmic_flops("gd_attn_fwd_generic", None) == 0.0


def test_launch_counters_are_per_thread_and_sum_up():
    """runner.EditWorkers drives one GPU from several host threads: each books its launches on its own counters (a CUDA-graph capture subtracts
    what it recorded from the capturing thread only); the module attributes read the totals; count_into books on another thread's counters."""
    import threading
    from geodiffuser_b200 import _lib

    base_l, base_f = _lib.LAUNCHES, _lib.FLOPS
    mine = _lib.counters()
    seen = {}

    def lane(n):
        c = _lib.counters()
        c.launches += n
        c.flops += 10.0 * n
        seen[n] = c
        with _lib.count_into(mine):
            _lib.counters().launches += 100 * n

    ts = [threading.Thread(target=lane, args=(n,)) for n in (1, 2)]
    l0 = mine.launches
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert seen[1] is not seen[2] and seen[1] is not mine
    assert (seen[1].launches, seen[2].launches) == (1, 2) and mine.launches - l0 == 300
    assert _lib.LAUNCHES - base_l == 303 and _lib.FLOPS - base_f == 30.0
