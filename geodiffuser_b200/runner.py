"""Request-level data parallelism over independent edits: one process per GPU, edit r -> rank r mod G, no collective
on the data path.  The natural seam in the reference is the folder loop of large_scale_editor.py:392-399.  torch.distributed is used
only for the end-of-run throughput report (max time / total count over ranks)."""
import time

import torch
import torch.distributed as dist


def shard_round_robin(n_requests, rank, world_size):
    """indices of the requests this rank serves"""
    return list(range(rank, n_requests, world_size))


def longest_first(costs):
    """order in which a rank starts its requests: descending cost, ties in arrival order (its lanes take requests as they become free, so the
    expensive ones -- a 768^2 edit is ~2.3 x a 512^2 one -- must not be the last to start)"""
    return sorted(range(len(costs)), key=lambda i: (-costs[i], i))


def reduce_throughput(n_done, seconds, device=None):
    """(total edits over all ranks, max seconds over ranks, edits/sec).  Works with nccl (GPU tensors) and gloo (CPU tensors)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
        cnt = torch.tensor([float(n_done)], device=dev, dtype=torch.float64)
        sec = torch.tensor([float(seconds)], device=dev, dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        n_done, seconds = float(cnt), float(sec)
    return n_done, seconds, n_done / max(seconds, 1e-12)


def run_requests(model, requests, rank=0, world_size=1, edit_fn=None):
    """Serves this rank's shard of `requests` (dicts as produced by editor.synthetic_request).  Returns ({index: latents}, seconds)."""
    from . import editor

    if edit_fn is None:
        edit_fn = lambda req: editor.perform_geometric_edit(model, req["depth"], req["image_mask"], req["transform_in"], req["text_embeddings"],
                                                            req["uncond_embeddings"], req["x0"], req["edit_type"])[0]
    out = {}
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in shard_round_robin(len(requests), rank, world_size):
        out[i] = edit_fn(requests[i])
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return out, time.perf_counter() - t0


class EditWorkers:
    """`lanes` independent edits in flight on ONE GPU: one persistent host thread, CUDA stream and model replica (shared weights, own
    processors / controller / graphs / caches: unet_sd15.replicate_model) per lane, requests dealt round-robin to the lanes.  Still
    request-level parallelism with no exchange between edits: every edit runs exactly the kernels, in the order, it runs alone, so its result
    does not depend on the lane count.  Why: one edit is a chain of ~90 000 dependent launches of mostly small kernels (8^2 .. 64^2 tokens,
    batch 2-3) that individually cannot fill 148 SMs; a second independent chain fills the gaps.
    The threads are persistent because cuDNN's autotune cache (torch.backends.cudnn.benchmark) is per host thread: a fresh thread would
    re-tune the body's convolutions, which is illegal inside the stream capture of an edit's optimisation pass."""

    def __init__(self, model, lanes=2):
        import queue
        import threading
        from .unet_sd15 import replicate_model

        dev = model.device
        self.device = dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
        self.models = [model] + [replicate_model(model) for _ in range(max(1, lanes) - 1)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.models]
        self._jobs = queue.Queue()          # ONE queue for all lanes: a lane takes the next request when it is free (requests differ in cost:
                                            # a 768^2 edit is ~2.3 x a 512^2 one, and a static deal leaves lanes idle at the end of a batch)
        self._own = [queue.Queue() for _ in self.models]      # requests pinned to one lane (warm-up: map_every_lane)
        self._done = queue.Queue()
        self._threads = [threading.Thread(target=self._lane, args=(w,), daemon=True) for w in range(len(self.models))]
        for t in self._threads:
            t.start()

    def _lane(self, w):
        from . import graphs

        torch.cuda.set_device(self.device)
        # cuBLAS / cuDNN handles are per host thread and created on first use (cublasCreate allocates and synchronises): create this thread's
        # now, under the capture lock, so that it never happens while another lane records a graph (that capture would be invalidated)
        with graphs._CAPTURE_LOCK, torch.cuda.stream(self.streams[w]), torch.no_grad():
            a = torch.ones(8, 8, device=self.device, dtype=torch.bfloat16)
            (a @ a).float() @ torch.ones(8, 8, device=self.device)
            torch.nn.functional.conv2d(a.reshape(1, 1, 8, 8), torch.ones(1, 1, 3, 3, device=self.device, dtype=torch.bfloat16), padding=1)
            self.streams[w].synchronize()
        import queue

        while True:
            try:
                job = self._own[w].get_nowait()
            except queue.Empty:
                try:
                    job = self._jobs.get(timeout=0.002)
                except queue.Empty:
                    continue
            if job is None:
                return
            fn, item, idx, after = job
            try:
                self.streams[w].wait_stream(after)
                with torch.cuda.stream(self.streams[w]):
                    self._done.put((idx, fn(self.models[w], item), None))
            except BaseException as e:   # noqa: BLE001 -- re-raised in the caller's thread
                import os
                import traceback

                if os.environ.get("GD_DEBUG_LANES"):
                    print(f"[lane {w}] failed:", flush=True)
                    traceback.print_exc()
                self._done.put((idx, None, e))

    def map(self, fn, items):
        """results[i] = fn(a lane's model, items[i]), in the order of `items`; requests are started in that order, each by the first lane that is
        free.  Returns once every request has been queued on its lane's stream, with the calling stream waiting for all lanes."""
        cur = torch.cuda.current_stream(self.device)
        for i, item in enumerate(items):
            self._jobs.put((fn, item, i, cur))
        results, error = [None] * len(items), None
        for _ in items:
            idx, res, err = self._done.get()
            results[idx], error = res, (err if err is not None and error is None else error)
        for s in self.streams:
            cur.wait_stream(s)
        if error is not None:
            raise error
        return results

    def map_every_lane(self, fn, items):
        """fn(lane's model, item) for every item on EVERY lane (warm-up: graphs, per-thread plan caches); -> results of lane 0"""
        n = len(self.models)
        cur = torch.cuda.current_stream(self.device)
        for w in range(n):
            for i, item in enumerate(items):
                self._own[w].put((fn, item, (w, i), cur))
        results, error = [None] * len(items), None
        for _ in range(n * len(items)):
            (w, i), res, err = self._done.get()
            if w == 0:
                results[i] = res
            error = err if err is not None and error is None else error
        for s in self.streams:
            cur.wait_stream(s)
        if error is not None:
            raise error
        return results

    def close(self):
        for q in self._own:
            q.put(None)


# ---- experiment-folder format of the reference's batch driver ---------------------------------------------------------
# ui_utils.save_exp / read_exp (:52-159) and large_scale_editor.py:133-178, 349-399: one edit = one folder holding input_image.png,
# input_mask.png, depth.npy, transform.npy, image_shape.npy; results are written next to them (loss.pkl; the reference also decodes
# result_ls.png through the VAE, which is outside this build, so the edited latents are stored as latents_ls.npy instead).
EXP_FILES = ("input_image.png", "depth.npy", "input_mask.png", "transform.npy", "image_shape.npy", "background_image.png")
SKIPPED_CATEGORIES = ("Rotation_2D", "Scaling")       # large_scale_editor.py:372


def save_exp(folder, input_img, input_depth, input_mask, transform_in, h=512, w=512):
    """ui_utils.save_exp (:52-112), the files the batch driver reads"""
    import os
    import cv2
    import numpy as np

    os.makedirs(folder, exist_ok=True)
    cv2.imwrite(os.path.join(folder, "input_image.png"), np.asarray(input_img, np.uint8)[..., ::-1])
    m = np.asarray(input_mask, np.float64)
    m8 = (m * 255.0 if m.max() <= 1.0 else m).astype(np.uint8)
    cv2.imwrite(os.path.join(folder, "input_mask.png"), np.repeat(m8[..., None], 3, -1) if m8.ndim == 2 else m8)
    np.save(os.path.join(folder, "depth.npy"), np.asarray(input_depth))
    np.save(os.path.join(folder, "transform.npy"), np.asarray(transform_in))
    np.save(os.path.join(folder, "image_shape.npy"), np.array([int(h), int(w)]))


def read_exp(d_path):
    """ui_utils.read_exp (:118-159): {'<name>_png' | '<name>_npy': array or None, 'path_name': folder}"""
    import os
    import cv2
    import numpy as np

    folder = d_path if d_path.endswith("/") else d_path + "/"
    out = {}
    for f in EXP_FILES:
        key, ext = f.split(".")
        path = folder + f
        val = None
        if os.path.exists(path):
            if ext == "png":
                val = cv2.imread(path, cv2.IMREAD_UNCHANGED)
                if val.ndim == 3:
                    val = val[..., :3][..., ::-1].copy()
            else:
                val = np.load(path)
        out[f"{key}_{ext}"] = val
    if out["image_shape_npy"] is None:
        out["image_shape_npy"] = np.array([512, 512])
    out["path_name"] = folder
    return out


def exp_type_of_category(category):
    """large_scale_editor.py:368-380: the folder's category decides the controller; two categories are skipped"""
    if category in SKIPPED_CATEGORIES:
        return None
    return "geometry_remover" if category == "Removal" else "geometry_editor"


def list_exp_folders(root):
    """[(folder, exp_type)] under an experiment root `root/<Category>/<n>/` (check_if_exp_root layout), sorted like the reference's glob"""
    import glob
    import os

    out = []
    for cat_dir in sorted(glob.glob(os.path.join(root, "*/"))):
        exp_type = exp_type_of_category(os.path.basename(os.path.normpath(cat_dir)))
        if exp_type is None:
            continue
        for f in sorted(glob.glob(os.path.join(cat_dir, "*/"))):
            if os.path.exists(os.path.join(f, "depth.npy")):
                out.append((f, exp_type))
    return out


def request_from_exp(exp_dict, exp_type, seed=1234):
    """perform_exp (:199-235): image_mask = input_mask[..., 0] / 255, depth, transform from the folder; the text / image encodings are
    synthetic (CLIP and the VAE are outside this build)."""
    import numpy as np
    from . import editor

    mask = exp_dict["input_mask_png"]
    mask = (mask[..., 0] if mask.ndim == 3 else mask) / 255.0
    size = int(mask.shape[0])
    text, uncond, x0 = editor.synthetic_embeddings(seed, "cpu", size)
    return dict(depth=np.asarray(exp_dict["depth_npy"], np.float64), image_mask=mask.astype(np.float64),
                transform_in=torch.tensor(np.asarray(exp_dict["transform_npy"])).float(), text_embeddings=text, uncond_embeddings=uncond, x0=x0,
                edit_type=exp_type)


def run_exp_on_folder_single(model, exp_folder, exp_type, **kw):
    """large_scale_editor.py:349-365: read the folder, edit, write the results into it.  Returns the edited latents (host)."""
    import pickle
    import numpy as np
    from . import editor

    exp = read_exp(exp_folder)
    req = request_from_exp(exp, exp_type)
    staged, _ = editor.stage_inputs(req["depth"], req["image_mask"], req["text_embeddings"], req["uncond_embeddings"], req["x0"], model.device)
    latents, log = editor.run_edit(model, staged, req["transform_in"], exp_type, return_log=True, **kw)
    out = latents.float().cpu().numpy()
    np.save(exp["path_name"] + "latents_ls.npy", out)
    with open(exp["path_name"] + "loss.pkl", "wb") as f:
        pickle.dump(log, f)
    return out


def run_exp_root(model, root, rank=0, world_size=1, **kw):
    """the reference's folder loop (:392-399), sharded round-robin over ranks.  Returns the folders this rank served."""
    folders = list_exp_folders(root)
    done = []
    for i in shard_round_robin(len(folders), rank, world_size):
        run_exp_on_folder_single(model, folders[i][0], folders[i][1], **kw)
        done.append(folders[i][0])
    return done
