"""Request-level data parallelism over independent edits: one process per GPU, edit r -> rank r mod G, no collective
on the data path.  The natural seam in the reference is the folder loop of large_scale_editor.py:392-399.  torch.distributed is used
only for the end-of-run throughput report (max time / total count over ranks)."""
import time

import torch
import torch.distributed as dist


def shard_round_robin(n_requests, rank, world_size):
    """indices of the requests this rank serves"""
    return list(range(rank, n_requests, world_size))


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
