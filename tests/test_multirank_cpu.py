"""N > 1 path on CPU: world_size-2 gloo processes shard a request list round-robin with no data-path collective and reduce the
throughput report (sum of counts, max of seconds) exactly as bench.py does over NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from geodiffuser_b200 import runner

    requests = [{"id": i} for i in range(7)]
    served, secs = runner.run_requests(None, requests, rank, world, edit_fn=lambda r: torch.tensor([r["id"] * 10.0]))
    n, t, thr = runner.reduce_throughput(len(served), 1.0 + rank)       # rank 1 is "slower": max must win
    gathered = [None] * world
    dist.all_gather_object(gathered, sorted(served))
    q.put((rank, sorted(served), n, t, thr, gathered))
    dist.destroy_process_group()


def test_round_robin_sharding_and_throughput_reduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    for _, _, n, t, thr, gathered in res:
        assert n == 7.0 and t == 2.0 and abs(thr - 3.5) < 1e-9
        assert sorted(sum(gathered, [])) == list(range(7))          # every request served exactly once, no overlap


def test_requests_of_a_rank_start_longest_first():
    from geodiffuser_b200 import runner

    assert runner.longest_first([1.0, 2.3, 1.0, 2.3, 1.0]) == [1, 3, 0, 2, 4]
    assert runner.longest_first([]) == []
    # every request of the shard is started exactly once, whatever the costs
    shard = runner.shard_round_robin(64, 3, 8)
    order = runner.longest_first([2.3 if i % 4 == 3 else 1.0 for i in shard])
    assert sorted(order) == list(range(len(shard)))
