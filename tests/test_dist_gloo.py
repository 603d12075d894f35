"""world_size-2 gloo test of the multi-GPU host logic (sharding, action scatter, learner gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from occlusionenv_b200.dist import LearnerGather, scatter_actions, shard_range


def test_shard_range_covers_everything():
    for n, w in [(65536, 8), (10, 3), (7, 8)]:
        ids = []
        for r in range(w):
            lo, hi = shard_range(n, r, w)
            ids += list(range(lo, hi))
        assert ids == list(range(n))
    assert shard_range(65536, 3, 8) == (3 * 8192, 4 * 8192)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local, S = 3, 4
    lo, hi = shard_range(n_local * world, rank, world)
    all_actions = torch.arange(n_local * world * 2, dtype=torch.float32).reshape(-1, 2) if rank == 0 else None
    a = scatter_actions(all_actions, n_local, src=0, device="cpu")
    ok = torch.equal(a, torch.arange(n_local * world * 2, dtype=torch.float32).reshape(-1, 2)[lo:hi])
    ids = torch.arange(lo, hi, dtype=torch.float32)
    obs = ids[:, None, None, None].expand(n_local, 4, S, S).contiguous()
    g = LearnerGather(n_local, (4, S, S), "cpu", dst=0)
    o, r, d = g.gather(obs, ids * 2, (ids % 2).to(torch.uint8))
    if rank == 0:
        want = torch.arange(n_local * world, dtype=torch.float32)
        ok &= torch.equal(o[:, 0, 0, 0], want) and torch.equal(r, want * 2) and torch.equal(d, (want % 2).to(torch.uint8))
    else:
        ok &= o is None
    o2, r2, d2 = LearnerGather(n_local, (4, S, S), "cpu").all_gather(obs, ids * 2, (ids % 2).to(torch.uint8))
    ok &= torch.equal(r2, torch.arange(n_local * world, dtype=torch.float32) * 2)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_and_scatter_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
