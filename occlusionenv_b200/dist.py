"""Multi-GPU plumbing: contiguous env sharding (no communication inside the step) and the one
collective at the learner boundary: gather observations / rewards / dones of all
ranks to the PPO learner rank, rank order = env order.  One process per GPU, ``torch.distributed`` with
NCCL over NVLink on the GPU box (gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous env ids owned by ``rank``: [lo, hi).  Remainders go to the lowest ranks."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scatter_actions(actions_all: Optional[torch.Tensor], n_local: int, src: int = 0, device=None) -> torch.Tensor:
    """Learner -> env ranks: every rank receives its (n_local, 2) slice of the (N_total, 2) actions.
    Equal shard sizes are required (weak-scaling layout)."""
    world = dist.get_world_size()
    out = torch.empty(n_local, 2, dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        assert actions_all.shape == (n_local * world, 2)
        chunks = list(actions_all.to(device=device, dtype=torch.float32).contiguous().chunk(world, dim=0))
        dist.scatter(out, chunks, src=src)
    else:
        dist.scatter(out, None, src=src)
    return out


class LearnerGather:
    """Pre-allocated gather of (obs, reward, done) to the learner rank.

    ``gather()`` issues ``dist.gather`` (NCCL: a grouped send/recv over NVLink) for the three tensors;
    on the learner the results land in rank-major order, i.e. global env order.  ``all_gather()`` is the
    variant in which every rank receives everything (``ncclAllGather``)."""

    def __init__(self, n_local: int, obs_shape, device, dst: int = 0):
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.dst = dst
        self.n_local = n_local
        self.device = device
        self.obs_shape = tuple(obs_shape)
        self._bufs = None

    def _alloc(self, everyone: bool):
        if self._bufs is None and (everyone or self.rank == self.dst):
            n = self.n_local * self.world
            self._bufs = (torch.empty((n,) + self.obs_shape, dtype=torch.float32, device=self.device),
                          torch.empty(n, dtype=torch.float32, device=self.device),
                          torch.empty(n, dtype=torch.uint8, device=self.device))
        return self._bufs

    def gather(self, obs: torch.Tensor, reward: torch.Tensor, done: torch.Tensor):
        bufs = self._alloc(False)
        outs = []
        for i, t in enumerate((obs, reward, done.to(torch.uint8))):
            t = t.contiguous()
            if self.rank == self.dst:
                lst = list(bufs[i].chunk(self.world, dim=0))
                dist.gather(t, lst, dst=self.dst)
                outs.append(bufs[i])
            else:
                dist.gather(t, None, dst=self.dst)
                outs.append(None)
        return tuple(outs)

    def all_gather(self, obs: torch.Tensor, reward: torch.Tensor, done: torch.Tensor):
        bufs = self._alloc(True)
        for i, t in enumerate((obs, reward, done.to(torch.uint8))):
            dist.all_gather_into_tensor(bufs[i], t.contiguous())
        return bufs
