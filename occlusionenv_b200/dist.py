"""Multi-GPU plumbing: contiguous env sharding (no communication inside the step) and the one exchange at the
learner boundary: observations / rewards / dones of all ranks to the PPO learner rank, rank
order = env order -- the ``torch.stack`` of ``/root/reference/SubProcVecEnv.py:219`` across GPUs.  One process per
GPU, ``torch.distributed`` with NCCL over NVLink on the GPU box (gloo in the CPU tests).

Two transports for the observations, both without a staging copy:

* ``transport="nccl"``: ONE grouped NCCL operation (``batch_isend_irecv`` = one ncclGroup) moves obs + reward + done
  of every rank straight into the rank's slice of the learner's contiguous (N_total, ...) buffers.  The rasteriser
  renders into ``obs_send_buffer()`` -- on the learner rank that IS its slice of the gathered tensor.
* ``transport="p2p"``: the learner's gather buffer is mapped into every env rank's address space (CUDA IPC over
  NVLink / NVSwitch peer memory) and ``obs_send_buffer()`` returns the rank's slice of THAT: the rasteriser's epilogue
  stores every observation row directly into the learner's HBM while the tile is being finished -- compute and
  transfer are one kernel, there is no collective for the observations at all.  ``gather()`` then only moves
  reward + done (5 B/env), which, being stream-ordered after the step, also tells the learner that the rows landed.

The transfer is bound by the learner GPU's NVLink ingest (measured peer copy 770 GB/s per direction,
B200_PROFILING.md): 262 144 B/env at 128^2.  ``planes=2`` (OcclConfig.obs_planes) halves it: the reference's R = G = B
(flat shading of white vertices) travel once as a grey plane next to the depth plane; ``expand_compact_obs`` restores
(N, 4, S, S) on the learner.
"""
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


def expand_compact_obs(obs2: torch.Tensor) -> torch.Tensor:
    """(N, 2, S, S) grey + depth -> the reference's (N, 4, S, S) RGB + depth (``environment.py:376-378``)."""
    return obs2[:, (0, 0, 0, 1)]


class PeerView:
    """A (slice of a) float32 tensor that lives in ANOTHER GPU's memory, mapped into this process with CUDA IPC on this
    rank's own device.  It only carries the address: ``engine.step(..., obs=view)`` lets the rasteriser store its
    observation rows there (``OcclOutputs.obs``), nothing else in this process reads or writes it."""

    def __init__(self, ptr: int, shape, itemsize: int = 4):
        self._ptr, self.shape, self._itemsize = int(ptr), tuple(shape), itemsize

    def data_ptr(self) -> int:
        return self._ptr

    def __getitem__(self, sl: slice) -> "PeerView":
        lo, hi, step = sl.indices(self.shape[0])
        assert step == 1
        row = self._itemsize
        for d in self.shape[1:]:
            row *= d
        return PeerView(self._ptr + lo * row, (hi - lo,) + self.shape[1:], self._itemsize)


_IPC_OPENED = {}  # exported block (64-byte handle) -> base address in this process


def _share_cuda_tensor(t: Optional[torch.Tensor], src: int):
    """Map ``t`` (a contiguous float32 tensor allocated on rank ``src``) into every other rank's address space through
    CUDA IPC, opened on the importing rank's OWN device (peer access over NVLink / NVSwitch).  Returns ``t`` on ``src``
    and a ``PeerView`` elsewhere."""
    import ctypes

    from . import _lib as L
    meta = None
    if dist.get_rank() == src:
        assert t.is_contiguous() and t.dtype == torch.float32
        hbuf, off = ctypes.create_string_buffer(64), ctypes.c_size_t()
        L.check(L.load().occl_ipc_export(ctypes.c_void_p(t.data_ptr()), hbuf, ctypes.byref(off)), "occl_ipc_export")
        meta = (hbuf.raw, int(off.value), tuple(t.shape))
    box = [meta]
    dist.broadcast_object_list(box, src=src)
    if dist.get_rank() == src:
        return t
    handle, offset, shape = box[0]
    base = _IPC_OPENED.get(handle)
    if base is None:
        out = ctypes.c_void_p()
        L.check(L.load().occl_ipc_open(handle, ctypes.byref(out)), "occl_ipc_open")
        base = _IPC_OPENED[handle] = int(out.value)
    return PeerView(base + offset, shape)


class LearnerGather:
    """Pre-allocated exchange of (obs, reward, done) to the learner rank; see the module docstring.

    ``gather()`` returns, on the learner, ``(obs (N_total, planes, S, S), reward (N_total,), done (N_total,) u8)`` in
    global env order; ``(None, None, None)`` elsewhere.  ``n_buffers`` > 1 rotates the destination buffers so that
    the learner can read step t while step t+1 is being written (``next_buffer()``)."""

    def __init__(self, n_local: int, obs_shape, device, dst: int = 0, transport: str = "nccl", n_buffers: int = 1):
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.dst = dst
        self.n_local = n_local
        self.device = device
        self.obs_shape = tuple(obs_shape)
        if transport not in ("nccl", "p2p"):
            raise ValueError("transport must be 'nccl' or 'p2p'")
        self.transport = transport
        self.n_buffers = int(n_buffers)
        self._cur = 0
        self._g_obs, self._g_small, self._send_obs, self._all = None, None, None, None
        lo = self.rank * n_local
        self._slice = slice(lo, lo + n_local)

    # -- buffers -----------------------------------------------------------------------------------
    def _alloc(self):
        if self._g_small is not None:
            return
        n = self.n_local * self.world
        learner = self.rank == self.dst
        if learner:
            self._g_obs = [torch.empty((n,) + self.obs_shape, dtype=torch.float32, device=self.device)
                           for _ in range(self.n_buffers)]
            self._g_small = (torch.empty(n, dtype=torch.float32, device=self.device),
                             torch.empty(n, dtype=torch.uint8, device=self.device))
        else:
            self._g_small = (None, None)
        if self.transport == "p2p":
            # every rank sees the learner's buffers; env ranks write their slice from inside the rasteriser
            self._g_obs = [_share_cuda_tensor(self._g_obs[b] if learner else None, self.dst) for b in range(self.n_buffers)]
        elif not learner:
            self._send_obs = [torch.empty((self.n_local,) + self.obs_shape, dtype=torch.float32, device=self.device)
                              for _ in range(self.n_buffers)]

    def obs_send_buffer(self) -> torch.Tensor:
        """Where the rasteriser should write this rank's observations (``engine.step(..., obs=...)``) so that
        ``gather`` needs no copy: the rank's slice of the learner's buffer (learner rank, or any rank with the p2p
        transport), else a local send buffer."""
        self._alloc()
        if self.transport == "p2p" or self.rank == self.dst:
            return self._g_obs[self._cur][self._slice]
        return self._send_obs[self._cur]

    def next_buffer(self):
        self._cur = (self._cur + 1) % self.n_buffers

    # -- the exchange --------------------------------------------------------------------------------
    def gather(self, obs: torch.Tensor, reward: torch.Tensor, done: torch.Tensor):
        self._alloc()
        learner = self.rank == self.dst
        done = done if done.dtype == torch.uint8 else done.to(torch.uint8)
        in_place = obs.data_ptr() == self.obs_send_buffer().data_ptr()
        g_obs = self._g_obs[self._cur] if (learner or self.transport == "p2p") else None
        move_obs = not (self.transport == "p2p" and in_place)
        if self.transport == "p2p" and not in_place:
            if not learner:
                raise ValueError("p2p transport: render into obs_send_buffer() (engine.step(..., obs=buffer))")
            g_obs[self._slice].copy_(obs)
            move_obs = False
        if dist.get_backend() != "nccl":
            return self._gather_collective(obs, reward, done, move_obs)
        ops = []
        if learner:
            g_rew, g_done = self._g_small
            for r in range(self.world):
                sl = slice(r * self.n_local, (r + 1) * self.n_local)
                if r == self.rank:
                    if move_obs and not in_place:
                        g_obs[sl].copy_(obs)
                    g_rew[sl].copy_(reward)
                    g_done[sl].copy_(done)
                    continue
                if move_obs:
                    ops.append(dist.P2POp(dist.irecv, g_obs[sl], r))
                ops.append(dist.P2POp(dist.irecv, g_rew[sl], r))
                ops.append(dist.P2POp(dist.irecv, g_done[sl], r))
        else:
            if move_obs:
                ops.append(dist.P2POp(dist.isend, obs.contiguous(), self.dst))
            ops.append(dist.P2POp(dist.isend, reward.contiguous(), self.dst))
            ops.append(dist.P2POp(dist.isend, done.contiguous(), self.dst))
        for w in dist.batch_isend_irecv(ops):  # one ncclGroup: a single grouped transfer for all three tensors
            w.wait()                           # (stream-ordered: no host block with NCCL)
        if learner:
            return self._g_obs[self._cur], self._g_small[0], self._g_small[1]
        return None, None, None

    def _gather_collective(self, obs, reward, done, move_obs=True):
        """gloo (CPU tests): plain ``dist.gather`` per tensor."""
        learner = self.rank == self.dst
        outs = []
        bufs = (self._g_obs[self._cur] if learner else None,) + tuple(self._g_small)
        for i, t in enumerate((obs, reward, done)):
            t = t.contiguous()
            if learner:
                dist.gather(t, list(bufs[i].chunk(self.world, dim=0)), dst=self.dst)
                outs.append(bufs[i])
            else:
                dist.gather(t, None, dst=self.dst)
                outs.append(None)
        return tuple(outs)

    def all_gather(self, obs: torch.Tensor, reward: torch.Tensor, done: torch.Tensor):
        """Variant in which every rank receives everything (``ncclAllGather``)."""
        if self._all is None:
            n = self.n_local * self.world
            self._all = (torch.empty((n,) + self.obs_shape, dtype=torch.float32, device=self.device),
                         torch.empty(n, dtype=torch.float32, device=self.device),
                         torch.empty(n, dtype=torch.uint8, device=self.device))
        for i, t in enumerate((obs, reward, done.to(torch.uint8))):
            dist.all_gather_into_tensor(self._all[i], t.contiguous())
        return self._all


class FeatureGather:
    """Row N-1: the env ranks run the policy's frozen encoder (occlusionenv_b200/features.py) on
    the observations they render and the learner receives (N_total, n_features) pooled features + reward + done --
    1 029 B/env instead of 262 149 B/env at 128^2 -- in one grouped NCCL transfer."""

    def __init__(self, n_local: int, n_features: int, device, dst: int = 0):
        self.world, self.rank, self.dst, self.n_local = dist.get_world_size(), dist.get_rank(), dst, n_local
        n = n_local * self.world
        self._bufs = None
        if self.rank == dst:
            self._bufs = (torch.empty(n, n_features, dtype=torch.float32, device=device),
                          torch.empty(n, dtype=torch.float32, device=device),
                          torch.empty(n, dtype=torch.uint8, device=device))

    def gather(self, features: torch.Tensor, reward: torch.Tensor, done: torch.Tensor):
        done = done if done.dtype == torch.uint8 else done.to(torch.uint8)
        tensors = (features.contiguous(), reward.contiguous(), done.contiguous())
        if dist.get_backend() != "nccl":
            for i, t in enumerate(tensors):
                dist.gather(t, list(self._bufs[i].chunk(self.world, dim=0)) if self.rank == self.dst else None, dst=self.dst)
            return self._bufs if self.rank == self.dst else (None, None, None)
        ops = []
        if self.rank == self.dst:
            for r in range(self.world):
                sl = slice(r * self.n_local, (r + 1) * self.n_local)
                for i, t in enumerate(tensors):
                    if r == self.rank:
                        self._bufs[i][sl].copy_(t)
                    else:
                        ops.append(dist.P2POp(dist.irecv, self._bufs[i][sl], r))
        else:
            ops = [dist.P2POp(dist.isend, t, self.dst) for t in tensors]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return self._bufs if self.rank == self.dst else (None, None, None)
