"""Data-parallel plumbing: one process per GPU, torch.distributed for rendezvous only.

Inference shards the batch across ranks with no collective.  Head training all-reduces one flat
gradient buffer per step inside ``hp_head_train_step`` with NCCL; the ncclUniqueId is created on
rank 0 by libhpose and broadcast through the already-initialised ``torch.distributed`` group.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DataParallel:
    def __init__(self, ctx=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("initialise torch.distributed first (torchrun sets RANK/WORLD_SIZE)")
        self.rank, self.world_size = dist.get_rank(), dist.get_world_size()
        self.ctx = ctx
        self._comm_ready = False

    def init_gradient_comm(self, p2p=True, cap_floats=65536):
        """Create the NCCL communicator used by hp_head_train_step and, with ``p2p`` (default), the peer-memory exchange used
        by hp_head_train_run (heads of up to ``cap_floats - 3`` parameters)."""
        import torch
        import torch.distributed as dist
        from .device import default_context
        ctx = self.ctx or default_context()
        buf = np.zeros(128, np.uint8)
        if self.rank == 0:
            _lib.check(_lib.lib().hp_comm_unique_id(buf.ctypes.data))
        t = torch.from_numpy(buf)
        if dist.get_backend() == "nccl":
            t = t.to(ctx.torch_device)
            dist.broadcast(t, 0)
            buf = t.cpu().numpy()
        else:
            dist.broadcast(t, 0)
            buf = t.numpy()
        buf = np.ascontiguousarray(buf)
        _lib.check(_lib.lib().hp_comm_init(ctx.handle, buf.ctypes.data, self.rank, self.world_size))
        self._comm_ready = True
        self.p2p = False
        if p2p and 2 <= self.world_size <= 8 and dist.get_backend() == "nccl":
            self.p2p = self._init_peer_exchange(ctx, cap_floats)
        return self

    def _init_peer_exchange(self, ctx, cap_floats):
        """Map every rank's gradient inbox into every process (CUDA IPC over NVLink): the fused all-reduce + optimizer kernel
        of hp_head_train_run then replaces ncclAllReduce + optimizer.  Every rank must take the same decision, so a rank that
        cannot map a peer reports it and all ranks fall back to NCCL together."""
        import torch
        import torch.distributed as dist
        L = _lib.lib()
        mine = np.zeros(64, np.uint8)
        ok = L.hp_p2p_alloc(ctx.handle, int(cap_floats), self.world_size, mine.ctypes.data) == 0
        t = torch.from_numpy(mine).to(ctx.torch_device)
        allh = [torch.empty_like(t) for _ in range(self.world_size)]
        dist.all_gather(allh, t)
        handles = np.ascontiguousarray(np.concatenate([a.cpu().numpy() for a in allh]))
        ok = ok and L.hp_p2p_open(ctx.handle, handles.ctypes.data, self.rank, self.world_size) == 0
        flag = torch.tensor([1 if ok else 0], device=ctx.torch_device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            L.hp_p2p_close(ctx.handle)
            return False
        return True
