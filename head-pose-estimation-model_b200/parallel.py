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

    def init_gradient_comm(self):
        """Create the NCCL communicator used by hp_head_train_step (CUDA only)."""
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
        return self
