"""Pose-sharded scans across the GPUs of one box: one process per GPU under
``torch.distributed``, the map and its distance field replicated on every GPU (each rank runs the
ingest kernels itself -- it is a sub-millisecond job), the pose batch split into contiguous index
ranges, and ONE collective: an all-gather (or gather-to-root) of the fp32 ranges over NCCL/NVLink.
There is no other exchange step -- every ray is independent.

The march itself is injected as a callable so that the partition / padding / gather logic can be
exercised on CPU with the ``gloo`` backend (tests/test_sharded_gloo.py); the product binding is
:func:`gpu_march_fn`.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous pose range [lo, hi) of `rank`: ceil(n / world) poses per rank, the last ranks
    possibly short or empty (pose-major output stays contiguous per rank)."""
    per = -(-n // world) if n > 0 else 0
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def gpu_march_fn(marcher, fov: float, num_rays: int) -> Callable:
    """The product march: ``PyRayMarchingGPU.calc_range_fan`` on device tensors."""
    def run(poses: torch.Tensor, out: torch.Tensor):
        if poses.shape[0]:
            marcher.calc_range_fan(poses, out, fov, num_rays)
    return run


class ShardedScanner:
    """``scan(poses)`` over the whole process group.

    poses:   (N, 3) float32, identical on every rank (only the rank's own rows are read).
    returns: (N * num_rays,) float32 ranges, pose-major -- on every rank for ``gather="all"``, on
             ``root`` only (None elsewhere) for ``gather="root"``, and just the local shard
             (``hi - lo`` poses) for ``gather="none"``.
    """

    def __init__(self, march_fn: Callable, num_rays: int, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None, chunks: int = 1):
        self.march_fn = march_fn
        self.num_rays = int(num_rays)
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.chunks = max(1, int(chunks))

    def scan(self, poses: torch.Tensor, gather: str = "all", root: int = 0):
        if gather not in ("all", "root", "none"):
            raise ValueError("gather must be 'all', 'root' or 'none'")
        n = poses.shape[0]
        lo, hi = shard_bounds(n, self.world, self.rank)
        per = -(-n // self.world) if n > 0 else 0
        R = self.num_rays
        # equal, padded shards: all_gather needs the same count from every rank
        local = torch.zeros(per * R, dtype=torch.float32, device=self.device)
        mine = poses[lo:hi].to(self.device).contiguous()
        self.march_fn(mine, local[:(hi - lo) * R])
        if gather == "none" or self.world == 1:
            out = local[:(hi - lo) * R]
            return out if (gather != "root" or self.rank == root) else None
        if gather == "all":
            full = torch.empty(self.world * per * R, dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(full, local, group=self.group)
            return full[:n * R]
        parts = [torch.empty_like(local) for _ in range(self.world)] if self.rank == root else None
        dist.gather(local, parts, dst=root, group=self.group)
        if self.rank != root:
            return None
        return torch.cat(parts)[:n * R]
