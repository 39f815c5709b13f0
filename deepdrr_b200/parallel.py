"""Multi-GPU view sharding for the projection path.

Views are independent units: every rank (one process per GPU) holds a replica of the volumes and
projects a contiguous chunk of the pose list.  There is no collective on the data path; the only
communication is the optional gather of finished images to rank 0 (``torch.distributed``, NCCL on
GPUs / gloo on CPU for tests).  The reference has nothing comparable (serial per-view loop on one GPU,
projector.py:679-685).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) chunk of ``n_items`` for ``rank``; chunk sizes differ by at most 1."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world_size: int) -> List[int]:
    return [shard_range(n_items, r, world_size)[1] - shard_range(n_items, r, world_size)[0] for r in range(world_size)]


class ViewShardedProjector:
    """Wraps one per-rank ``Projector``; ``project`` takes the GLOBAL pose list on every rank.

    ``projector`` is any object with ``project(*poses) -> [n, H, W]`` (the rank's own GPU projector).
    """

    def __init__(self, projector, rank: Optional[int] = None, world_size: Optional[int] = None):
        import torch.distributed as dist

        self.projector = projector
        self.rank = dist.get_rank() if rank is None else rank
        self.world_size = dist.get_world_size() if world_size is None else world_size

    def project_local(self, poses: Sequence) -> np.ndarray:
        a, b = shard_range(len(poses), self.rank, self.world_size)
        if b == a:
            return np.zeros((0, 0, 0), dtype=np.float32)
        out = self.projector.project(*poses[a:b])
        return out[None] if out.ndim == 2 else out

    def project(self, poses: Sequence, gather_to: Optional[int] = 0) -> Optional[np.ndarray]:
        """Every rank projects its chunk; if ``gather_to`` is a rank, that rank returns all images in
        pose order (others return their local chunk)."""
        import torch
        import torch.distributed as dist

        local = self.project_local(poses)
        if gather_to is None or self.world_size == 1:
            return local
        sizes = shard_sizes(len(poses), self.world_size)
        shapes = [None] * self.world_size
        dist.all_gather_object(shapes, tuple(local.shape))
        hw = next((s[1:] for s in shapes if s[0] > 0), (0, 0))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        mine = torch.from_numpy(np.ascontiguousarray(local.reshape(sizes[self.rank], *hw))).to(dev)
        bufs = [torch.empty((sizes[r],) + tuple(hw), dtype=torch.float32, device=dev) for r in range(self.world_size)]
        dist.all_gather(bufs, mine) if len(set(sizes)) == 1 else _all_gather_uneven(bufs, mine, self.rank, self.world_size)
        if self.rank == gather_to:
            return torch.cat(bufs, dim=0).cpu().numpy()
        return local


def _all_gather_uneven(bufs, mine, rank, world):
    import torch.distributed as dist

    for r in range(world):
        if r == rank:
            bufs[r].copy_(mine)
        dist.broadcast(bufs[r], src=r)
