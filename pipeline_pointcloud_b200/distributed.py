"""Multi-GPU plumbing for the rasterization path: one process per GPU,
``torch.distributed`` (NCCL over NVLink 5 / NVSwitch on the B200 box, gloo in CPU tests of the host logic).

(i) camera data-parallel: Gaussians replicated, each rank renders its own cameras, dense-gradient all-reduce.
(ii) Gaussian-sharded projection + tile-stripe ownership with an all-to-all of projected splats.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def cameras_of_rank(n_cameras: int, rank: int, world_size: int) -> List[int]:
    """Camera data-parallel assignment: rank r renders cameras c with c % world_size == r."""
    return [c for c in range(n_cameras) if c % world_size == rank]


def allreduce_gradients(params: Iterable[Tensor], group=None, average: bool = False) -> None:
    """Sums (or averages) ``p.grad`` over ranks, one collective per parameter group so that the
    SH-coefficient bucket (81 % of the bytes) is in flight while the small ones are issued."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    ws = dist.get_world_size(group)
    works = []
    for p in sorted((p for p in params if p.grad is not None), key=lambda t: -t.grad.numel()):
        works.append((dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group, async_op=True), p))
    for w, p in works:
        w.wait()
        if average:
            p.grad.div_(ws)


def rasterization_distributed(*args, **kwargs):
    raise NotImplementedError(
        "distributed=True (Gaussian-sharded projection + all-to-all) is not implemented yet; "
        "camera data-parallel training uses rasterization() per rank + allreduce_gradients()")
