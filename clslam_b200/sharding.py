"""Host-side logic of the data-parallel adaptation step.

The (current + replay) mini-batch is sharded by sample: rank r owns global samples
``[r*B/G, (r+1)*B/G)``; the online sample (global index 0) therefore lives on rank 0.  The loss is
``sum_b w_b * L_b`` with GLOBAL weights ``w_b`` (depth_pose_prediction.py:1032, :1073, :1090,
:1108), so each rank evaluates its local terms with its slice of the global weight vector and a
single SUM all-reduce of the flat gradient buffer (and of the loss scalars) reproduces the
single-process result.  The reference's smoothness term reads only batch element 0
(SURVEY.md A.7), so rank 0 alone evaluates it, for the whole global batch.

Pure PyTorch / no CUDA: exercised with world_size-2 gloo tests on the CPU.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np
import torch


def shard_range(batch_global: int, rank: int, world: int) -> Tuple[int, int]:
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f'bad rank/world {rank}/{world}')
    if batch_global % world:
        raise ValueError(f'global batch {batch_global} is not divisible by world size {world}')
    per = batch_global // world
    return rank * per, (rank + 1) * per


def shard_batch(batch: Dict[Any, torch.Tensor], rank: int, world: int) -> Dict[Any, torch.Tensor]:
    """Slice every value of a reference-schema batch dict along dim 0."""
    some = batch[('rgb', 0, 0)]
    lo, hi = shard_range(some.shape[0], rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def global_sample_weights(batch_size: int, online_loss_weight: Optional[float] = None, online_index: int = 0,
                          ) -> torch.Tensor:
    """Per-sample loss weights of ``adapt`` (depth_pose_prediction.py:297-305) /
    ``_compute_loss`` default (:1031-1032), for the GLOBAL batch, on the CPU."""
    if online_loss_weight is None:
        return torch.ones(batch_size) / batch_size
    if batch_size == 1:
        return torch.ones(1)
    w = torch.empty(batch_size)
    w[np.arange(batch_size) != online_index] = (1 - online_loss_weight) / (batch_size - 1)
    w[online_index] = online_loss_weight
    return w


def shard_noise(noise_global: Sequence[torch.Tensor], rank: int, world: int) -> Sequence[torch.Tensor]:
    """The tie-break noise is drawn as one global (B,2,H,W) tensor per scale in the reference's
    order (depth_pose_prediction.py:1055-1056) and sliced per rank (parity mode)."""
    lo, hi = shard_range(noise_global[0].shape[0], rank, world)
    return [n[lo:hi] for n in noise_global]


def allreduce_sum_(tensors: Sequence[torch.Tensor], group=None) -> None:
    """The one exchange step of the path: SUM all-reduce (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
