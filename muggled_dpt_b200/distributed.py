"""
Multi-GPU plumbing for the depth path: frames of a batch never interact
(muggled_dpt/dpt_model.py:61-83 has no cross-batch op), so the batch is sharded over ranks - one process per GPU, the
weights replicated - and the only collective is ONE all-gather of the [B/G, H, W] depth maps per step.
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is plumbing, not compute.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of the frames rank `rank` owns; the first (global_batch % world) ranks take one extra frame"""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_depth(local_depth: torch.Tensor, global_batch: int, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """local_depth [b_local, H, W] (this rank's shard, b_local from shard_range) -> [global_batch, H, W] on every rank.
    Even shards use a single all_gather_into_tensor straight into the output; uneven ones are padded to the largest
    shard first."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    H, W = local_depth.shape[1:]
    sizes = [shard_range(global_batch, r, world)[1] - shard_range(global_batch, r, world)[0] for r in range(world)]
    if local_depth.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local_depth.shape[0]} frames, expected {sizes[rank]}")
    if out is None:
        out = torch.empty((global_batch, H, W), dtype=local_depth.dtype, device=local_depth.device)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out, local_depth.contiguous(), group=group)
        return out
    mx = max(sizes)
    padded = torch.zeros((mx, H, W), dtype=local_depth.dtype, device=local_depth.device)
    padded[: sizes[rank]] = local_depth
    gathered = torch.empty((world * mx, H, W), dtype=local_depth.dtype, device=local_depth.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    pos = 0
    for r, n in enumerate(sizes):
        out[pos:pos + n] = gathered[r * mx:r * mx + n]
        pos += n
    return out


def sharded_forward(model, images_global: torch.Tensor, group=None) -> torch.Tensor:
    """Every rank passes the same [B, 3, H, W] batch (or at least its own slice filled in); each runs model() on its
    shard and the depth maps are all-gathered."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B = images_global.shape[0]
    lo, hi = shard_range(B, rank, world)
    local = model(images_global[lo:hi].contiguous())
    return all_gather_depth(local, B, group=group)
