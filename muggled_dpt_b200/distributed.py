"""
Multi-GPU plumbing for the depth path: frames of a batch never interact
(muggled_dpt/dpt_model.py:61-83 has no cross-batch op), so the batch is sharded over ranks - one process per GPU, the
weights replicated - and the only collective is ONE all-gather of the [B/G, H, W] depth maps per step.
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is plumbing, not compute.
"""

from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _native as N


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


class _NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


class NativeDepthAllGather:
    """The path's one collective through the C ABI: `dpt_allgather_depth` (include/dpt_b200.h) on a communicator this
    object owns - created with ncclCommInitRank on the NCCL library the process already carries (PyTorch's), the unique
    id exchanged once over the existing torch.distributed group. One process per GPU; every rank must construct it."""

    def __init__(self, group=None):
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._nccl = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)  # resolves to the copy torch has loaded
        self._nccl.ncclGetUniqueId.argtypes = [C.POINTER(_NcclUniqueId)]
        self._nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _NcclUniqueId, C.c_int]
        self._nccl.ncclCommDestroy.argtypes = [C.c_void_p]
        uid = _NcclUniqueId()
        payload = [None]
        if self.rank == 0:
            if self._nccl.ncclGetUniqueId(C.byref(uid)) != 0:
                raise RuntimeError("ncclGetUniqueId failed")
            payload[0] = C.string_at(C.byref(uid), 128)
        dist.broadcast_object_list(payload, src=0, group=group)
        C.memmove(C.byref(uid), payload[0], 128)
        self._comm = C.c_void_p()
        rc = self._nccl.ncclCommInitRank(C.byref(self._comm), self.world, uid, self.rank)
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed with ncclResult_t {rc}")

    def __call__(self, local_depth: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """local_depth [b, H, W] (same b on every rank) -> out [world * b, H, W], enqueued on the current stream"""
        if not (local_depth.is_contiguous() and out.is_contiguous()) or out.numel() != self.world * local_depth.numel():
            raise ValueError("all-gather buffers must be contiguous with out = world x local elements")
        code = {torch.bfloat16: N.DPT_BF16, torch.float16: N.DPT_F16}[local_depth.dtype]
        stream = C.c_void_p(torch.cuda.current_stream(local_depth.device).cuda_stream)
        rc = N.lib().dpt_allgather_depth(self._comm, C.c_void_p(local_depth.data_ptr()), C.c_void_p(out.data_ptr()),
                                         local_depth.numel(), code, stream)
        N.check(rc, None, "dpt_allgather_depth")
        return out

    def close(self):
        if self._comm:
            self._nccl.ncclCommDestroy(self._comm)
            self._comm = C.c_void_p()
