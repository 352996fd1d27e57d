"""Multi-GPU layer for the view transformation: batch sharding + output gather.

The path shards by independent units: ``ranks_bev = b * (Z*Y*X) + ...`` so samples (and
temporal frames, which FusionOcc treats as extra independent pooling calls,
``fusion_occ.py:289-326``) never share a voxel.  One process per GPU
(``torchrun``), each rank runs rank precompute + forward (+ backward) on its own
contiguous slice of the batch with *local* batch indices; there is NO collective
on the data path.  BASELINE.json names one collective only: gathering the voxel
outputs, which is what :func:`gather_outputs` does (NCCL ``all_gather`` /
``gather`` over NVLink; ``gloo`` in the CPU tests).  In real DDP training the
outputs are consumed locally and never gathered.

Nothing here touches CUDA directly, so the logic is testable with ``gloo`` on CPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ['shard_bounds', 'shard_batch', 'gather_outputs', 'sharded_view_transform']


def shard_bounds(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of ``batch`` samples owned by ``rank``.
    The first ``batch % world_size`` ranks get one extra sample; slices may be empty."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank {rank} / world_size {world_size}')
    base, rem = divmod(batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[Optional[torch.Tensor]], world_size: int, rank: int) -> List[Optional[torch.Tensor]]:
    """Slice every per-sample tensor (leading dim = global batch) to this rank's samples."""
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        lo, hi = shard_bounds(t.shape[0], world_size, rank)
        out.append(t[lo:hi])
    return out


def gather_outputs(local: torch.Tensor, global_batch: int, dst: Optional[int] = None,
                   group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
    """Assemble the per-rank ``(B_local, C, Z, Y, X)`` voxel tensors into ``(B, C, Z, Y, X)``.

    ``dst=None``: every rank gets the full tensor (all_gather); otherwise only ``dst`` does
    (gather) and the others return ``None``.  Ragged shards (batch not divisible by the world
    size) are padded to the largest shard for the collective and trimmed afterwards.
    """
    if not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(global_batch, world, r) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    if local.shape[0] != counts[rank]:
        raise ValueError(f'rank {rank} holds {local.shape[0]} samples, expected {counts[rank]}')
    max_n = max(counts)
    local = local.contiguous()
    if max_n == 0:
        return local if (dst is None or dst == rank) else None
    if local.shape[0] < max_n:
        pad = local.new_zeros((max_n - local.shape[0],) + tuple(local.shape[1:]))
        send = torch.cat([local, pad], 0)
    else:
        send = local
    if dst is None:
        if len(set(counts)) == 1:
            out = local.new_empty((world * max_n,) + tuple(local.shape[1:]))
            dist.all_gather_into_tensor(out, send, group=group)
            return out
        bufs = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(bufs, send, group=group)
        return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def sharded_view_transform(view_transformer, input: Sequence[torch.Tensor], depth: torch.Tensor,
                           tran_feat: torch.Tensor, gather: bool = False, dst: Optional[int] = None):
    """Run ``view_transformer.view_transform`` on this rank's batch slice.

    ``input`` is the reference's list ``[x(B,N,C,H,W), sensor2ego, ego2global, cam2img,
    post_rots, post_trans, bda, ...]`` for the GLOBAL batch; ``depth`` / ``tran_feat`` are
    ``(B*N, D|C, H, W)``.  Returns this rank's ``bev_feat`` (or the gathered tensor)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    B, N = input[0].shape[:2]
    lo, hi = shard_bounds(B, world, rank)
    local_in = [t[lo:hi] if isinstance(t, torch.Tensor) and t.shape[0] == B else t for t in input]
    d = depth.view(B, N, *depth.shape[1:])[lo:hi].reshape(-1, *depth.shape[1:])
    f = tran_feat.view(B, N, *tran_feat.shape[1:])[lo:hi].reshape(-1, *tran_feat.shape[1:])
    bev, _ = view_transformer.view_transform(local_in, d, f)
    if gather:
        return gather_outputs(bev, B, dst=dst)
    return bev
