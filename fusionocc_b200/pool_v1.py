"""Sibling pooling ops on the same native pipeline.

``bev_pool(feats, coords, B, D, H, W)`` mirrors the BEVFusion / BEVDet v1 op
(``/root/reference/projects/BEVFusion/bevfusion/ops/bev_pool/bev_pool.py:85-99`` + ``src/bev_pool_cuda.cu:21-45``)
and ``occ_pool`` the CONet / SparseOcc occupancy pooling
(``projects/CONet/mmdet3d_plugin/ops/occ_pooling/OCC_Pool.py:39-71``): both sum pre-multiplied point
features ``feats[N, C]`` into the voxel ``(b, z, x, y)`` named by integer ``coords[N, 4] = (x, y, z, b)`` and
return fp32 ``(B, C, D, H, W)``.

Underneath: the stable bucket sort of the rank pipeline (``fo_rank_from_keys``) replaces the reference's
``argsort`` + ``kept`` / ``where`` interval rebuild, and the splat is the bev_pool_v2 forward with a unit depth
(``fma(x, 1, s) == s + x`` exactly, so the per-voxel sum is the reference kernel's sequential ``psum += x`` in
sorted order) writing the dense ``(B, C, D, H, W)`` tensor once — no zero fill, no permute copy.  The backward is
the bev_pool_v2 backward: ``feats.grad[i] = out_grad[voxel(i)]`` exactly (``bev_pool_cuda.cu:66-91``).
No CPU path.
"""
from __future__ import annotations

import torch

from . import _cabi
from .bev_pool import _p, _require_cuda, _stream, bev_pool_v2

__all__ = ['bev_pool', 'occ_pool', 'rank_from_keys']


def rank_from_keys(keys: torch.Tensor, n_buckets: int):
    """Stable sort of point indices by integer bucket id (fo_rank_from_keys).  Returns capacity-sized int32
    ``(sorted_keys, order, interval_starts, interval_lengths, counts_dev)``; live sizes are
    ``counts_dev[0]`` (kept points) and ``counts_dev[1]`` (intervals).  No host sync."""
    _require_cuda(keys)
    lib = _cabi.load()
    keys = keys.int().contiguous()
    n = keys.numel()
    dev = keys.device
    cap_iv = max(1, min(n, n_buckets))
    i32 = dict(dtype=torch.int32, device=dev)
    sorted_keys, order = torch.empty(max(n, 1), **i32), torch.empty(max(n, 1), **i32)
    starts, lengths = torch.empty(cap_iv, **i32), torch.empty(cap_iv, **i32)
    counts = torch.empty(4, **i32)
    sbytes = lib.fo_rank_from_keys_scratch_bytes(n, n_buckets)
    scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(lib.fo_rank_from_keys(_stream(dev), _p(keys), n, n_buckets, _p(sorted_keys), _p(order), _p(starts),
                                          _p(lengths), _p(counts), _p(scratch), sbytes), 'fo_rank_from_keys')
    return sorted_keys, order, starts, lengths, counts


def bev_pool(feats: torch.Tensor, coords: torch.Tensor, B: int, D: int, H: int, W: int) -> torch.Tensor:
    """Drop-in for ``bev_pool(feats, coords, B, D, H, W)`` (bev_pool.py:85-99): ``coords[:, (0,1,2,3)] =
    (x, y, z, batch)`` with ``x < H``, ``y < W``, ``z < D``; returns fp32 contiguous ``(B, C, D, H, W)``.
    Points with coordinates outside the grid are dropped (the reference requires the caller to filter them)."""
    assert feats.shape[0] == coords.shape[0]
    _require_cuda(feats, coords)
    B, D, H, W = (int(v.item()) if hasattr(v, 'item') else int(v) for v in (B, D, H, W))
    C = feats.shape[1]
    if feats.shape[0] == 0:
        return feats.new_zeros((B, C, D, H, W), dtype=torch.float32)
    c = coords.long()
    x, y, z, b = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    ok = (x >= 0) & (x < H) & (y >= 0) & (y < W) & (z >= 0) & (z < D) & (b >= 0) & (b < B)
    keys = torch.where(ok, ((b * D + z) * H + x) * W + y, torch.full_like(x, -1)).int()
    rb, order, st, ln, counts = rank_from_keys(keys, B * D * H * W)
    n_kept, n_iv = (int(v) for v in counts[:2].tolist())      # like the reference's torch.where (bev_pool.py:46)
    if n_kept == 0:
        return feats.new_zeros((B, C, D, H, W), dtype=torch.float32)
    rb, order, st, ln = rb[:n_kept], order[:n_kept], st[:n_iv], ln[:n_iv]
    ones = torch.ones(feats.shape[0], dtype=torch.float32, device=feats.device)
    return bev_pool_v2(ones, feats, order, order, rb, (B, D, H, W, C), st, ln)


def occ_pool(feats: torch.Tensor, coords: torch.Tensor, B, D, H, W) -> torch.Tensor:
    """Drop-in for ``occ_pool`` (OCC_Pool.py:74-104): same contract as :func:`bev_pool`."""
    assert feats.shape[0] == coords.shape[0], 'feats and coords must have same number of points'
    return bev_pool(feats, coords, B, D, H, W)
