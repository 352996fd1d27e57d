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
import ctypes

from .bev_pool import (FO_LAYOUT_BCZYX, FO_LAYOUT_BZYXC, VoxelPoolPlan, _p, _require_cuda, _stream, native_backward,
                       native_forward)

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
    # no host read-back: the live counts stay on the device (the reference syncs in torch.where, bev_pool.py:46)
    n = feats.shape[0]
    plan = build_plan_from_counts(rb, st, ln, counts, B, D * H * W, n)
    return _UnitDepthPool.apply(feats, order, rb, st, ln, plan, (B, D, H, W, C))


def build_plan_from_counts(rb, st, ln, counts, B: int, n_vox: int, n_points: int) -> VoxelPoolPlan:
    """Forward plan over capacity-sized rank arrays whose live sizes are ``counts`` on the device."""
    lib = _cabi.load()
    dev = rb.device
    cap_iv = st.numel()
    nbytes = lib.fo_fwd_plan_bytes(B * n_vox, max(n_points, cap_iv))
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(lib.fo_fwd_plan_build(_stream(dev), _p(rb), _p(st), _p(ln), n_points, cap_iv,
                                          ctypes.c_void_p(counts.data_ptr() + 4), B, n_vox, _p(buf), nbytes),
                    'fo_fwd_plan_build')
    return VoxelPoolPlan(buf, B, n_vox, n_points, cap_iv, counts_dev=counts)


class _UnitDepthPool(torch.autograd.Function):
    """Sum of pre-multiplied point features per voxel = bev_pool_v2 with a unit depth; the gradient w.r.t. the
    (constant) depth vector is never formed into a tensor the caller sees."""

    @staticmethod
    def forward(ctx, feats, order, rb, st, ln, plan, shape):
        feats32 = feats.contiguous().float()
        ones = torch.ones(feats32.shape[0], dtype=torch.float32, device=feats32.device)
        out = native_forward(ones, feats32, order, order, rb, st, ln, shape, plan)
        ctx.save_for_backward(ones, feats32, order)
        ctx.plan, ctx.shape, ctx.in_dtype = plan, shape, feats.dtype
        return out

    @staticmethod
    def backward(ctx, out_grad):
        ones, feats32, order = ctx.saved_tensors
        if out_grad.dtype != torch.float32:
            out_grad = out_grad.float()
        if out_grad.is_contiguous():
            layout = FO_LAYOUT_BCZYX
        elif out_grad.permute(0, 2, 3, 4, 1).is_contiguous():
            layout = FO_LAYOUT_BZYXC
        else:
            out_grad, layout = out_grad.contiguous(), FO_LAYOUT_BCZYX
        _dg, fg = native_backward(out_grad, layout, ones, feats32, order, order, ctx.shape, ctx.plan)
        return fg.to(ctx.in_dtype), None, None, None, None, None, None


def occ_pool(feats: torch.Tensor, coords: torch.Tensor, B, D, H, W) -> torch.Tensor:
    """Drop-in for ``occ_pool`` (OCC_Pool.py:74-104): same contract as :func:`bev_pool`."""
    assert feats.shape[0] == coords.shape[0], 'feats and coords must have same number of points'
    return bev_pool(feats, coords, B, D, H, W)
