"""ORACLE / CPU BASELINE (test + measurement infrastructure, NOT product code).

The "reference pure-PyTorch scatter CPU path" named by BASELINE.json
configs[0] and BASELINE.md §4: the rank precompute is the reference's own
eager-torch algorithm (``view_transformer.py:135-173,223-281``) restated
op-for-op on CPU tensors, and the pooling is the reference tree's own
pure-PyTorch pooling pattern (``occ_pool_pure_pytorch``,
``projects/CONet/mmdet3d_plugin/ops/occ_pooling/OCC_Pool.py:39-71``) adapted to
the v2 signature; backward is autograd.  Tolerance-grade (torch CPU rounds
the product before the add), not the bit-exact oracle — that is
``bevpool_oracle.c``.

Timed by ``bench.py`` (``cpu_baseline`` and ``--impl reference``), cross-checked
in ``tests/`` against the numpy/C oracle.
"""
from __future__ import annotations

import torch


def voxel_pooling_prepare_v2_torch(coor, grid_lower_bound, grid_interval, grid_size):
    """view_transformer.py:223-281 in eager torch (argsort made explicitly stable)."""
    B, N, D, H, W, _ = coor.shape
    num_points = B * N * D * H * W
    ranks_depth = torch.arange(0, num_points, dtype=torch.int, device=coor.device)
    ranks_feat = torch.arange(0, num_points // D, dtype=torch.int, device=coor.device)
    ranks_feat = ranks_feat.reshape(B, N, 1, H, W).expand(B, N, D, H, W).flatten()
    coor = ((coor - grid_lower_bound.to(coor)) / grid_interval.to(coor))
    coor = coor.long().view(num_points, 3)
    batch_idx = torch.arange(0, B, dtype=torch.float32, device=coor.device).reshape(B, 1) \
        .expand(B, num_points // B).reshape(num_points, 1)
    coor = torch.cat((coor, batch_idx), 1)
    gs = grid_size.to(coor.device)
    kept = (coor[:, 0] >= 0) & (coor[:, 0] < gs[0]) & (coor[:, 1] >= 0) & (coor[:, 1] < gs[1]) & \
           (coor[:, 2] >= 0) & (coor[:, 2] < gs[2])
    if len(kept) == 0:
        return None, None, None, None, None
    coor, ranks_depth, ranks_feat = coor[kept], ranks_depth[kept], ranks_feat[kept]
    ranks_bev = coor[:, 3] * (gs[2] * gs[1] * gs[0])
    ranks_bev += coor[:, 2] * (gs[1] * gs[0])
    ranks_bev += coor[:, 1] * gs[0] + coor[:, 0]
    order = ranks_bev.argsort(stable=True)
    ranks_bev, ranks_depth, ranks_feat = ranks_bev[order], ranks_depth[order], ranks_feat[order]
    kept = torch.ones(ranks_bev.shape[0], device=ranks_bev.device, dtype=torch.bool)
    kept[1:] = ranks_bev[1:] != ranks_bev[:-1]
    interval_starts = torch.where(kept)[0].int()
    if len(interval_starts) == 0:
        return None, None, None, None, None
    interval_lengths = torch.zeros_like(interval_starts)
    interval_lengths[:-1] = interval_starts[1:] - interval_starts[:-1]
    interval_lengths[-1] = ranks_bev.shape[0] - interval_starts[-1]
    return (ranks_bev.int().contiguous(), ranks_depth.int().contiguous(), ranks_feat.int().contiguous(),
            interval_starts.int().contiguous(), interval_lengths.int().contiguous())


def bev_pool_v2_pure_torch(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                           interval_starts=None, interval_lengths=None):
    """OCC_Pool.py:39-71 pattern with the v2 signature: gather, multiply,
    ``index_add_`` into (B,Z,Y,X,C), permute to contiguous (B,C,Z,Y,X).
    Differentiable through autograd."""
    C = feat.shape[-1]
    v = depth.reshape(-1)[ranks_depth.long(), None] * feat.reshape(-1, C)[ranks_feat.long()]
    out = torch.zeros(bev_feat_shape, dtype=v.dtype, device=v.device)
    out = out.view(-1, C).index_add(0, ranks_bev.long(), v).view(bev_feat_shape)
    return out.permute(0, 4, 1, 2, 3).contiguous()


def view_transform_step_cpu(frustum, calib, depth, feat_nchw, grid, out_grad=None):
    """One full pass of the hot path on CPU: geometry -> ranks -> splat (-> backward).

    ``grid`` = (lower_bound, interval, size) fp32 tensors.  Returns
    (out, depth_grad, feat_grad, n_kept, n_intervals)."""
    from .rank_oracle import get_lidar_coor
    lb, itv, gs = grid
    coor = get_lidar_coor(frustum, *calib)
    rb, rd, rf, st, ln = voxel_pooling_prepare_v2_torch(coor, lb, itv, gs)
    B = depth.shape[0]
    shape = (B, int(gs[2]), int(gs[1]), int(gs[0]), feat_nchw.shape[2])
    if out_grad is not None:
        depth = depth.detach().requires_grad_(True)
        feat_nchw = feat_nchw.detach().requires_grad_(True)
    feat = feat_nchw.permute(0, 1, 3, 4, 2)
    out = bev_pool_v2_pure_torch(depth, feat, rd, rf, rb, shape)
    dg = fg = None
    if out_grad is not None:
        out.backward(out_grad)
        dg, fg = depth.grad, feat_nchw.grad
    return out.detach(), dg, fg, int(rb.shape[0]), int(st.shape[0])
