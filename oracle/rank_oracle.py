"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the
reference's rank-precompute stage.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The shipped path
(``fusionocc_b200``) never does; it fails loudly when its CUDA library is
missing.

Parity status: PINNED.  Every function below is checked against golden
vectors produced by executing the reference's *own* code
(``projects/FusionOcc/fusionocc/necks/view_transformer.py``) in the build
container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``), and, when
``/root/reference`` is present, live against that code.

All citations are relative to ``/root/reference``.

Arithmetic notes
----------------
* numpy float32 ``-`` and ``/`` are IEEE-754 correctly rounded, identical to
  the torch CPU/CUDA elementwise kernels the reference runs, so the voxel
  index arithmetic here is bit-equivalent to the reference on either device.
* ``.long()`` is truncation toward zero -> ``np.trunc``.
* The FusionOcc copy concatenates an fp32 batch column onto the int64 voxel
  indices (view_transformer.py:249-251), which silently promotes everything to
  fp32; ranks are then exact only below 2**24.  ``rank_dtype='fp32'``
  reproduces that quirk, ``rank_dtype='int64'`` is the exact-integer mode the
  CUDA path implements (identical whenever B*Z*Y*X < 2**24).
* ``get_lidar_coor`` depends on third-party arithmetic (``torch.inverse`` +
  broadcast ``matmul``, torch pinned at 2.4.0 in docker/Dockerfile:41, 2.11 in
  this image).  It is restated op-for-op with the same torch calls.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- a1
def create_grid_infos(x, y, z, **kwargs):
    """view_transformer.py:87-103 — lower bound, interval, size as fp32[3]."""
    lb = np.array([cfg[0] for cfg in (x, y, z)], dtype=F32)
    itv = np.array([cfg[2] for cfg in (x, y, z)], dtype=F32)
    # Python-float division first, then cast to fp32 (torch.Tensor([...]) does that)
    size = np.array([(cfg[1] - cfg[0]) / cfg[2] for cfg in (x, y, z)], dtype=F32)
    return lb, itv, size


# --------------------------------------------------------------------------- a2
def _torch_cpu_linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """Bit-restatement of torch CPU ``linspace(dtype=float32)``.

    The ATen CPU kernel computes ``step=(end-start)/(steps-1)`` in fp32 and fills
    the first half as ``fma(step, i, start)`` and the second half as
    ``fma(-step, steps-1-i, end)`` (single rounding).  Emulated in fp64, where
    the product of an fp32 and a small integer is exact.
    """
    if steps == 1:
        return np.array([start], dtype=F32)
    s, e = F32(start), F32(end)
    step = F32((e - s) / F32(steps - 1))
    i = np.arange(steps, dtype=np.int64)
    lo = (np.float64(s) + np.float64(step) * i).astype(F32)
    hi = (np.float64(e) - np.float64(step) * (steps - 1 - i)).astype(F32)
    return np.where(i < steps // 2, lo, hi).astype(F32)


def create_frustum(depth_cfg, input_size, downsample, sid: bool = False) -> np.ndarray:
    """view_transformer.py:105-133 — (D,H,W,3) fp32 of (x_px, y_px, depth_m)."""
    H_in, W_in = input_size
    H_feat, W_feat = H_in // downsample, W_in // downsample
    lo, hi, st = depth_cfg
    # torch.arange(*cfg, dtype=float): count = ceil((hi-lo)/st) in double, value = lo + i*st
    n = int(np.ceil((float(hi) - float(lo)) / float(st)))
    d = (float(lo) + np.arange(n, dtype=np.float64) * float(st)).astype(F32)
    D = d.shape[0]
    if sid:
        # :121-126 — spacing-increasing discretisation, all in fp32 torch ops
        import torch
        d_sid = torch.arange(D).float()
        cfg_t = torch.tensor(depth_cfg).float()
        d_sid = torch.exp(torch.log(cfg_t[0]) + d_sid / (D - 1) *
                          torch.log((cfg_t[1] - 1) / cfg_t[0]))
        d = d_sid.numpy()
    xs = _torch_cpu_linspace_f32(0, W_in - 1, W_feat)
    ys = _torch_cpu_linspace_f32(0, H_in - 1, H_feat)
    fr = np.empty((D, H_feat, W_feat, 3), dtype=F32)
    fr[..., 0] = xs[None, None, :]
    fr[..., 1] = ys[None, :, None]
    fr[..., 2] = d[:, None, None]
    return fr


# --------------------------------------------------------------------------- a3
def get_lidar_coor(frustum, sensor2ego, ego2global, cam2imgs, post_rots, post_trans, bda):
    """view_transformer.py:135-173, op-for-op in torch (third-party arithmetic:
    torch.inverse + broadcast matmul).  ``ego2global`` is unused there too.
    Accepts numpy or torch inputs; returns a torch tensor (B,N,D,H,W,3)."""
    import torch
    as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a))
    frustum, sensor2ego, cam2imgs, post_rots, post_trans, bda = map(
        as_t, (frustum, sensor2ego, cam2imgs, post_rots, post_trans, bda))
    B, N, _, _ = sensor2ego.shape
    points = frustum.to(sensor2ego) - post_trans.view(B, N, 1, 1, 1, 3)            # :161
    points = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(points.unsqueeze(-1))  # :162-163
    points = torch.cat((points[..., :2, :] * points[..., 2:3, :], points[..., 2:3, :]), 5)    # :166-167
    combine = sensor2ego[:, :, :3, :3].matmul(torch.inverse(cam2imgs))                         # :168
    points = combine.view(B, N, 1, 1, 1, 3, 3).matmul(points).squeeze(-1)                     # :169
    points += sensor2ego[:, :, :3, 3].view(B, N, 1, 1, 1, 3)                                  # :170
    points = bda.view(B, 1, 1, 1, 1, 3, 3).matmul(points.unsqueeze(-1)).squeeze(-1)           # :171-172
    return points


# --------------------------------------------------------------------------- a4
def intervals_from_sorted(ranks: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """view_transformer.py:270-278 / bev_pool.py:50-57 — run starts and lengths."""
    n = ranks.shape[0]
    kept = np.ones(n, dtype=bool)
    kept[1:] = ranks[1:] != ranks[:-1]
    starts = np.nonzero(kept)[0].astype(np.int32)
    lengths = np.zeros_like(starts)
    if starts.shape[0]:
        lengths[:-1] = starts[1:] - starts[:-1]
        lengths[-1] = n - starts[-1]
    return starts, lengths


def voxel_index(coor: np.ndarray, lb: np.ndarray, itv: np.ndarray) -> np.ndarray:
    """view_transformer.py:246-248 — ((coor - lb) / itv).long(): fp32 sub, fp32
    true division, truncation toward zero.  Returns int64 (..., 3)."""
    c = (coor.astype(F32, copy=False) - lb.astype(F32)).astype(F32)
    c = (c / itv.astype(F32)).astype(F32)
    with np.errstate(invalid='ignore'):
        return np.trunc(c).astype(np.int64)


def voxel_pooling_prepare_v2(coor: np.ndarray, lb: np.ndarray, itv: np.ndarray,
                             grid_size: np.ndarray, rank_dtype: str = 'fp32'
                             ) -> Tuple[Optional[np.ndarray], ...]:
    """view_transformer.py:223-281.

    coor: fp32 (B,N,D,H,W,3).  Returns int32 arrays
    (ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths), or
    five Nones when nothing survives (:257-258, :274-275).  The sort is the
    stable one (ties in ascending point index), SURVEY.md §A.2.
    """
    B, N, D, H, W, _ = coor.shape
    num_points = B * N * D * H * W
    ranks_depth = np.arange(num_points, dtype=np.int32)                                  # :237-238
    ranks_feat = np.arange(num_points // D, dtype=np.int32).reshape(B, N, 1, H, W)       # :239-242
    ranks_feat = np.broadcast_to(ranks_feat, (B, N, D, H, W)).reshape(-1)
    idx = voxel_index(coor, lb, itv).reshape(num_points, 3)                              # :246-248
    batch_idx = np.repeat(np.arange(B, dtype=np.int64), num_points // B)                 # :249-250
    gs = grid_size.astype(F32)
    if rank_dtype == 'fp32':
        c = idx.astype(F32)                                                              # :251 (cat promotes to fp32)
        kept = ((c[:, 0] >= 0) & (c[:, 0] < gs[0]) & (c[:, 1] >= 0) & (c[:, 1] < gs[1]) &
                (c[:, 2] >= 0) & (c[:, 2] < gs[2]))                                      # :254-256
    else:
        gi = gs.astype(np.int64)
        kept = ((idx[:, 0] >= 0) & (idx[:, 0] < gi[0]) & (idx[:, 1] >= 0) & (idx[:, 1] < gi[1]) &
                (idx[:, 2] >= 0) & (idx[:, 2] < gi[2]))
    if kept.shape[0] == 0:                                                               # :257 (length, not population)
        return None, None, None, None, None
    ranks_depth, ranks_feat = ranks_depth[kept], ranks_feat[kept]                        # :259-260
    if rank_dtype == 'fp32':
        c = c[kept]
        b = batch_idx[kept].astype(F32)
        rb = (b * F32(gs[2] * gs[1] * gs[0])).astype(F32)                                # :262-263
        rb = (rb + (c[:, 2] * F32(gs[1] * gs[0])).astype(F32)).astype(F32)               # :264
        rb = (rb + ((c[:, 1] * gs[0]).astype(F32) + c[:, 0]).astype(F32)).astype(F32)    # :265
    else:
        i = idx[kept]
        gi = gs.astype(np.int64)
        rb = batch_idx[kept] * (gi[2] * gi[1] * gi[0]) + i[:, 2] * (gi[1] * gi[0]) + i[:, 1] * gi[0] + i[:, 0]
    order = np.argsort(rb, kind='stable')                                                # :266
    rb, ranks_depth, ranks_feat = rb[order], ranks_depth[order], ranks_feat[order]       # :267-268
    starts, lengths = intervals_from_sorted(rb)                                          # :270-278
    if starts.shape[0] == 0:                                                             # :274-275
        return None, None, None, None, None
    return (rb.astype(np.int32), ranks_depth.astype(np.int32), ranks_feat.astype(np.int32),
            starts.astype(np.int32), lengths.astype(np.int32))                           # :279-281


# --------------------------------------------------------------------------- a10 (backward re-sort)
def backward_resort(ranks_bev, ranks_depth, ranks_feat):
    """bev_pool.py:47-57 — stable argsort by ranks_feat of the bev-sorted arrays,
    then intervals over ranks_feat."""
    order = np.argsort(ranks_feat, kind='stable')
    rf, rd, rb = ranks_feat[order], ranks_depth[order], ranks_bev[order]
    starts, lengths = intervals_from_sorted(rf)
    return rb, rd, rf, starts, lengths
