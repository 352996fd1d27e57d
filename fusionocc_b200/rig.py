"""Synthetic nuScenes-shaped calibration + value rig.

Everything here is seeded and deterministic so that the oracle, the CUDA path,
the CPU baseline and the golden fixtures all see identical inputs.  Nothing in
this file touches the GPU unless the caller passes ``device=``.

Shapes follow the reference's ``img_inputs`` convention
(``projects/FusionOcc/fusionocc/necks/view_transformer.py:135-173``,
``transforms/loading.py:148-160``): ``sensor2ego (B,N,4,4)``,
``ego2global (B,N,4,4)``, ``cam2img (B,N,3,3)``, ``post_rots (B,N,3,3)``,
``post_trans (B,N,3)``, ``bda (B,3,3)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Tuple

import torch

# Grid of configs/fusion_occ.py:38-43
GRID_CONFIG_FUSIONOCC = {
    'x': [-40, 40, 0.4],
    'y': [-40, 40, 0.4],
    'z': [-1, 5.4, 0.4],
    'depth': [1.0, 45.0, 0.5],
}


@dataclass(frozen=True)
class Shape:
    """One BASELINE.json workload shape."""
    name: str
    input_size: Tuple[int, int]          # (H_in, W_in)
    downsample: int
    depth_cfg: Tuple[float, float, float]
    channels: int
    n_cams: int = 6
    grid_config: Dict[str, list] = field(default_factory=lambda: dict(GRID_CONFIG_FUSIONOCC))

    @property
    def feat_hw(self) -> Tuple[int, int]:
        return (self.input_size[0] // self.downsample, self.input_size[1] // self.downsample)

    @property
    def D(self) -> int:
        lo, hi, st = self.depth_cfg
        return int(torch.arange(lo, hi, st, dtype=torch.float).shape[0])

    def grid_cfg(self) -> Dict[str, list]:
        g = dict(self.grid_config)
        g['depth'] = list(self.depth_cfg)
        return g


SHAPES = {
    # BASELINE.json headline: 6 cams 256x704 -> 16x44, D=88, C=32
    'base': Shape('base', (256, 704), 16, (1.0, 45.0, 0.5), 32),
    # configs/fusion_occ.py:26 native resolution
    'native': Shape('native', (512, 1408), 16, (1.0, 45.0, 0.5), 32),
    # BASELINE.json configs[4]: 512x1408, D=118, C=80
    'stress': Shape('stress', (512, 1408), 16, (1.0, 60.0, 0.5), 80),
    # tiny frustum on the full grid, for CPU-speed tests
    'tiny': Shape('tiny', (64, 176), 16, (1.0, 45.0, 4.0), 8, n_cams=6),
    # small frustum on a coarse 50x50x4 grid (V=10 000, not a multiple of the 128-voxel tile; long intervals)
    'small': Shape('small', (128, 352), 16, (1.0, 45.0, 2.0), 32, n_cams=6,
                   grid_config={'x': [-40, 40, 1.6], 'y': [-40, 40, 1.6], 'z': [-1, 5.4, 1.6],
                                'depth': [1.0, 45.0, 2.0]}),
}

_YAW_DEG = [55.0, 0.0, -55.0, 110.0, 180.0, -110.0]   # FL, F, FR, BL, B, BR
_CAM_AXES = torch.tensor([[0., 0., 1.], [-1., 0., 0.], [0., -1., 0.]], dtype=torch.float64)


def _rz(yaw_rad: torch.Tensor) -> torch.Tensor:
    c, s = torch.cos(yaw_rad), torch.sin(yaw_rad)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.stack([c, -s, z], -1),
                        torch.stack([s, c, z], -1),
                        torch.stack([z, z, o], -1)], -2)


def make_calibration(shape: Shape, B: int, *, frame_shift: bool = False,
                     device='cpu') -> Tuple[torch.Tensor, ...]:
    """Returns (sensor2ego, ego2global, cam2img, post_rots, post_trans, bda) in fp32.

    Sample 0 is jitter-free (so B=1 reproduces the SURVEY probe numbers);
    samples b>=1 get a per-(b,cam) yaw jitter U(-2deg, 2deg) from
    ``torch.Generator().manual_seed(1000 + b)``.  ``frame_shift`` applies the
    adjacent-frame transform ``T(x=-2.5 m, yaw=+1deg)`` of config C4.
    Geometry is assembled in fp64 and cast to fp32, like ``prepare_inputs``
    (fusion_occ.py:245-248).
    """
    N = shape.n_cams
    H_in, W_in = shape.input_size
    yaw = torch.tensor([_YAW_DEG[i % 6] for i in range(N)], dtype=torch.float64).repeat(B, 1)
    for b in range(1, B):
        g = torch.Generator().manual_seed(1000 + b)
        yaw[b] += (torch.rand(N, generator=g, dtype=torch.float64) * 4.0 - 2.0)
    yaw_r = yaw * (math.pi / 180.0)
    R = _rz(yaw_r) @ _CAM_AXES                                     # (B,N,3,3)
    t = torch.stack([1.5 * torch.cos(yaw_r), 0.5 * torch.sin(yaw_r),
                     torch.full_like(yaw_r, 1.5)], -1)            # (B,N,3)
    s2e = torch.zeros(B, N, 4, 4, dtype=torch.float64)
    s2e[..., :3, :3] = R
    s2e[..., :3, 3] = t
    s2e[..., 3, 3] = 1.0
    if frame_shift:
        T = torch.eye(4, dtype=torch.float64)
        a = torch.tensor(math.pi / 180.0, dtype=torch.float64)
        T[:3, :3] = _rz(a)
        T[0, 3] = -2.5
        s2e = T @ s2e
    K = torch.zeros(B, N, 3, 3, dtype=torch.float64)
    f = torch.tensor([809.2 if (i % 6) == 4 else 1266.4 for i in range(N)], dtype=torch.float64)
    K[..., 0, 0] = f
    K[..., 1, 1] = f
    K[..., 0, 2] = 816.3
    K[..., 1, 2] = 491.5
    K[..., 2, 2] = 1.0
    resize = W_in / 1600.0
    crop_h = int(900 * resize) - H_in
    post_rots = torch.zeros(B, N, 3, 3, dtype=torch.float64)
    post_rots[..., 0, 0] = resize
    post_rots[..., 1, 1] = resize
    post_rots[..., 2, 2] = 1.0
    post_trans = torch.zeros(B, N, 3, dtype=torch.float64)
    post_trans[..., 1] = -float(crop_h)
    bda = torch.eye(3, dtype=torch.float64).repeat(B, 1, 1)
    e2g = torch.eye(4, dtype=torch.float64).repeat(B, N, 1, 1)
    out = (s2e, e2g, K, post_rots, post_trans, bda)
    return tuple(x.float().to(device) for x in out)


def make_values(shape: Shape, B: int, *, device='cpu', with_grad_seed: bool = True):
    """depth = softmax(N(0,1), dim=D) seed 0; feat ~ N(0,1) seed 1; out_grad ~ N(0,1) seed 2.

    Returns ``depth (B,N,D,H,W)``, ``feat (B,N,C,H,W)`` (NCHW, as the reference's
    ``tran_feat`` arrives, view_transformer.py:309-311) and ``out_grad (B,C,Z,Y,X)``
    (or None).  Generated per sample so that sample b is identical for every B.
    """
    N, D, C = shape.n_cams, shape.D, shape.channels
    H, W = shape.feat_hw
    depth = torch.empty(B, N, D, H, W)
    feat = torch.empty(B, N, C, H, W)
    for b in range(B):
        g0 = torch.Generator().manual_seed(0 + 7919 * b)
        g1 = torch.Generator().manual_seed(1 + 7919 * b)
        depth[b] = torch.randn(N, D, H, W, generator=g0).softmax(dim=1)
        feat[b] = torch.randn(N, C, H, W, generator=g1)
    return depth.to(device), feat.to(device)


def make_out_grad(B: int, C: int, Z: int, Y: int, X: int, *, device='cpu') -> torch.Tensor:
    """Upstream gradient of the (B,C,Z,Y,X) output, N(0,1), seed 2 (+ per-sample stride)."""
    g = torch.empty(B, C, Z, Y, X)
    for b in range(B):
        gen = torch.Generator().manual_seed(2 + 7919 * b)
        g[b] = torch.randn(C, Z, Y, X, generator=gen)
    return g.to(device)
