"""get_lidar_coor fused into the rank precompute (fo_rank_prepare_calib).

The per-point arithmetic of view_transformer.py:161-172 is three 3x3 matrix-vector products that the
reference runs through a batched library GEMM, whose fp32 summation order is not documented.  These tests
MEASURE, on the GPU, how each of four candidate orders of the fused kernel compares with the reference's torch ops:
bitwise equality of the frustum points, and the voxel-index mismatch rate that survives truncation.  The
default order must reproduce every rank array exactly on the rig shapes; whatever the order, the fused path
must stay consistent with itself (ranks from the fused call == ranks from fo_rank_prepare on its own points).
"""
import json
import os

import numpy as np
import pytest
import torch

from fusionocc_b200 import LSSViewTransformer
from fusionocc_b200.rig import SHAPES, make_calibration, make_values
from fusionocc_b200.view_transformer import (DEFAULT_MATVEC_MODE, pack_calibration, rank_prepare,
                                             rank_prepare_calib)

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _setup(name, B, frame_shift=False, bda4=False):
    sh = SHAPES[name]
    vt = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels,
                            collapse_z=False)
    cal = [c.to(DEV) for c in make_calibration(sh, B, frame_shift=frame_shift)]
    if bda4:                                     # STCOcc's 4x4 bda with a translation and a small rotation/flip
        g = torch.Generator().manual_seed(5)
        bda = torch.eye(4).repeat(B, 1, 1)
        ang = (torch.rand(B, generator=g) - 0.5) * 0.4
        bda[:, 0, 0], bda[:, 0, 1], bda[:, 1, 0], bda[:, 1, 1] = ang.cos(), -ang.sin(), ang.sin(), ang.cos()
        bda[:, :3, :3] *= 1.05
        bda[:, :3, 3] = torch.rand(B, 3, generator=g) - 0.5
        cal[5] = bda.to(DEV)
        c, s_ = float(np.cos(0.05)), float(np.sin(0.05))      # train-time image rotation: full 2x2 block in post_rots
        rot = torch.tensor([[c, -s_, 0.], [s_, c, 0.], [0., 0., 1.]], device=DEV)
        cal[3] = cal[3] @ rot
    return sh, vt, cal


def _fused(vt, cal, mode, return_coor=True):
    s2e, _e2g, k, pr, pt, bda = cal
    B, N = s2e.shape[:2]
    cam, bda12, has_t = pack_calibration(s2e, k, pr, pt, bda)
    return rank_prepare_calib(vt._frustum_on(s2e), cam, bda12, has_t, B, N, vt.grid_lower_bound.tolist(),
                              vt.grid_interval.tolist(), vt._grid_xyz(), matvec_mode=mode, return_coor=return_coor)


def _live(arrs, counts):
    nk, ni = (int(v) for v in counts[:2].tolist())
    rb, rd, rf, st, ln = arrs
    return [x.cpu().numpy() for x in (rb[:nk], rd[:nk], rf[:nk], st[:ni], ln[:ni])]


@pytest.mark.parametrize('name,B,bda4', [('small', 2, False), ('base', 2, False), ('base', 1, True), ('stress', 1, True)])
def test_matvec_order_vs_reference_ops(name, B, bda4):
    """Measured comparison of the three summation orders with torch's get_lidar_coor on this GPU."""
    sh, vt, cal = _setup(name, B, bda4=bda4)
    want = vt.get_lidar_coor(*cal).contiguous()
    ref = _live(rank_prepare(want, vt.grid_lower_bound.tolist(), vt.grid_interval.tolist(), vt._grid_xyz())[:5],
                rank_prepare(want, vt.grid_lower_bound.tolist(), vt.grid_interval.tolist(), vt._grid_xyz())[5])
    report = {}
    for mode in (0, 1, 2, 3):
        out = _fused(vt, cal, mode)
        coor = out[7]
        diff_bits = int((coor.view(torch.int32) != want.view(torch.int32)).sum())
        got = _live(out[:5], out[5])
        same_ranks = all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(got, ref))
        # voxel-index mismatches: points whose kept flag or voxel differs
        lb, itv = vt.grid_lower_bound.to(DEV), vt.grid_interval.to(DEV)
        ia, ib = ((coor - lb) / itv).long(), ((want - lb) / itv).long()
        vox_diff = int((ia != ib).any(dim=-1).sum())
        report[mode] = dict(coor_floats_differing=diff_bits, of=coor.numel(), voxel_index_mismatches=vox_diff,
                            points=coor.numel() // 3, rank_arrays_identical=bool(same_ranks),
                            max_abs_diff=float((coor - want).abs().max()))
        # every order is a correct evaluation of the same formula: differences are rounding-level
        assert torch.allclose(coor, want, rtol=1e-5, atol=1e-4), (mode, report[mode])
        assert vox_diff <= 1e-3 * (coor.numel() // 3), (mode, report[mode])
    print(f'\n[fused-geometry] {name} B={B} bda4={bda4}: ' + json.dumps(report))
    os.makedirs('gpurun_out', exist_ok=True)
    with open(f'gpurun_out/fused_geometry_{name}_B{B}_{int(bda4)}.json', 'w') as f:
        json.dump(report, f, indent=1)
    d = report[DEFAULT_MATVEC_MODE]
    assert d['coor_floats_differing'] == 0 and d['rank_arrays_identical'], \
        f'default matvec_mode {DEFAULT_MATVEC_MODE} no longer reproduces the reference ops bit for bit: {report}'


@pytest.mark.parametrize('mode', [0, 1, 2, 3])
def test_fused_call_is_self_consistent(mode):
    """Ranks from the fused call == ranks from fo_rank_prepare run on the points the fused call emitted."""
    sh, vt, cal = _setup('base', 2, frame_shift=True)
    out = _fused(vt, cal, mode)
    got = _live(out[:5], out[5])
    sep = rank_prepare(out[7].contiguous(), vt.grid_lower_bound.tolist(), vt.grid_interval.tolist(), vt._grid_xyz())
    want = _live(sep[:5], sep[5])
    for a, b, nm in zip(got, want, ('ranks_bev', 'ranks_depth', 'ranks_feat', 'interval_starts', 'interval_lengths')):
        assert np.array_equal(a, b), nm


def test_module_fuse_geometry_matches_unfused_module():
    """LSSViewTransformer(fuse_geometry=True).view_transform == the unfused module, bit for bit, fwd and bwd."""
    sh, vt, cal = _setup('small', 2)
    vt_f = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels,
                              collapse_z=False, fuse_geometry=True)
    depth, feat = (t.to(DEV) for t in make_values(sh, 2))
    B, N, C, H, W = feat.shape
    inp = [torch.empty(B, N, 8, H, W, device=DEV)] + cal
    outs = []
    for m in (vt, vt_f):
        d = depth.clone().requires_grad_()
        f = feat.clone().requires_grad_()
        out, _ = m.view_transform(inp, d.view(B * N, -1, H, W), f.view(B * N, C, H, W))
        out.backward(torch.ones_like(out) * 0.5)
        outs.append((out.detach(), d.grad, f.grad))
    for a, b, nm in zip(outs[0], outs[1], ('out', 'depth_grad', 'feat_grad')):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), nm
