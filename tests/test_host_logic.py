"""CPU: host-side mirror of the reference interface (no CUDA needed)."""
import sys
import os

import numpy as np
import pytest
import torch

from fusionocc_b200.rig import SHAPES, make_calibration, make_values
from oracle import rank_oracle as ro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _vt(name='tiny', **kw):
    from fusionocc_b200 import LSSViewTransformer
    sh = SHAPES[name]
    return sh, LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8,
                                  out_channels=sh.channels, **kw)


@pytest.mark.parametrize('name', ['tiny', 'base', 'stress'])
def test_module_surface_and_geometry_match_oracle(name):
    """Attribute names / values of view_transformer.py:61-133 and get_lidar_coor (:135-173) on CPU."""
    sh, vt = _vt(name, collapse_z=False)
    lb, itv, gs = ro.create_grid_infos(**sh.grid_cfg())
    np.testing.assert_array_equal(vt.grid_lower_bound.numpy(), lb)
    np.testing.assert_array_equal(vt.grid_interval.numpy(), itv)
    np.testing.assert_array_equal(vt.grid_size.numpy(), gs)
    assert vt.D == sh.D and vt.initial_flag is True and vt.accelerate is False
    np.testing.assert_array_equal(vt.frustum.numpy(), ro.create_frustum(sh.depth_cfg, sh.input_size, sh.downsample))
    for attr in ('grid_config', 'downsample', 'out_channels', 'in_channels', 'depth_net', 'collapse_z', 'sid'):
        assert hasattr(vt, attr)
    for meth in ('create_grid_infos', 'create_frustum', 'get_lidar_coor', 'init_acceleration_v2', 'voxel_pooling_v2',
                 'voxel_pooling_prepare_v2', 'pre_compute', 'view_transform_core', 'view_transform', 'forward',
                 'get_mlp_input'):
        assert callable(getattr(vt, meth))
    if name == 'tiny':
        cal = make_calibration(sh, 2)
        got = vt.get_lidar_coor(*cal)
        want = ro.get_lidar_coor(vt.frustum.numpy(), *cal)
        assert torch.equal(got, want)
        # STCOcc variant: 4x4 bda with translation
        bda4 = torch.eye(4).repeat(2, 1, 1)
        bda4[:, :3, 3] = torch.tensor([1.0, -2.0, 0.5])
        got4 = vt.get_lidar_coor(cal[0], cal[1], cal[2], cal[3], cal[4], bda4)
        assert torch.allclose(got4, want + torch.tensor([1.0, -2.0, 0.5]))


def test_sid_frustum():
    sh, vt = _vt('tiny', sid=True)
    np.testing.assert_array_equal(vt.frustum.numpy(),
                                  ro.create_frustum(sh.depth_cfg, sh.input_size, sh.downsample, sid=True))


def test_ops_refuse_cpu_tensors_loudly():
    """No CPU fallback: the op and the rank precompute raise on CPU tensors."""
    from fusionocc_b200 import bev_pool_v2, rank_prepare
    z = torch.zeros
    with pytest.raises(RuntimeError, match='CUDA'):
        bev_pool_v2(z(1, 1, 2, 2, 2), z(1, 1, 2, 2, 4), z(1).int(), z(1).int(), z(1).int(), (1, 1, 2, 2, 4), z(1).int(),
                    z(1).int())
    with pytest.raises(RuntimeError, match='CUDA'):
        rank_prepare(z(1, 1, 2, 2, 2, 3), [0, 0, 0], [1, 1, 1], [4, 4, 4])
    sh, vt = _vt('tiny')
    with pytest.raises(RuntimeError, match='CUDA'):
        vt.voxel_pooling_prepare_v2(z(1, 1, 2, 2, 2, 3))


def test_out_grad_layout_classification():
    from fusionocc_b200.bev_pool import FO_LAYOUT_BCZYX, FO_LAYOUT_BZYXC, _classify_out_grad
    B, Z, Y, X, C = 2, 3, 4, 5, 8
    g_bczyx = torch.randn(B, C, Z, Y, X)
    view = g_bczyx.permute(0, 2, 3, 4, 1)                  # what autograd hands QuickCumsumCuda.backward
    t, lay = _classify_out_grad(view)
    assert lay == FO_LAYOUT_BCZYX and t.data_ptr() == g_bczyx.data_ptr()
    g_cl = torch.randn(B, Z, Y, X, C)
    t, lay = _classify_out_grad(g_cl)
    assert lay == FO_LAYOUT_BZYXC and t.data_ptr() == g_cl.data_ptr()
    odd = torch.randn(B, Z, Y, C, X).permute(0, 1, 2, 4, 3)
    t, lay = _classify_out_grad(odd)
    assert lay == FO_LAYOUT_BCZYX and t.permute(0, 4, 1, 2, 3).is_contiguous() and torch.equal(t, odd)
    t, lay = _classify_out_grad(view.half())
    assert t.dtype == torch.float32


def test_overlay_import_paths():
    """The overlay keeps the reference's import paths: mmdet3d.ops.bev_pool_v2.bev_pool (+ LiCROcc copy)."""
    sys.path.insert(0, os.path.join(ROOT, 'overlay'))
    try:
        for m in [k for k in sys.modules if k == 'mmdet3d' or k.startswith('mmdet3d.')]:
            del sys.modules[m]
        from mmdet3d.ops.bev_pool_v2.bev_pool import TRTBEVPoolv2, bev_pool_v2
        from mmdet3d.ops.bev_pool_v2 import bev_pool_v2 as again
        import fusionocc_b200
        assert bev_pool_v2 is fusionocc_b200.bev_pool_v2 is again and TRTBEVPoolv2 is fusionocc_b200.TRTBEVPoolv2
        import importlib.util
        p = os.path.join(ROOT, 'overlay', 'projects', 'LiCROcc', 'projects', 'mmdet3d_plugin', 'ops', 'bev_pool_v2',
                         'bev_pool.py')
        spec = importlib.util.spec_from_file_location('licrocc_bev_pool', p)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert mod.bev_pool_v2 is fusionocc_b200.bev_pool_v2
        # LiCROcc's vendored TRTBEVPoolv2 has its own signature (bev_pool.py:108-159 of that project)
        import inspect
        assert mod.TRTBEVPoolv2 is not TRTBEVPoolv2
        assert list(inspect.signature(mod.TRTBEVPoolv2.forward).parameters)[-3:] == \
            ['output_height', 'output_width', 'output_z']
        assert list(inspect.signature(TRTBEVPoolv2.forward).parameters)[-2:] == ['out_height', 'out_width']
    finally:
        sys.path.remove(os.path.join(ROOT, 'overlay'))
        for m in [k for k in sys.modules if k == 'mmdet3d' or k.startswith('mmdet3d.')]:
            del sys.modules[m]


def test_rig_is_deterministic_and_sample_stable():
    sh = SHAPES['tiny']
    a = make_calibration(sh, 3)
    b = make_calibration(sh, 3)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    c1 = make_calibration(sh, 1)
    for x, y in zip(a, c1):
        assert torch.equal(x[:1], y), 'sample 0 must not depend on the batch size'
    d3, f3 = make_values(sh, 3)
    d1, f1 = make_values(sh, 1)
    assert torch.equal(d3[:1], d1) and torch.equal(f3[:1], f1)
    assert torch.allclose(d3.sum(2), torch.ones(3, sh.n_cams, *sh.feat_hw))


def test_bench_byte_model_matches_baseline_md():
    """bench.py's algorithmic-byte formula reproduces BASELINE.md's per-sample table."""
    sys.path.insert(0, ROOT)
    import bench
    ab = bench.algorithmic_bytes(1, 6, 88, 16, 44, 32, 640000, 211434, 138852)
    assert round(ab['fwd'] / 1e6, 2) == 87.60 and round(ab['bwd'] / 1e6, 2) == 24.40
    assert round(ab['pre'] / 1e6, 2) == 8.11 and round(ab['total'] / 1e6, 2) == 120.10


def test_new_entry_points_have_no_cpu_path():
    """Sibling ops, the channel-slice op and the fused-geometry precompute raise on CPU tensors too."""
    from fusionocc_b200 import bev_pool_v2_cat, pack_calibration, rank_prepare_calib
    from fusionocc_b200.pool_v1 import bev_pool, occ_pool, rank_from_keys
    z = torch.zeros
    with pytest.raises(RuntimeError, match='CUDA'):
        bev_pool(z(4, 8), z(4, 4).int(), 1, 1, 2, 2)
    with pytest.raises(RuntimeError, match='CUDA'):
        occ_pool(z(4, 8), z(4, 4).int(), 1, 1, 2, 2)
    with pytest.raises(RuntimeError, match='CUDA'):
        rank_from_keys(z(4).int(), 8)
    frame = (z(1, 1, 2, 2, 2), z(1, 1, 2, 2, 4), z(1).int(), z(1).int(), z(1).int(), z(1).int(), z(1).int())
    with pytest.raises(RuntimeError, match='CUDA'):
        bev_pool_v2_cat([frame, frame], (1, 1, 2, 2, 4))
    with pytest.raises(ValueError):
        bev_pool_v2_cat([], (1, 1, 2, 2, 4))
    sh, vt = _vt('tiny')
    cal = make_calibration(sh, 1)
    cam, bda12, has_t = pack_calibration(cal[0], cal[2], cal[3], cal[4], cal[5])
    with pytest.raises(RuntimeError, match='CUDA'):
        rank_prepare_calib(vt.frustum, cam, bda12, has_t, 1, sh.n_cams, [0, 0, 0], [1, 1, 1], [4, 4, 4])


def test_pack_calibration_layout_and_formula():
    """cam_mats / bda12 are the matrices of view_transformer.py:161-172; evaluating the per-point formula with
    them in fp64 reproduces the oracle's get_lidar_coor to fp32 rounding (3x3 and 4x4 bda)."""
    from fusionocc_b200 import pack_calibration
    from oracle import rank_oracle as ro
    sh, vt = _vt('tiny')
    B = 2
    s2e, e2g, k, pr, pt, bda = make_calibration(sh, B, frame_shift=True)
    for use_bda4 in (False, True):
        b_in = bda
        if use_bda4:
            b_in = torch.eye(4).repeat(B, 1, 1)
            b_in[:, :3, :3] = bda * 1.02
            b_in[:, :3, 3] = torch.tensor([0.25, -0.5, 0.125])
        cam, bda12, has_t = pack_calibration(s2e, k, pr, pt, b_in)
        N = s2e.shape[1]
        assert cam.shape == (B * N, 24) and bda12.shape == (B, 12) and has_t == use_bda4
        assert torch.equal(cam[:, :9].reshape(B, N, 3, 3), torch.inverse(pr))
        assert torch.equal(cam[:, 9:12].reshape(B, N, 3), pt)
        assert torch.equal(cam[:, 21:24].reshape(B, N, 3), s2e[:, :, :3, 3])
        fr = vt.frustum.double()                                   # (D,H,W,3)
        c = cam.double().reshape(B, N, 24)
        p = fr[None, None] - c[:, :, None, None, None, 9:12]
        p = torch.einsum('bnij,bndhwj->bndhwi', c[..., :9].reshape(B, N, 3, 3), p)
        p = torch.cat((p[..., :2] * p[..., 2:3], p[..., 2:3]), -1)
        p = torch.einsum('bnij,bndhwj->bndhwi', c[..., 12:21].reshape(B, N, 3, 3), p) + c[:, :, None, None, None, 21:24]
        bd = bda12.double()
        p = torch.einsum('bij,bndhwj->bndhwi', bd[:, :9].reshape(B, 3, 3), p) + bd[:, None, None, None, None, 9:12]
        want = vt.get_lidar_coor(s2e, e2g, k, pr, pt, b_in)
        assert torch.allclose(p.float(), want, rtol=1e-5, atol=1e-4)
        if not use_bda4:
            ref = ro.get_lidar_coor(ro.create_frustum(sh.depth_cfg, sh.input_size, sh.downsample), s2e, e2g, k, pr, pt, bda)
            assert torch.allclose(p.float(), ref, rtol=1e-5, atol=1e-4)


def test_overlay_import_paths_sibling_ops():
    """overlay/projects/{BEVFusion,CONet} keep the sibling ops' module layout (bev_pool.py:85, OCC_Pool.py:74)."""
    import importlib.util
    import fusionocc_b200.pool_v1 as pv1
    for rel, name in ((('projects', 'BEVFusion', 'bevfusion', 'ops', 'bev_pool', 'bev_pool.py'), 'bev_pool'),
                      (('projects', 'CONet', 'mmdet3d_plugin', 'ops', 'occ_pooling', 'OCC_Pool.py'), 'occ_pool')):
        p = os.path.join(ROOT, 'overlay', *rel)
        spec = importlib.util.spec_from_file_location('overlay_' + name, p)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert getattr(mod, name) is getattr(pv1, name)
