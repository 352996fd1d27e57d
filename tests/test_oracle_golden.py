"""CPU: the oracle against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py) and against the reference's known-answer test."""
import json
import os

import numpy as np
import pytest
import torch

from fusionocc_b200.rig import SHAPES, make_calibration
from oracle import kernels as ok
from oracle import rank_oracle as ro
from tests.helpers import canon, rig_case, sha

NAMES = ('ranks_bev', 'ranks_depth', 'ranks_feat', 'interval_starts', 'interval_lengths')


def test_kat_through_c_oracle(golden_dir):
    """mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176: loss 4.4, both gradients."""
    k = np.load(os.path.join(golden_dir, 'kat_bev_pool_v2.npz'))
    st, ln = ro.intervals_from_sorted(k['ranks_bev'])
    shape = tuple(int(v) for v in k['bev_feat_shape'])
    out = ok.bev_pool_v2(k['depth'], k['feat'], k['ranks_depth'], k['ranks_feat'], k['ranks_bev'], shape, st, ln)
    assert out.shape == (1, 2, 1, 2, 2)
    assert abs(float(out.sum()) - float(k['loss'])) < 1e-6
    dg, fg = ok.bev_pool_v2_backward(np.ones_like(out), k['depth'], k['feat'], k['ranks_depth'], k['ranks_feat'],
                                     k['ranks_bev'])
    np.testing.assert_allclose(dg, k['grad_depth'], rtol=0, atol=1e-7)
    np.testing.assert_allclose(fg, k['grad_feat'], rtol=0, atol=1e-7)


@pytest.mark.parametrize('fixture', ['geom_tiny.npz', 'geom_tiny_sid_aug.npz'])
def test_geometry_chain_matches_reference(golden_dir, fixture):
    """create_grid_infos / create_frustum / get_lidar_coor restatements == reference output, bit for bit."""
    g = np.load(os.path.join(golden_dir, fixture))
    sh = SHAPES['tiny']
    lb, itv, gs = ro.create_grid_infos(**sh.grid_cfg())
    if 'grid_size' in g:
        np.testing.assert_array_equal(lb, g['grid_lower_bound'])
        np.testing.assert_array_equal(itv, g['grid_interval'])
        np.testing.assert_array_equal(gs, g['grid_size'])
    fr = ro.create_frustum(sh.depth_cfg, sh.input_size, sh.downsample, sid='sid' in fixture)
    np.testing.assert_array_equal(fr, g['frustum'])
    coor = ro.get_lidar_coor(fr, g['sensor2ego'], g['ego2global'], g['cam2img'], g['post_rots'], g['post_trans'],
                             g['bda'])
    np.testing.assert_array_equal(coor.numpy(), g['coor'])


@pytest.mark.parametrize('fixture', ['geom_tiny.npz', 'geom_tiny_sid_aug.npz', 'edge_coords.npz', 'fp32_hazard_b28.npz'])
@pytest.mark.parametrize('mode', ['fp32', 'int64'])
def test_rank_oracle_matches_reference(golden_dir, fixture, mode):
    """voxel_pooling_prepare_v2 restatement == reference arrays.  The fp32 mode reproduces the
    reference everywhere (incl. its inexact ranks at B=28); the int64 mode agrees whenever
    B*Z*Y*X < 2^24 and must differ on the hazard fixture."""
    g = np.load(os.path.join(golden_dir, fixture))
    lb, itv, gs = ro.create_grid_infos(**(SHAPES['tiny'] if fixture.startswith('geom') else SHAPES['base']).grid_cfg())
    got = ro.voxel_pooling_prepare_v2(g['coor'], lb, itv, gs, mode)
    if fixture.startswith('fp32_hazard') and mode == 'int64':
        assert not np.array_equal(got[0], g['ranks_bev'])
        return
    for name, a in zip(NAMES, got):
        assert a.dtype == np.int32
        np.testing.assert_array_equal(a, g[name], err_msg=f'{fixture}:{name}')
    # the raw reference output (CPU argsort, unstable) is a tie-permutation of the canonical one
    crb, crd, crf = canon(g['ranks_bev'], g['ranks_depth_raw'], g['ranks_feat_raw'])
    np.testing.assert_array_equal(crd, g['ranks_depth'])
    np.testing.assert_array_equal(crf, g['ranks_feat'])


def test_all_filtered_returns_five_nones():
    lb, itv, gs = ro.create_grid_infos(**SHAPES['base'].grid_cfg())
    out = ro.voxel_pooling_prepare_v2(np.full((1, 1, 2, 2, 2, 3), 1e6, np.float32), lb, itv, gs)
    assert out == (None,) * 5


def test_fullsize_digest_base_b1(golden_dir):
    """Headline shape: geometry + ranks digests of the reference's own output."""
    with open(os.path.join(golden_dir, 'fullsize_digests.json')) as f:
        dig = json.load(f)['digests']['base_B1']
    c = rig_case('base', 1)
    assert sha(c['frustum']) == dig['frustum']
    assert sha(c['coor'].numpy()) == dig['coor']
    assert c['ranks'][0].shape[0] == dig['n_kept'] == 211434
    assert c['ranks'][3].shape[0] == dig['n_intervals'] == 138852
    for name, a in zip(NAMES, c['ranks']):
        assert sha(a) == dig[name], name


def test_torch_cpu_path_agrees_with_c_oracle():
    """The pure-PyTorch scatter path (CPU baseline) vs the bit-exact C oracle: rtol=atol=1e-5."""
    from oracle.torch_cpu_path import bev_pool_v2_pure_torch, voxel_pooling_prepare_v2_torch
    c = rig_case('small', 2)
    lb, itv, gs = (torch.from_numpy(x) for x in (c['lb'], c['itv'], c['gs']))
    t_ranks = voxel_pooling_prepare_v2_torch(c['coor'], lb, itv, gs)
    for a, b in zip(t_ranks, c['ranks']):
        np.testing.assert_array_equal(a.numpy(), b)
    rb, rd, rf, st, ln = c['ranks']
    B, N, D, H, W, _ = c['coor'].shape
    C = 16
    g = torch.Generator().manual_seed(0)
    depth = torch.rand(B, N, D, H, W, generator=g)
    feat = torch.randn(B, N, H, W, C, generator=g)
    X, Y, Z = (int(v) for v in c['gs'])
    shape = (B, Z, Y, X, C)
    want = ok.bev_pool_v2(depth.numpy(), feat.numpy(), rd, rf, rb, shape, st, ln)
    d = depth.clone().requires_grad_()
    f = feat.clone().requires_grad_()
    got = bev_pool_v2_pure_torch(d, f, t_ranks[1], t_ranks[2], t_ranks[0], shape)
    np.testing.assert_allclose(got.detach().numpy(), want, rtol=1e-5, atol=1e-5)
    og = torch.randn(got.shape, generator=g)
    got.backward(og)
    dg, fg = ok.bev_pool_v2_backward(og.numpy(), depth.numpy(), feat.numpy(), rd, rf, rb)
    np.testing.assert_allclose(d.grad.numpy(), dg, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(f.grad.numpy(), fg, rtol=1e-5, atol=1e-5)


def test_oracle_vs_reference_live():
    """When /root/reference is mounted (build container only): run the reference's own
    voxel_pooling_prepare_v2 / get_lidar_coor now and compare."""
    from tests.golden._ref_import import load_reference_view_transformer, reference_available
    if not reference_available():
        pytest.skip('reference tree not mounted')
    mod = load_reference_view_transformer('fusionocc')
    sh = SHAPES['small']
    vt = mod.LSSViewTransformer(grid_config=sh.grid_cfg(), input_size=sh.input_size, downsample=sh.downsample,
                                in_channels=8, out_channels=sh.channels, collapse_z=False)
    cal = make_calibration(sh, 3)
    coor = vt.get_lidar_coor(*cal)
    fr = ro.create_frustum(sh.depth_cfg, sh.input_size, sh.downsample)
    np.testing.assert_array_equal(ro.get_lidar_coor(fr, *cal).numpy(), coor.numpy())
    ref = [a.numpy() for a in vt.voxel_pooling_prepare_v2(coor)]
    lb, itv, gs = ro.create_grid_infos(**sh.grid_cfg())
    got = ro.voxel_pooling_prepare_v2(coor.numpy(), lb, itv, gs, 'fp32')
    crb, crd, crf = canon(ref[0], ref[1], ref[2])
    for a, b in zip(got, (crb, crd, crf, ref[3], ref[4])):
        np.testing.assert_array_equal(a, b)


def test_pool_v1_oracle_vs_reference_generated_golden(golden_dir):
    """The sibling-op oracle (oracle/pool_v1.py) against the output of the reference's own occ_pool_pure_pytorch
    (OCC_Pool.py:39-71, executed by tests/golden/make_golden.py): rtol=atol=1e-5 (index_add_ order is not defined)."""
    from oracle.pool_v1 import pool_v1
    g = np.load(os.path.join(golden_dir, 'occ_pool_ref.npz'))
    B, D, H, W = (int(v) for v in g['dims'])
    got = pool_v1(g['feats'], g['coords'], B, D, H, W)
    np.testing.assert_allclose(got, g['out'], rtol=1e-5, atol=1e-5)
