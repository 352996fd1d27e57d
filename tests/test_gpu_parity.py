"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle and the golden
fixtures.  Integer/index outputs must be exactly equal; fp32 outputs are compared BIT-EXACTLY
wherever the reference's summation order is preserved (all of forward, depth_grad and feat_grad on
the vector and scalar paths), and the tolerance north_star states (rtol = atol = 1e-5) is written
where a looser comparison is the contract (pure-torch scatter path).
"""
import json
import os

import numpy as np
import pytest
import torch

from tests.helpers import canon, rig_case, sha

pytestmark = pytest.mark.gpu

RTOL = ATOL = 1e-5          # BASELINE.json north_star tolerance for fp32


def dev():
    return torch.device('cuda:0')


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        x = x.to(dtype)
    return x.to(dev())


def bits(x: torch.Tensor) -> np.ndarray:
    return x.detach().cpu().contiguous().numpy().view(np.uint32)


def assert_bit_equal(got: torch.Tensor, want: np.ndarray, what: str):
    g = got.detach().cpu().contiguous().numpy()
    assert g.shape == want.shape, f'{what}: shape {g.shape} vs {want.shape}'
    # -0.0 vs +0.0 matter too: compare raw bits
    same = g.view(np.uint32) == np.ascontiguousarray(want).view(np.uint32)
    if not same.all():
        idx = np.argwhere(~same)[:5]
        raise AssertionError(f'{what}: {(~same).sum()} of {same.size} elements differ bitwise, first at {idx.tolist()}: '
                             f'got {g[tuple(idx[0])]!r} want {want[tuple(idx[0])]!r}')


# ------------------------------------------------------------------------------------------ KAT
def test_kat_reference_known_answer(golden_dir):
    """The reference's only test of this path, mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176."""
    from fusionocc_b200 import bev_pool_v2
    k = np.load(os.path.join(golden_dir, 'kat_bev_pool_v2.npz'))
    depth = t(k['depth']).requires_grad_()
    feat = t(k['feat']).requires_grad_()
    rd, rf, rb = t(k['ranks_depth']), t(k['ranks_feat']), t(k['ranks_bev'])
    kept = torch.ones(rb.shape[0], device=rb.device, dtype=torch.bool)
    kept[1:] = rb[1:] != rb[:-1]
    st = torch.where(kept)[0].int()
    ln = torch.zeros_like(st)
    ln[:-1] = st[1:] - st[:-1]
    ln[-1] = rb.shape[0] - st[-1]
    out = bev_pool_v2(depth, feat, rd, rf, rb, (1, 1, 2, 2, 2), st, ln)
    assert out.shape == (1, 2, 1, 2, 2) and out.is_contiguous() and out.dtype == torch.float32
    loss = torch.sum(out)
    loss.backward()
    assert loss.item() == pytest.approx(4.4, abs=1e-6)
    assert torch.allclose(depth.grad.cpu(), torch.from_numpy(k['grad_depth']))
    assert torch.allclose(feat.grad.cpu(), torch.from_numpy(k['grad_feat']))


# ------------------------------------------------------------------------------------------ ranks
def _run_rank_prepare(coor_np, lb, itv, gs):
    from fusionocc_b200 import rank_prepare
    coor = t(coor_np)
    rb, rd, rf, st, ln, counts, plan = rank_prepare(coor, lb.tolist(), itv.tolist(), [int(v) for v in gs])
    nk, ni = (int(v) for v in counts[:2].tolist())
    return (rb[:nk].cpu().numpy(), rd[:nk].cpu().numpy(), rf[:nk].cpu().numpy(), st[:ni].cpu().numpy(),
            ln[:ni].cpu().numpy()), plan


@pytest.mark.parametrize('fixture', ['geom_tiny.npz', 'geom_tiny_sid_aug.npz', 'edge_coords.npz'])
def test_rank_prepare_matches_reference_golden(golden_dir, fixture):
    """fo_rank_prepare == the reference's voxel_pooling_prepare_v2 (view_transformer.py:223-281) run on
    CPU (fixtures), all five arrays exactly; tie order = ascending point index (the device sort's)."""
    from oracle import rank_oracle as ro
    g = np.load(os.path.join(golden_dir, fixture))
    if 'grid_lower_bound' in g:
        lb, itv, gs = g['grid_lower_bound'], g['grid_interval'], g['grid_size']
    else:
        from fusionocc_b200.rig import SHAPES
        lb, itv, gs = ro.create_grid_infos(**SHAPES['tiny'].grid_cfg())
    got, _ = _run_rank_prepare(g['coor'], lb, itv, gs)
    for name, a in zip(('ranks_bev', 'ranks_depth', 'ranks_feat', 'interval_starts', 'interval_lengths'), got):
        assert a.dtype == np.int32
        np.testing.assert_array_equal(a, g[name], err_msg=f'{fixture}:{name}')


def test_rank_prepare_fp32_hazard_is_exact_integers(golden_dir):
    """B=28 on the full grid: the reference's fp32 ranks are inexact.  The CUDA path
    must equal the int64-exact oracle, and must differ from the reference's fp32 result."""
    from oracle import rank_oracle as ro
    from fusionocc_b200.rig import SHAPES
    g = np.load(os.path.join(golden_dir, 'fp32_hazard_b28.npz'))
    lb, itv, gs = ro.create_grid_infos(**SHAPES['base'].grid_cfg())
    got, _ = _run_rank_prepare(g['coor'], lb, itv, gs)
    want = ro.voxel_pooling_prepare_v2(g['coor'], lb, itv, gs, 'int64')
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    assert not np.array_equal(got[0], g['ranks_bev']), 'expected the fp32 reference ranks to be inexact at B=28'


@pytest.mark.parametrize('case', ['base_B1', 'base_B2', 'base_B8', 'native_B1', 'native_B8', 'stress_B1', 'stress_B2'])
def test_rank_prepare_fullsize_digests(golden_dir, case):
    """Full BASELINE shapes: sha256 of all five arrays against digests of the reference's own output."""
    with open(os.path.join(golden_dir, 'fullsize_digests.json')) as f:
        dig = json.load(f)['digests'][case]
    name, B = case.split('_B')
    c = rig_case(name, int(B), with_ranks=False)
    assert sha(c['coor'].numpy()) == dig['coor'], 'rig/geometry drifted from the golden generation'
    got, _ = _run_rank_prepare(c['coor'].numpy(), c['lb'], c['itv'], c['gs'])
    assert got[0].shape[0] == dig['n_kept'] and got[3].shape[0] == dig['n_intervals']
    for nm, a in zip(('ranks_bev', 'ranks_depth', 'ranks_feat', 'interval_starts', 'interval_lengths'), got):
        assert sha(a) == dig[nm], f'{case}:{nm}'


def test_device_argsort_is_the_stable_order():
    """Pins the tie order: the reference's own torch ops (view_transformer.py:223-281) run ON THE GPU,
    with its default ``argsort()``, give exactly our arrays (device radix sort keeps ties ascending)."""
    from oracle.torch_cpu_path import voxel_pooling_prepare_v2_torch  # torch restatement, run on CUDA here
    c = rig_case('base', 2)
    coor = c['coor'].to(dev())
    lb, itv, gs = (torch.from_numpy(x) for x in (c['lb'], c['itv'], c['gs']))
    B, N, D, H, W, _ = coor.shape
    # the reference's default (non-stable API) argsort on device
    num_points = B * N * D * H * W
    idx = ((coor - lb.to(coor)) / itv.to(coor)).long().view(num_points, 3)
    batch_idx = torch.arange(0, B, dtype=torch.float32, device=coor.device).reshape(B, 1) \
        .expand(B, num_points // B).reshape(num_points, 1)
    cc = torch.cat((idx, batch_idx), 1)
    g = gs.to(coor.device)
    kept = (cc[:, 0] >= 0) & (cc[:, 0] < g[0]) & (cc[:, 1] >= 0) & (cc[:, 1] < g[1]) & (cc[:, 2] >= 0) & (cc[:, 2] < g[2])
    rd = torch.arange(num_points, dtype=torch.int, device=coor.device)[kept]
    cc = cc[kept]
    rbev = cc[:, 3] * (g[2] * g[1] * g[0]) + cc[:, 2] * (g[1] * g[0]) + cc[:, 1] * g[0] + cc[:, 0]
    order = rbev.argsort()                       # view_transformer.py:266
    ref_rd = rd[order].cpu().numpy()
    ref_rb = rbev[order].int().cpu().numpy()
    got, _ = _run_rank_prepare(c['coor'].numpy(), c['lb'], c['itv'], c['gs'])
    np.testing.assert_array_equal(got[0], ref_rb)
    np.testing.assert_array_equal(got[1], ref_rd)
    stable = voxel_pooling_prepare_v2_torch(coor, lb, itv, gs)
    for a, b in zip(got, stable):
        np.testing.assert_array_equal(a, b.cpu().numpy())


def test_rank_prepare_degenerate_long_intervals():
    """Every point of a camera in ONE voxel (intervals of 6 000 and 40 points: the global-memory and
    shared-memory in-segment sorts) plus ordinary short ones; equals the stable-sort oracle."""
    from oracle import rank_oracle as ro
    from fusionocc_b200.rig import SHAPES
    lb, itv, gs = ro.create_grid_infos(**SHAPES['base'].grid_cfg())
    rng = np.random.default_rng(3)
    coor = np.empty((2, 3, 10, 20, 30, 3), dtype=np.float32)
    coor[:, 0] = np.array([1.0, 2.0, 0.5], dtype=np.float32)                     # 6 000 points, one voxel
    coor[:, 1] = (rng.random((2, 10, 20, 30, 3)) * [80, 80, 6.4] - [40, 40, 1]).astype(np.float32)
    coor[:, 2] = np.array([-3.0, 7.0, 2.5], dtype=np.float32)
    coor[:, 2, :, :, 2:] = 1e5                                                   # 40*... partly outside
    coor[:, 2, 0, :2, :20] = np.array([-3.0, 7.0, 2.5], dtype=np.float32)        # 40 points, one voxel
    got, _ = _run_rank_prepare(coor, lb, itv, gs)
    want = ro.voxel_pooling_prepare_v2(coor, lb, itv, gs, 'int64')
    assert want[4].max() >= 6000
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


def test_rank_prepare_all_filtered_returns_nones():
    from fusionocc_b200 import LSSViewTransformer
    from fusionocc_b200.rig import SHAPES
    sh = SHAPES['tiny']
    vt = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels)
    coor = torch.full((1, 1, 2, 2, 2, 3), 1e6, device=dev())
    assert vt.voxel_pooling_prepare_v2(coor) == (None, None, None, None, None)


@pytest.mark.parametrize('name,B', [('tiny', 2), ('small', 2), ('base', 2)])
def test_chunk_sort_rank_pipeline_matches_the_default(name, B, monkeypatch):
    """FO_RANK_IMPL=1 (two-level sort, 1024-voxel chunks ordered in shared memory, csrc/rank_chunk.cuh) produces the
    same five rank arrays as the default global bucket sort, hence the reference's (both are checked against the
    oracle here), from coor and from the calibration."""
    from fusionocc_b200 import pack_calibration, rank_prepare, rank_prepare_calib
    case = rig_case(name, B)
    want = case['ranks']
    lb, itv, gs = case['lb'].tolist(), case['itv'].tolist(), [int(v) for v in case['gs']]
    coor = case['coor'].to(dev())
    cal = [c.to(dev()) for c in case['calib']]
    cam, bda12, has_t = pack_calibration(cal[0], cal[2], cal[3], cal[4], cal[5])
    for impl in ('1', '0'):
        monkeypatch.setenv('FO_RANK_IMPL', impl)
        for mode in ('coor', 'calib'):
            if mode == 'coor':
                rb, rd, rf, st, ln, counts, plan = rank_prepare(coor, lb, itv, gs)
            else:
                rb, rd, rf, st, ln, counts, plan = rank_prepare_calib(torch.as_tensor(case['frustum']).to(dev()), cam, bda12, has_t, B,
                                                                      coor.shape[1], lb, itv, gs)
            nk, ni = (int(v) for v in counts[:2].tolist())
            assert (nk, ni) == (len(want[0]), len(want[3])), f'impl {impl} {mode}: counts'
            for nm, a, b in zip(('ranks_bev', 'ranks_depth', 'ranks_feat', 'interval_starts', 'interval_lengths'),
                                (rb[:nk], rd[:nk], rf[:nk], st[:ni], ln[:ni]), want):
                assert np.array_equal(a.cpu().numpy(), b), f'impl {impl} {mode}: {nm} differs from the oracle'


# ------------------------------------------------------------------------------------------ forward/backward
def _values(case, C, seed=0):
    g = torch.Generator().manual_seed(seed)
    B, N, D, H, W, _ = case['coor'].shape
    depth = torch.randn(B, N, D, H, W, generator=g).softmax(dim=2)
    feat = torch.randn(B, N, H, W, C, generator=g)
    return depth, feat


@pytest.mark.parametrize('C', [32, 8, 20, 64, 80, 128, 3, 33, 132])
def test_forward_backward_bit_exact_vs_oracle_tiny(C):
    """Forward, depth_grad and feat_grad are bit-identical to the C restatement of
    bev_pool_cuda.cu:21-48,67-121 (sequential FMA orders preserved) for vector (C%4==0, <=128) and
    scalar channel counts.  Coarse 50x50x4 grid: partial last tile, intervals up to hundreds of points."""
    from fusionocc_b200 import bev_pool_v2
    from oracle import kernels as ok
    case = rig_case('small', 2)
    rb, rd, rf, st, ln = case['ranks']
    depth, feat = _values(case, C)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (2, Z, Y, X, C)
    want = ok.bev_pool_v2(depth.numpy(), feat.numpy(), rd, rf, rb, shape, st, ln)
    d = depth.to(dev()).requires_grad_()
    f = feat.to(dev()).requires_grad_()
    out = bev_pool_v2(d, f, t(rd), t(rf), t(rb), shape, t(st), t(ln))
    assert_bit_equal(out, want, f'forward C={C}')
    gen = torch.Generator().manual_seed(2)
    og = torch.randn(out.shape, generator=gen)
    out.backward(og.to(dev()))
    dg, fg = ok.bev_pool_v2_backward(og.numpy(), depth.numpy(), feat.numpy(), rd, rf, rb)
    assert_bit_equal(d.grad, dg, f'depth_grad C={C}')
    assert_bit_equal(f.grad, fg, f'feat_grad C={C}')


def test_forward_backward_base_shape_bit_exact():
    """BASELINE headline shape, B=1 (6 cams, D=88, C=32, 200x200x16): bit-exact forward and grads."""
    from fusionocc_b200 import bev_pool_v2
    from fusionocc_b200.rig import make_out_grad, make_values
    from oracle import kernels as ok
    case = rig_case('base', 1)
    sh = case['shape']
    rb, rd, rf, st, ln = case['ranks']
    depth, feat_nchw = make_values(sh, 1)
    feat = feat_nchw.permute(0, 1, 3, 4, 2)          # the non-contiguous NHWC view callers pass (:211)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (1, Z, Y, X, sh.channels)
    want = ok.bev_pool_v2(depth.numpy(), feat.contiguous().numpy(), rd, rf, rb, shape, st, ln)
    d = depth.to(dev()).requires_grad_()
    fn = feat_nchw.to(dev()).requires_grad_()
    out = bev_pool_v2(d, fn.permute(0, 1, 3, 4, 2), t(rd), t(rf), t(rb), shape, t(st), t(ln))
    assert_bit_equal(out, want, 'forward base')
    og = make_out_grad(1, sh.channels, Z, Y, X)
    out.backward(og.to(dev()))
    dg, fg = ok.bev_pool_v2_backward(og.numpy(), depth.numpy(), feat.contiguous().numpy(), rd, rf, rb)
    assert_bit_equal(d.grad, dg, 'depth_grad base')
    assert_bit_equal(fn.grad.permute(0, 1, 3, 4, 2).contiguous(), fg, 'feat_grad base')


def test_unsorted_and_duplicate_free_user_intervals_take_scatter_path():
    """The op must accept arbitrary user-supplied intervals (TRT path / KAT).  Intervals given in
    reverse voxel order are detected by the plan (flag) and produce the same dense result."""
    from fusionocc_b200 import bev_pool_v2, build_plan
    from oracle import kernels as ok
    case = rig_case('tiny', 1)
    rb, rd, rf, st, ln = case['ranks']
    C = 16
    depth, feat = _values(case, C)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (1, Z, Y, X, C)
    st_r, ln_r = st[::-1].copy(), ln[::-1].copy()
    want = ok.bev_pool_v2(depth.numpy(), feat.numpy(), rd, rf, rb, shape, st_r, ln_r)
    plan = build_plan(t(rb), t(st_r), t(ln_r), 1, X * Y * Z)
    assert plan.flags() & 1, 'reverse-ordered intervals must raise the unsorted flag'
    out = bev_pool_v2(depth.to(dev()), feat.to(dev()), t(rd), t(rf), t(rb), shape, t(st_r), t(ln_r))
    assert_bit_equal(out, want, 'forward (scatter path)')
    plan2 = build_plan(t(rb), t(st), t(ln), 1, X * Y * Z)
    assert plan2.flags() == 0


@pytest.mark.parametrize('order', ['reversed', 'shuffled'])
def test_backward_with_non_canonical_user_intervals(order):
    """the backward must not index the sub-tile tables of a plan the device flagged as
    non-canonical.  The reference backward ignores the forward intervals altogether (bev_pool.py:44-83), so the
    gradients for reversed / shuffled interval lists equal the ones for the canonical list."""
    from fusionocc_b200 import bev_pool_v2
    from oracle import kernels as ok
    case = rig_case('tiny', 2)
    rb, rd, rf, st, ln = case['ranks']
    C = 32
    depth, feat = _values(case, C)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (2, Z, Y, X, C)
    if order == 'reversed':
        perm = np.arange(len(st))[::-1].copy()
    else:
        perm = np.random.default_rng(3).permutation(len(st))
    st_p, ln_p = st[perm].copy(), ln[perm].copy()
    want = ok.bev_pool_v2(depth.numpy(), feat.numpy(), rd, rf, rb, shape, st_p, ln_p)
    d = depth.to(dev()).requires_grad_()
    f = feat.to(dev()).requires_grad_()
    out = bev_pool_v2(d, f, t(rd), t(rf), t(rb), shape, t(st_p), t(ln_p))
    assert_bit_equal(out, want, f'forward ({order} intervals)')
    og = torch.randn(out.shape, generator=torch.Generator().manual_seed(11))
    out.backward(og.to(dev()))
    dg, fg = ok.bev_pool_v2_backward(og.numpy(), depth.numpy(), feat.numpy(), rd, rf, rb)
    assert_bit_equal(d.grad, dg, f'depth_grad ({order} intervals)')
    assert_bit_equal(f.grad, fg, f'feat_grad ({order} intervals)')


def test_backward_plan_follows_ranks_depth():
    """the cached backward plan bakes ranks_depth in.  The same (ranks_bev, starts, lengths,
    ranks_feat) objects with ranks_depth edited IN PLACE, or with another ranks_depth tensor, must not reuse it."""
    from fusionocc_b200 import bev_pool_v2
    from oracle import kernels as ok
    case = rig_case('tiny', 1)
    rb, rd, rf, st, ln = case['ranks']
    C = 8
    depth, feat = _values(case, C)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (1, Z, Y, X, C)
    trb, trf, tst, tln, trd = t(rb), t(rf), t(st), t(ln), t(rd)
    og = torch.randn((1, C, Z, Y, X), generator=torch.Generator().manual_seed(4))
    # a second, different but valid (collision-free) depth index per point: the same pixel, depth bins mirrored
    HW = case['coor'].shape[3] * case['coor'].shape[4]
    D = case['coor'].shape[2]
    dbin = (rd // HW) % D
    rd2 = (rd + (D - 1 - 2 * dbin) * HW).astype(rd.dtype)

    def run(rd_tensor, rd_np):
        d = depth.to(dev()).requires_grad_()
        f = feat.to(dev()).requires_grad_()
        out = bev_pool_v2(d, f, rd_tensor, trf, trb, shape, tst, tln)
        out.backward(og.to(dev()))
        dg, fg = ok.bev_pool_v2_backward(og.numpy(), depth.numpy(), feat.numpy(), rd_np, rf, rb)
        assert_bit_equal(d.grad, dg, 'depth_grad')
        assert_bit_equal(f.grad, fg, 'feat_grad')

    run(trd, rd)
    run(t(rd2), rd2)                         # another tensor object
    trd.copy_(t(rd2))                        # the first object, edited in place
    run(trd, rd2)


def test_channels_last_out_grad_layout():
    """A (B,Z,Y,X,C)-contiguous upstream gradient (channels-last consumer) is read in place."""
    from fusionocc_b200.bev_pool import QuickCumsumCuda
    from oracle import kernels as ok
    case = rig_case('tiny', 2)
    rb, rd, rf, st, ln = case['ranks']
    C = 32
    depth, feat = _values(case, C)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (2, Z, Y, X, C)
    d = depth.to(dev()).requires_grad_()
    f = feat.to(dev()).requires_grad_()
    out5 = QuickCumsumCuda.apply(d, f, t(rd), t(rf), t(rb), shape, t(st), t(ln))
    assert tuple(out5.shape) == shape
    gen = torch.Generator().manual_seed(5)
    og_bzyxc = torch.randn(shape, generator=gen)
    out5.backward(og_bzyxc.to(dev()))                 # contiguous in (B,Z,Y,X,C)
    dg, fg = ok.bev_pool_v2_backward(og_bzyxc.permute(0, 4, 1, 2, 3).contiguous().numpy(), depth.numpy(),
                                     feat.numpy(), rd, rf, rb)
    assert_bit_equal(d.grad, dg, 'depth_grad (BZYXC og)')
    assert_bit_equal(f.grad, fg, 'feat_grad (BZYXC og)')


def test_half_inputs_are_computed_in_fp32():
    """bev_pool.py:20-21: any float dtype in, fp32 out; grads come back in the input dtype."""
    from fusionocc_b200 import bev_pool_v2
    from oracle import kernels as ok
    case = rig_case('tiny', 1)
    rb, rd, rf, st, ln = case['ranks']
    C = 8
    depth, feat = _values(case, C)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (1, Z, Y, X, C)
    dh, fh = depth.half(), feat.half()
    want = ok.bev_pool_v2(dh.float().numpy(), fh.float().numpy(), rd, rf, rb, shape, st, ln)
    d = dh.to(dev()).requires_grad_()
    f = fh.to(dev()).requires_grad_()
    out = bev_pool_v2(d, f, t(rd), t(rf), t(rb), shape, t(st), t(ln))
    assert out.dtype == torch.float32
    assert_bit_equal(out, want, 'forward (fp16 inputs)')
    out.sum().backward()
    assert d.grad.dtype == torch.float16 and f.grad.dtype == torch.float16


def test_trt_bev_pool_v2():
    """TRTBEVPoolv2 eager forward (bev_pool.py:121-142): Z=1, returns (B,Y,X,C)."""
    from fusionocc_b200 import TRTBEVPoolv2
    from oracle import kernels as ok
    rng = np.random.default_rng(0)
    N, D, H, W, C, OH, OW = 2, 3, 4, 5, 8, 16, 16
    depth = rng.random((N, D, H, W), dtype=np.float32)
    feat = rng.standard_normal((N, H, W, C)).astype(np.float32)
    P = N * D * H * W
    keep = rng.random(P) < 0.7
    rd = np.nonzero(keep)[0].astype(np.int32)
    rf = ((rd // (D * H * W)) * (H * W) + rd % (H * W)).astype(np.int32)
    rb = rng.integers(0, OH * OW, size=rd.shape[0]).astype(np.int32)
    order = np.argsort(rb, kind='stable')
    rb, rd, rf = rb[order], rd[order], rf[order]
    from oracle.rank_oracle import intervals_from_sorted
    st, ln = intervals_from_sorted(rb)
    out = TRTBEVPoolv2.apply(t(depth), t(feat), t(rd), t(rf), t(rb), t(st), t(ln), OH, OW)
    want = ok.bev_pool_v2(depth[None], feat[None], rd, rf, rb, (1, 1, OH, OW, C), st, ln)   # (1,C,1,OH,OW)
    assert tuple(out.shape) == (1, OH, OW, C)
    assert_bit_equal(out.permute(0, 3, 1, 2).contiguous(), want[:, :, 0], 'TRTBEVPoolv2')


@pytest.mark.parametrize('OZ', [1, 3])
def test_trt_bev_pool_v2_licrocc_variant(OZ):
    """LiCROcc's vendored TRTBEVPoolv2 (projects/LiCROcc/.../bev_pool_v2/bev_pool.py:108-159): output_z argument,
    squeeze + NHWC permute only for output_z == 1, (B,C,Z,Y,X) otherwise.  Imported through the overlay path."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'overlay', 'projects', 'LiCROcc',
                        'projects', 'mmdet3d_plugin', 'ops', 'bev_pool_v2', 'bev_pool.py')
    spec = importlib.util.spec_from_file_location('licrocc_bev_pool_gpu', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from oracle import kernels as ok
    from oracle.rank_oracle import intervals_from_sorted
    rng = np.random.default_rng(1)
    N, D, H, W, C, OH, OW = 2, 3, 4, 5, 8, 12, 16
    depth = rng.random((N, D, H, W), dtype=np.float32)
    feat = rng.standard_normal((N, H, W, C)).astype(np.float32)
    P = N * D * H * W
    rd = np.nonzero(rng.random(P) < 0.7)[0].astype(np.int32)
    rf = ((rd // (D * H * W)) * (H * W) + rd % (H * W)).astype(np.int32)
    rb = rng.integers(0, OZ * OH * OW, size=rd.shape[0]).astype(np.int32)
    order = np.argsort(rb, kind='stable')
    rb, rd, rf = rb[order], rd[order], rf[order]
    st, ln = intervals_from_sorted(rb)
    out = mod.TRTBEVPoolv2.apply(t(depth), t(feat), t(rd), t(rf), t(rb), t(st), t(ln), OH, OW, OZ)
    want = ok.bev_pool_v2(depth[None], feat[None], rd, rf, rb, (1, OZ, OH, OW, C), st, ln)   # (1,C,OZ,OH,OW)
    if OZ == 1:
        assert tuple(out.shape) == (1, OH, OW, C)
        assert_bit_equal(out.permute(0, 3, 1, 2).contiguous(), want[:, :, 0], 'LiCROcc TRTBEVPoolv2 (Z=1)')
    else:
        assert tuple(out.shape) == (1, C, OZ, OH, OW)
        assert_bit_equal(out, want, 'LiCROcc TRTBEVPoolv2 (Z=3)')


# ------------------------------------------------------------------------------------------ module level
@pytest.mark.parametrize('sync_free', [False, True])
@pytest.mark.parametrize('collapse_z', [False, True])
def test_view_transformer_end_to_end_vs_torch_cpu_path(sync_free, collapse_z):
    """LSSViewTransformer.view_transform (geometry -> ranks -> splat) against the reference-style
    pure-PyTorch CPU path; tolerance rtol=atol=1e-5 (torch's index_add_ rounds the product first)."""
    from fusionocc_b200 import LSSViewTransformer
    from fusionocc_b200.rig import SHAPES, make_calibration, make_values
    from oracle.torch_cpu_path import view_transform_step_cpu
    sh = SHAPES['tiny']
    B = 2
    vt = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels,
                            collapse_z=collapse_z, sync_free=sync_free).to(dev())
    cal = make_calibration(sh, B)
    depth, feat = make_values(sh, B)
    H, W = sh.feat_hw
    x = torch.zeros(B, sh.n_cams, 8, H, W)
    inp = [x.to(dev())] + [c.to(dev()) for c in cal]
    d = depth.to(dev()).requires_grad_()
    f = feat.to(dev()).requires_grad_()
    bev, _ = vt.view_transform(inp, d.view(B * sh.n_cams, vt.D, H, W), f.view(B * sh.n_cams, sh.channels, H, W))
    X, Y, Z = vt._grid_xyz()
    og = torch.randn(B, sh.channels, Z, Y, X, generator=torch.Generator().manual_seed(3))
    ref_out, ref_dg, ref_fg, nk, ni = view_transform_step_cpu(vt.frustum, cal, depth, feat,
                                                              (vt.grid_lower_bound, vt.grid_interval, vt.grid_size), og)
    want = torch.cat(ref_out.unbind(dim=2), 1) if collapse_z else ref_out
    assert bev.shape == want.shape
    torch.testing.assert_close(bev.cpu(), want, rtol=RTOL, atol=ATOL)
    g = torch.cat(og.unbind(dim=2), 1) if collapse_z else og
    bev.backward(g.to(dev()))
    torch.testing.assert_close(d.grad.cpu(), ref_dg, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(f.grad.cpu(), ref_fg, rtol=RTOL, atol=ATOL)


def test_accelerate_mode_caches_ranks_and_plan():
    from fusionocc_b200 import LSSViewTransformer
    from fusionocc_b200.rig import SHAPES, make_calibration, make_values
    sh = SHAPES['tiny']
    B = 1
    vt = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels,
                            accelerate=True, collapse_z=True).to(dev())
    vt2 = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels,
                             accelerate=False, collapse_z=True).to(dev())
    cal = make_calibration(sh, B)
    depth, feat = make_values(sh, B)
    H, W = sh.feat_hw
    inp = [torch.zeros(B, sh.n_cams, 8, H, W, device=dev())] + [c.to(dev()) for c in cal]
    dd, ff = depth.to(dev()).view(-1, vt.D, H, W), feat.to(dev()).view(-1, sh.channels, H, W)
    a1, _ = vt.view_transform(inp, dd, ff)
    assert vt.initial_flag is False and vt.ranks_bev.dtype == torch.int32
    plan1 = vt.ranks_bev._fo_plan[0]                  # attached by the rank precompute, lives with the tensor
    a2, _ = vt.view_transform(inp, dd, ff)
    assert vt.ranks_bev._fo_plan[0] is plan1, 'second accelerate call must reuse the attached plan'
    b1, _ = vt2.view_transform(inp, dd, ff)
    # accelerate path squeezes Z (:305) — with Z>1 the shapes differ from the collapse path by design
    assert torch.equal(a1, a2)
    assert torch.equal(torch.cat(a1.unbind(dim=2), 1), b1)


# ------------------------------------------------------------------------------------------ raw C ABI
def test_compat_l0_symbols_match_reference_launcher_semantics():
    """fo_compat_bev_pool_v2 / _grad: caller-zeroed (B,Z,Y,X,C) out, backward arrays pre-sorted by
    ranks_feat — the contract of bev_pool.cpp:7-14."""
    import ctypes
    from fusionocc_b200 import _cabi
    from oracle import kernels as ok, rank_oracle as ro
    lib = _cabi.load()
    case = rig_case('tiny', 1)
    rb, rd, rf, st, ln = case['ranks']
    C = 32
    depth, feat = _values(case, C)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (1, Z, Y, X, C)
    out = torch.zeros(shape, device=dev())
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    dd, ff = depth.to(dev()), feat.to(dev())
    trb, trd, trf, tst, tln = t(rb), t(rd), t(rf), t(st), t(ln)
    torch.cuda.synchronize()
    lib.fo_compat_bev_pool_v2(C, len(ln), p(dd), p(ff), p(trd), p(trf), p(trb), p(tst), p(tln), p(out))
    torch.cuda.synchronize()
    want = ok.bev_pool_v2_forward_bzyxc(depth.numpy(), feat.numpy(), rd, rf, rb, shape, st, ln)
    assert_bit_equal(out, want, 'compat forward')
    og = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    brb, brd, brf, bst, bln = ro.backward_resort(rb, rd, rf)
    dg = torch.zeros_like(dd)
    fg = torch.zeros_like(ff)
    tog = og.to(dev())
    a = [t(np.ascontiguousarray(x)) for x in (brd, brf, brb, bst, bln)]
    torch.cuda.synchronize()
    lib.fo_compat_bev_pool_v2_grad(C, len(bln), p(tog), p(dd), p(ff), p(a[0]), p(a[1]), p(a[2]), p(a[3]), p(a[4]),
                                   p(dg), p(fg))
    torch.cuda.synchronize()
    wdg, wfg = ok.bev_pool_v2_backward(og.permute(0, 4, 1, 2, 3).contiguous().numpy(), depth.numpy(), feat.numpy(),
                                       rd, rf, rb)
    assert_bit_equal(dg, wdg, 'compat depth_grad')
    assert_bit_equal(fg, wfg, 'compat feat_grad')


@pytest.mark.parametrize('two_streams', [False, True])
def test_host_buffer_entry_point(two_streams):
    """fo_view_transform_host: pinned host buffers in, host buffers out (the e2e leg of bench.py); with a
    second stream the out_grad upload overlaps the forward and the voxel download."""
    import ctypes
    from fusionocc_b200 import _cabi
    from fusionocc_b200.rig import make_out_grad
    from oracle import kernels as ok
    lib = _cabi.load()
    case = rig_case('tiny', 2)
    rb, rd, rf, st, ln = case['ranks']
    C = 8
    depth, feat = _values(case, C)
    B, N, D, H, W, _ = case['coor'].shape
    X, Y, Z = (int(v) for v in case['gs'])
    og = make_out_grad(B, C, Z, Y, X)
    pin = lambda x: x.contiguous().pin_memory()
    h_coor, h_depth, h_feat, h_og = pin(case['coor']), pin(depth), pin(feat), pin(og)
    h_out = torch.empty(B, C, Z, Y, X).pin_memory()
    h_dg = torch.empty_like(depth).pin_memory()
    h_fg = torch.empty_like(feat).pin_memory()
    h_counts = torch.zeros(4, dtype=torch.int32).pin_memory()
    ws_bytes = lib.fo_view_transform_host_workspace_bytes(B, N, D, H, W, C, X, Y, Z, 1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev())
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    s = torch.cuda.current_stream()
    up = torch.cuda.Stream() if two_streams else None
    rc = lib.fo_view_transform_host(ctypes.c_void_p(s.cuda_stream), p(h_coor), p(h_depth), p(h_feat), p(h_og),
                                    B, N, D, H, W, C, _cabi.f3(case['lb']), _cabi.f3(case['itv']), X, Y, Z,
                                    p(h_out), p(h_dg), p(h_fg), p(h_counts), p(ws), ws_bytes,
                                    ctypes.c_void_p(up.cuda_stream) if up is not None else None)
    _cabi.check(rc, 'fo_view_transform_host')
    s.synchronize()
    assert h_counts[0].item() == len(rb) and h_counts[1].item() == len(st)
    shape = (B, Z, Y, X, C)
    want = ok.bev_pool_v2(depth.numpy(), feat.numpy(), rd, rf, rb, shape, st, ln)
    assert_bit_equal(h_out, want, 'host entry forward')
    dg, fg = ok.bev_pool_v2_backward(og.numpy(), depth.numpy(), feat.numpy(), rd, rf, rb)
    assert_bit_equal(h_dg, dg, 'host entry depth_grad')
    assert_bit_equal(h_fg, fg, 'host entry feat_grad')


def test_invalid_arguments_raise():
    from fusionocc_b200 import _cabi, bev_pool_v2
    z = lambda *s: torch.zeros(*s, device=dev())
    zi = lambda n: torch.zeros(n, dtype=torch.int32, device=dev())
    with pytest.raises(ValueError):
        bev_pool_v2(z(1, 1, 2, 2, 2), z(1, 1, 2, 2, 4), zi(1), zi(1), zi(1), (1, 1, 2, 2, 8), zi(1), zi(1))
    with pytest.raises(RuntimeError):
        bev_pool_v2(torch.zeros(1, 1, 2, 2, 2), z(1, 1, 2, 2, 4), zi(1), zi(1), zi(1), (1, 1, 2, 2, 4), zi(1), zi(1))
    lib = _cabi.load()
    rc = lib.fo_bev_pool_v2_forward(None, 0, None, None, None, None, None, None, None, 0, 0, None, 1, 1, None, 0, 0,
                                    None, 0)
    assert rc == 1 and b'channels' in lib.fo_last_error()
    # the entry points added for rows f-1 .. f-4 validate their arguments the same way
    rc = lib.fo_bev_pool_v2_forward_slice(None, 8, None, None, None, None, None, None, None, 0, 0, None, 1, 64, None,
                                          0, 12, 6, 0, None, 0)
    assert rc == 1 and b'slice' in lib.fo_last_error()
    rc = lib.fo_bev_pool_v2_backward_slice(None, 8, None, 0, 4, 0, None, None, 0, 0, 1, 64, 0, 1, None, None, None, 0,
                                           None, 0, None, 0)
    assert rc == 1 and b'slice' in lib.fo_last_error()
    f3 = _cabi.f3([0, 0, 0])
    rc = lib.fo_rank_prepare_calib(None, None, None, None, 0, 3, None, 1, 1, 1, 1, 1, f3, f3, 1, 1, 1, None, None, None,
                                   None, None, None, None, 0, None, 0)
    assert rc == 1 and b'calibration' in lib.fo_last_error()
    fr = z(3)
    rc = lib.fo_rank_prepare_calib(None, fr.data_ptr(), fr.data_ptr(), fr.data_ptr(), 0, 7, None, 1, 1, 1, 1, 1, f3, f3,
                                   1, 1, 1, None, None, None, None, None, None, None, 0, None, 0)
    assert rc == 1 and b'matvec_mode' in lib.fo_last_error()
    rc = lib.fo_rank_from_keys(None, None, 5, 0, None, None, None, None, None, None, 0)
    assert rc == 1
    rc = lib.fo_lift_prepare_forward(None, fr.data_ptr(), 9, 1, 4, 2, 2, 4, None, None)
    assert rc == 1 and b'dtype' in lib.fo_last_error()
    rc = lib.fo_lift_prepare_forward(None, fr.data_ptr(), 0, 1, 3, 2, 2, 4, None, None)
    assert rc == 1 and b'sizes' in lib.fo_last_error()


def test_size_independent_properties_at_full_size():
    """BASELINE full size (base, B=2): linearity in feat, checksum of checksums, and empty voxels are
    exactly +0.0f (bit pattern 0)."""
    from fusionocc_b200 import bev_pool_v2
    from fusionocc_b200.rig import make_values
    case = rig_case('base', 2)
    sh = case['shape']
    rb, rd, rf, st, ln = (t(a) for a in case['ranks'])
    depth, feat_nchw = make_values(sh, 2)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (2, Z, Y, X, sh.channels)
    d = depth.to(dev())
    f = feat_nchw.to(dev()).permute(0, 1, 3, 4, 2)
    o1 = bev_pool_v2(d, f, rd, rf, rb, shape, st, ln)
    o2 = bev_pool_v2(d, 2.0 * f, rd, rf, rb, shape, st, ln)
    assert torch.equal(o2, 2.0 * o1), 'scaling feat by 2 must scale the output exactly (power of two)'
    # total mass: sum_out == sum_i depth[p_i] * sum_c feat[q_i, c]   (fp64 accumulate, 1e-5 relative)
    mass = (d.reshape(-1)[rd.long()].double() * f.reshape(-1, sh.channels)[rf.long()].double().sum(1)).sum()
    assert abs(o1.double().sum().item() - mass.item()) <= 1e-5 * abs(mass.item()) + 1e-3
    occupied = torch.zeros(2 * Z * Y * X, dtype=torch.bool, device=dev())
    occupied[rb.long()] = True
    empty = ~occupied.view(2, 1, Z, Y, X).expand(-1, sh.channels, -1, -1, -1)
    assert (o1[empty].view(torch.int32) == 0).all(), 'untouched voxels must be +0.0f'
    # idempotence: same inputs, same bits
    assert torch.equal(o1, bev_pool_v2(d, f, rd, rf, rb, shape, st, ln))


def test_batch32_shard_matches_single_sample_runs():
    """BASELINE configs[2]: batch 64 over 2 GPUs = 32 samples per rank.  B*V = 20.5 M exceeds 2^24 — where the
    reference's fp32 ranks merge voxels and where the packed-key backward plan hands over to the
    64-bit variant.  Every sample of the batched run must equal, bit for bit, the same sample run alone (samples
    never share a voxel), forward and both gradients."""
    from fusionocc_b200 import LSSViewTransformer
    from fusionocc_b200.rig import SHAPES, make_calibration, make_values
    sh = SHAPES['base']
    B = 32
    vt = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels,
                            collapse_z=False, sync_free=True)
    cal = [c.to(dev()) for c in make_calibration(sh, B)]
    depth1, feat1 = make_values(sh, 1)                      # the same values for every sample keep the host side small
    depth = depth1.expand(B, -1, -1, -1, -1).contiguous().to(dev()).requires_grad_()
    feat = feat1.expand(B, -1, -1, -1, -1).contiguous().to(dev()).requires_grad_()
    out = vt.voxel_pooling_v2(vt.get_lidar_coor(*cal), depth, feat)
    X, Y, Z = vt._grid_xyz()
    assert out.shape == (B, sh.channels, Z, Y, X)
    g1 = torch.randn(1, sh.channels, Z, Y, X, generator=torch.Generator().manual_seed(4)).to(dev())
    out.backward(g1.expand(B, -1, -1, -1, -1))
    for b in (0, 13, 31):
        d = depth1.to(dev()).requires_grad_()
        f = feat1.to(dev()).requires_grad_()
        o = vt.voxel_pooling_v2(vt.get_lidar_coor(*[c[b:b + 1] for c in cal]), d, f)
        assert torch.equal(o[0].view(torch.int32), out[b].detach().view(torch.int32)), f'sample {b}: forward'
        o.backward(g1)
        assert torch.equal(d.grad[0].view(torch.int32), depth.grad[b].view(torch.int32)), f'sample {b}: depth_grad'
        assert torch.equal(f.grad[0].view(torch.int32), feat.grad[b].view(torch.int32)), f'sample {b}: feat_grad'


@pytest.mark.parametrize('grid,C,B,D', [((7, 5, 3), 5, 3, 7), ((1, 1, 1), 32, 2, 7), ((9, 2, 2), 36, 1, 7),
                                        ((33, 1, 1), 8, 2, 7), ((7, 5, 3), 12, 2, 120), ((1, 1, 1), 32, 1, 60)])
def test_odd_grids_take_the_scalar_paths(grid, C, B, D):
    """Voxel counts that are not a multiple of 4 / 32 (ragged last sub-tile, no 128-bit stores or loads), one-voxel
    grids and channel counts that are not a multiple of 4: ranks exact, forward and both gradients bit-exact vs the
    oracle, for the gradient of bev_pool_v2's output and for a channels-last gradient; the dense cases put more than
    256 points into a sub-tile, so the forward's front CTAs write ragged, unaligned voxel columns."""
    from fusionocc_b200.bev_pool import bev_pool_v2_with_plan
    from fusionocc_b200.view_transformer import rank_prepare
    from oracle import kernels as ok, rank_oracle as ro
    X, Y, Z = grid
    N, H, W = 2, 3, 5                            # D = 120 / 60: > 256 points per sub-tile -> the front CTAs, ragged
    g = torch.Generator().manual_seed(X * 100 + Y * 10 + Z)
    lb = np.array([-1.0, -2.0, 0.5], np.float32)
    itv = np.array([0.5, 0.25, 1.0], np.float32)
    span = torch.tensor([X * 0.5, Y * 0.25, Z * 1.0])
    coor = (torch.rand(B, N, D, H, W, 3, generator=g) * 1.4 - 0.2) * span + torch.from_numpy(lb)   # ~30 % outside
    want = ro.voxel_pooling_prepare_v2(coor.numpy(), lb, itv, np.array([X, Y, Z], np.float32), 'int64')
    rb, rd, rf, st, ln, counts, plan = rank_prepare(coor.to(dev()), lb.tolist(), itv.tolist(), (X, Y, Z))
    nk, ni = (int(v) for v in counts[:2].tolist())
    assert nk == len(want[0]) and ni == len(want[3])
    for a, b, nm in zip((rb[:nk], rd[:nk], rf[:nk], st[:ni], ln[:ni]), want,
                        ('ranks_bev', 'ranks_depth', 'ranks_feat', 'interval_starts', 'interval_lengths')):
        assert np.array_equal(a.cpu().numpy(), b), nm
    depth = torch.rand(B, N, D, H, W, generator=g)
    feat = torch.randn(B, N, H, W, C, generator=g)
    shape = (B, Z, Y, X, C)
    w_out = ok.bev_pool_v2(depth.numpy(), feat.numpy(), want[1], want[2], want[0], shape, want[3], want[4])
    og = torch.randn(B, C, Z, Y, X, generator=g)
    w_dg, w_fg = ok.bev_pool_v2_backward(og.numpy(), depth.numpy(), feat.numpy(), want[1], want[2], want[0])
    for channels_last in (False, True):
        d = depth.to(dev()).requires_grad_()
        f = feat.to(dev()).requires_grad_()
        out = bev_pool_v2_with_plan(d, f, rd, rf, rb, shape, st, ln, plan)
        assert_bit_equal(out, w_out, f'forward {grid}')
        gr = og.to(dev())
        if channels_last:
            gr = gr.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
        out.backward(gr)
        assert_bit_equal(d.grad, w_dg, f'depth_grad {grid} cl={channels_last}')
        assert_bit_equal(f.grad, w_fg, f'feat_grad {grid} cl={channels_last}')


def test_empty_rank_arrays_give_zero_output_and_zero_gradients():
    """Op-level empty case (the reference would launch a 0-block grid): zero points / intervals
    must produce an all-zero voxel tensor and all-zero gradients, through the dense path and its plans."""
    from fusionocc_b200 import bev_pool_v2
    B, N, D, H, W, C, Z, Y, X = 2, 2, 3, 2, 3, 8, 2, 4, 8
    e = torch.empty(0, dtype=torch.int32, device=dev())
    d = torch.rand(B, N, D, H, W, device=dev()).requires_grad_()
    f = torch.randn(B, N, H, W, C, device=dev()).requires_grad_()
    out = bev_pool_v2(d, f, e, e, e, (B, Z, Y, X, C), e, e)
    assert out.shape == (B, C, Z, Y, X) and bool((out.view(torch.int32) == 0).all())
    out.backward(torch.ones_like(out))
    assert bool((d.grad == 0).all()) and bool((f.grad == 0).all())
