"""The new CUDA path against the REFERENCE CUDA EXTENSION ITSELF (oracle/_ref, compiled unmodified
from /root/reference/mmdet3d/ops/bev_pool_v2/src by oracle/build_ref.py) on identical inputs at the
BASELINE shapes.  Forward and both gradients must be bit-identical (same FFMA chains)."""
import numpy as np
import pytest
import torch

from tests.helpers import rig_case

pytestmark = pytest.mark.gpu


def _ref():
    from oracle import ref_ext
    if not ref_ext.available():
        pytest.skip('oracle/_ref not built (reference tree absent at build time)')
    return ref_ext


@pytest.mark.parametrize('name,B', [('base', 1), ('base', 8), ('native', 2), ('stress', 1), ('stress', 2)])
def test_forward_and_backward_bit_identical_to_reference_extension(name, B):
    ref = _ref()
    from fusionocc_b200 import bev_pool_v2
    from fusionocc_b200.rig import make_out_grad, make_values
    dev = torch.device('cuda:0')
    case = rig_case(name, B)
    sh = case['shape']
    rb, rd, rf, st, ln = (torch.from_numpy(a).to(dev) for a in case['ranks'])
    depth, feat_nchw = make_values(sh, B)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (B, Z, Y, X, sh.channels)
    d = depth.to(dev).requires_grad_()
    f = feat_nchw.to(dev).requires_grad_()
    feat_view = f.permute(0, 1, 3, 4, 2)
    out = bev_pool_v2(d, feat_view, rd, rf, rb, shape, st, ln)
    want = ref.forward(d.detach(), feat_view.detach(), rd, rf, rb, shape, st, ln)
    assert out.shape == want.shape and out.is_contiguous()
    assert torch.equal(out.view(torch.int32), want.view(torch.int32)), 'forward differs from the reference ext'
    og = make_out_grad(B, sh.channels, Z, Y, X).to(dev)
    out.backward(og)
    dg, fg = ref.backward(og, d.detach(), feat_view.detach(), rd, rf, rb)
    assert torch.equal(d.grad.view(torch.int32), dg.view(torch.int32)), 'depth_grad differs from the reference ext'
    got_fg = f.grad.permute(0, 1, 3, 4, 2).contiguous()
    assert torch.equal(got_fg.view(torch.int32), fg.view(torch.int32)), 'feat_grad differs from the reference ext'


def test_reference_argsort_on_device_orders_backward_like_our_plan():
    """bev_pool.py:47: ranks_feat.argsort() on device == stable order (our backward plan's order)."""
    _ref()
    dev = torch.device('cuda:0')
    case = rig_case('base', 2)
    rf = torch.from_numpy(case['ranks'][2]).to(dev)
    a = rf.argsort()
    b = rf.argsort(stable=True)
    assert torch.equal(a, b)


def test_whole_chain_at_the_headline_batch_vs_reference_precompute_and_extension():
    """VERDICT r1 (parity holes): nothing fed from the oracle.  At the headline batch (base shape, B = 8) the
    product's own chain — fused geometry + rank precompute -> bev_pool_v2 -> backward, through the module surface —
    against the REFERENCE's chain on the same GPU: its eager-torch geometry + rank precompute
    (view_transformer.py:135-173,223-281, restated op for op) feeding the UNMODIFIED reference CUDA extension."""
    ref = _ref()
    from fusionocc_b200 import LSSViewTransformer
    from fusionocc_b200.rig import SHAPES, make_calibration, make_out_grad, make_values
    from oracle.torch_cpu_path import voxel_pooling_prepare_v2_torch
    dev = torch.device('cuda:0')
    sh, B = SHAPES['base'], 8
    vt = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels,
                            collapse_z=False).to(dev)
    assert vt.fuse_geometry, 'the fused geometry is the default path'
    cal = [c.to(dev) for c in make_calibration(sh, B)]
    depth, feat_nchw = make_values(sh, B)
    H, W = sh.feat_hw
    X, Y, Z = vt._grid_xyz()
    d = depth.to(dev).view(-1, vt.D, H, W).requires_grad_()
    f = feat_nchw.to(dev).view(-1, sh.channels, H, W).requires_grad_()
    inp = [torch.zeros(B, sh.n_cams, 8, H, W, device=dev)] + cal
    out, _ = vt.view_transform(inp, d, f)
    og = make_out_grad(B, sh.channels, Z, Y, X).to(dev)
    out.backward(og)
    # the reference's chain
    coor = vt.get_lidar_coor(*cal)                       # the reference's torch ops (same library kernels)
    rb, rd, rf, st, ln = voxel_pooling_prepare_v2_torch(coor, vt.grid_lower_bound, vt.grid_interval, vt.grid_size)
    feat_view = feat_nchw.to(dev).permute(0, 1, 3, 4, 2)
    want = ref.forward(depth.to(dev), feat_view, rd, rf, rb, (B, Z, Y, X, sh.channels), st, ln)
    assert torch.equal(out.view(torch.int32), want.view(torch.int32)), 'voxels differ from the reference chain'
    dg, fg = ref.backward(og, depth.to(dev), feat_view, rd, rf, rb)
    assert torch.equal(d.grad.view(-1).view(torch.int32), dg.view(-1).view(torch.int32)), 'depth_grad differs'
    got_fg = f.grad.view(B, sh.n_cams, sh.channels, H, W).permute(0, 1, 3, 4, 2).contiguous()
    assert torch.equal(got_fg.view(torch.int32), fg.view(torch.int32)), 'feat_grad differs'


def test_quick_cumsum_output_is_a_view_with_the_reference_shape_and_values():
    """QuickCumsumCuda.apply returns a (B,Z,Y,X,C)-SHAPED tensor like the reference (bev_pool.py:27,41), backed by
    (B,C,Z,Y,X) memory (so the wrapper's permute().contiguous() is free).  Callers that index it, reshape() it or
    call .contiguous() get the reference's values; only .view() on it needs .contiguous() first — documented contract."""
    ref = _ref()
    from fusionocc_b200.bev_pool import QuickCumsumCuda
    dev = torch.device('cuda:0')
    case = rig_case('tiny', 2)
    rb, rd, rf, st, ln = (torch.from_numpy(a).to(dev) for a in case['ranks'])
    B, N, D, H, W, _ = case['coor'].shape
    g = torch.Generator().manual_seed(0)
    depth = torch.rand(B, N, D, H, W, generator=g).to(dev)
    feat = torch.randn(B, N, H, W, 8, generator=g).to(dev)
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (B, Z, Y, X, 8)
    x = QuickCumsumCuda.apply(depth, feat, rd, rf, rb, shape, st, ln)
    assert tuple(x.shape) == shape and not x.is_contiguous()
    from oracle.ref_ext import ext
    want = feat.new_zeros(shape)                          # what the reference's forward returns (contiguous B,Z,Y,X,C)
    ext().bev_pool_v2_forward(depth, feat, want, rd.int(), rf.int(), rb.int(), ln.int(), st.int())
    assert torch.equal(x.contiguous().view(torch.int32), want.view(torch.int32))
    assert torch.equal(x.reshape(-1, 8).view(torch.int32), want.view(-1, 8).view(torch.int32))
    assert torch.equal(x[1, 2, 3].view(torch.int32), want[1, 2, 3].view(torch.int32))
