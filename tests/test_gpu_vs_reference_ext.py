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
