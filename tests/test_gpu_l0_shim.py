"""L0 drop-in proof: the reference's UNMODIFIED pybind translation unit
mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp, compiled by oracle/build_ref.py --which shim together with the forwarding
file oracle/shim/bev_pool_shim.cpp and LINKED AGAINST libfusionocc_b200.so, runs the reference's own Python-side op
sequence (bev_pool.py:17-92, restated in oracle/ref_ext.py) on the new kernels: the reference's known-answer test
(bev_pool.py:145-176) and a full BASELINE-shape sample, bit-identical to the unmodified reference extension."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import rig_case

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _ref():
    from oracle import ref_ext
    if not ref_ext.available('shim'):
        pytest.skip('oracle/_shim not built (reference tree absent at build time)')
    return ref_ext


def test_shim_module_is_linked_against_the_product_library():
    ref = _ref()
    m = ref.ext('shim')
    assert hasattr(m, 'bev_pool_v2_forward') and hasattr(m, 'bev_pool_v2_backward')
    with open('/proc/self/maps') as f:
        maps = f.read()
    assert 'libfusionocc_b200.so' in maps and 'bev_pool_v2_ext_shim' in maps


def test_reference_kat_through_the_shim(golden_dir):
    """bev_pool.py:145-176: loss 4.4, the two gradient tensors."""
    ref = _ref()
    k = np.load(os.path.join(golden_dir, 'kat_bev_pool_v2.npz'))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    depth, feat = t(k['depth']), t(k['feat'])
    rd, rf, rb = t(k['ranks_depth']), t(k['ranks_feat']), t(k['ranks_bev'])
    kept = torch.ones(rb.shape[0], device=DEV, dtype=torch.bool)
    kept[1:] = rb[1:] != rb[:-1]
    st = torch.where(kept)[0].int()
    ln = torch.zeros_like(st)
    ln[:-1] = st[1:] - st[:-1]
    ln[-1] = rb.shape[0] - st[-1]
    out = ref.forward(depth, feat, rd, rf, rb, (1, 1, 2, 2, 2), st, ln, which='shim')
    assert out.shape == (1, 2, 1, 2, 2)
    assert float(out.sum().item()) == pytest.approx(4.4, abs=1e-6)
    dg, fg = ref.backward(torch.ones_like(out), depth, feat, rd, rf, rb, which='shim')
    assert torch.allclose(dg.cpu(), torch.from_numpy(k['grad_depth']))
    assert torch.allclose(fg.cpu(), torch.from_numpy(k['grad_feat']))


def test_shim_equals_unmodified_reference_extension_at_the_headline_shape():
    ref = _ref()
    if not ref.available('v2'):
        pytest.skip('oracle/_ref not built')
    from fusionocc_b200.rig import make_out_grad, make_values
    case = rig_case('base', 1)
    sh = case['shape']
    rb, rd, rf, st, ln = (torch.from_numpy(a).to(DEV) for a in case['ranks'])
    depth, feat_nchw = make_values(sh, 1)
    d, f = depth.to(DEV), feat_nchw.to(DEV).permute(0, 1, 3, 4, 2).contiguous()
    X, Y, Z = (int(v) for v in case['gs'])
    shape = (1, Z, Y, X, sh.channels)
    a = ref.forward(d, f, rd, rf, rb, shape, st, ln, which='shim')
    b = ref.forward(d, f, rd, rf, rb, shape, st, ln, which='v2')
    assert torch.equal(a.view(torch.int32), b.view(torch.int32)), 'forward through the shim differs'
    og = make_out_grad(1, sh.channels, Z, Y, X).to(DEV)
    ga = ref.backward(og, d, f, rd, rf, rb, which='shim')
    gb = ref.backward(og, d, f, rd, rf, rb, which='v2')
    assert torch.equal(ga[0].view(torch.int32), gb[0].view(torch.int32)), 'depth_grad through the shim differs'
    assert torch.equal(ga[1].view(torch.int32), gb[1].view(torch.int32)), 'feat_grad through the shim differs'
