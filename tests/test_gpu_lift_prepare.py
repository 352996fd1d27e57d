"""depth softmax + channel split + NCHW->NHWC + fp16/bf16->fp32 in one pass (lift_prepare).
Oracle: the reference's own torch expressions (view_transformer.py:333-335, bev_pool.py:20-21).  The features
are pure data movement: exact.  The softmax sums in a different order than torch's kernel: the tolerance is the
one north_star states for fp32, rtol = atol = 1e-5 (observed differences are ~1e-7)."""
import pytest
import torch

from fusionocc_b200 import LSSViewTransformer, lift_prepare
from fusionocc_b200.rig import SHAPES, make_calibration

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
RTOL = ATOL = 1e-5


def _ref(x, D, C):
    depth = x[:, :D].float().softmax(dim=1)
    feat = x[:, D:D + C].float().permute(0, 2, 3, 1).contiguous()
    return depth, feat


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('BN,D,C,H,W,extra', [(12, 88, 32, 16, 44, 0), (3, 118, 80, 7, 13, 5), (2, 5, 3, 4, 9, 0)])
def test_forward_and_backward_vs_torch(dtype, BN, D, C, H, W, extra):
    g = torch.Generator().manual_seed(0)
    x0 = (torch.randn(BN, D + C + extra, H, W, generator=g) * 3).to(dtype).to(DEV)
    xa = x0.clone().requires_grad_()
    xb = x0.clone().requires_grad_()
    want_d, want_f = _ref(xa, D, C)
    got_d, got_f = lift_prepare(xb, D, C)
    assert got_d.dtype == torch.float32 and got_f.dtype == torch.float32 and got_f.is_contiguous()
    assert torch.equal(got_f, want_f), 'features are pure data movement: must be exact'
    torch.testing.assert_close(got_d, want_d, rtol=RTOL, atol=ATOL)
    assert torch.allclose(got_d.sum(1), torch.ones_like(got_d[:, 0]), atol=1e-5)
    gd = torch.randn(want_d.shape, generator=g).to(DEV)
    gf = torch.randn(want_f.shape, generator=g).to(DEV)
    (want_d * gd).sum().add((want_f * gf).sum()).backward()
    (got_d * gd).sum().add((got_f * gf).sum()).backward()
    assert xb.grad.dtype == dtype and xb.grad.shape == x0.shape
    tol = dict(rtol=RTOL, atol=ATOL) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)   # one half/bf16 rounding
    torch.testing.assert_close(xb.grad.float(), xa.grad.float(), **tol)
    if extra:
        assert bool((xb.grad[:, D + C:] == 0).all())


@pytest.mark.parametrize('fuse_geometry', [False, True])
def test_module_fuse_lift_matches_unfused_forward(fuse_geometry):
    """LSSViewTransformer(fuse_lift=True[, fuse_geometry=True]).forward == the reference-order forward within
    tolerance, gradients included."""
    sh = SHAPES['small']
    B = 2
    torch.manual_seed(0)
    kw = dict(in_channels=16, out_channels=sh.channels, collapse_z=False)
    vt = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, **kw).to(DEV)
    vt_f = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, fuse_lift=True,
                              fuse_geometry=fuse_geometry, **kw).to(DEV)
    vt_f.load_state_dict(vt.state_dict())
    cal = [c.to(DEV) for c in make_calibration(sh, B)]
    H, W = sh.feat_hw
    img = torch.randn(B, sh.n_cams, 16, H, W, generator=torch.Generator().manual_seed(1)).to(DEV)
    outs = []
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the 1x1 depth_net conv must not add TF32 noise to the comparison
    try:
        for m in (vt, vt_f):
            xin = img.clone().requires_grad_()
            out, depth = m([xin] + cal)
            (out.square().mean() + depth.mean()).backward()
            outs.append((out.detach(), depth.detach(), xin.grad, m.depth_net.weight.grad.clone()))
            m.zero_grad()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    for a, b, nm in zip(outs[0], outs[1], ('bev_feat', 'depth', 'input grad', 'depth_net weight grad')):
        torch.testing.assert_close(b, a, rtol=1e-4, atol=1e-4, msg=lambda m, nm=nm: f'{nm}: {m}')
