"""temporal frames written straight into channel slices of one voxel tensor
(bev_pool_v2_cat / fo_bev_pool_v2_forward_slice / fo_bev_pool_v2_backward_slice).  The bar: bit-identical to
torch.cat of the per-frame bev_pool_v2 results (fusion_occ.py:316-326 pattern), forward and both gradients,
for contiguous and channels-last-3d incoming gradients, trusted plans and caller-supplied ranks."""
import numpy as np
import pytest
import torch

from fusionocc_b200 import LSSViewTransformer, bev_pool_v2, bev_pool_v2_cat
from fusionocc_b200.rig import SHAPES, make_calibration, make_values

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _frames(name, B, n_frames, chans):
    sh = SHAPES[name]
    vt = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels,
                            collapse_z=False)
    X, Y, Z = vt._grid_xyz()
    frames = []
    for i in range(n_frames):
        cal = [c.to(DEV) for c in make_calibration(sh, B, frame_shift=(i % 2 == 1))]
        coor = vt.get_lidar_coor(*cal)
        rb, rd, rf, st, ln = vt.voxel_pooling_prepare_v2(coor)
        depth, feat = make_values(sh, B)
        g = torch.Generator().manual_seed(100 + i)
        N, H, W = feat.shape[1], feat.shape[3], feat.shape[4]
        f = torch.randn(B, N, H, W, chans[i], generator=g)
        d = torch.rand(depth.shape, generator=g)
        frames.append((d.to(DEV), f.to(DEV), rd, rf, rb, st, ln))
    return frames, (B, Z, Y, X)


def _bits(t):
    return t.detach().contiguous().view(torch.int32)


@pytest.mark.parametrize('chans,cl3d', [((8, 8), False), ((32, 16, 20), False), ((8, 12), True)])
def test_cat_equals_concatenation_bitwise(chans, cl3d):
    frames, (B, Z, Y, X) = _frames('small', 2, len(chans), chans)
    leaves_a = [(f[0].clone().requires_grad_(), f[1].clone().requires_grad_()) for f in frames]
    leaves_b = [(f[0].clone().requires_grad_(), f[1].clone().requires_grad_()) for f in frames]
    want = torch.cat([bev_pool_v2(d, ft, f[2], f[3], f[4], (B, Z, Y, X, ft.shape[-1]), f[5], f[6])
                      for (d, ft), f in zip(leaves_a, frames)], dim=1)
    got = bev_pool_v2_cat([(d, ft) + f[2:] for (d, ft), f in zip(leaves_b, frames)], (B, Z, Y, X, chans[0]))
    assert got.shape == want.shape and got.is_contiguous()
    assert torch.equal(_bits(got), _bits(want))
    og = torch.randn(want.shape, generator=torch.Generator().manual_seed(3)).to(DEV)
    if cl3d:
        og = og.contiguous(memory_format=torch.channels_last_3d)
    want.backward(og)
    got.backward(og)
    for (da, fa), (db, fb) in zip(leaves_a, leaves_b):
        assert torch.equal(_bits(da.grad), _bits(db.grad)), 'depth_grad'
        assert torch.equal(_bits(fa.grad), _bits(fb.grad)), 'feat_grad'


def test_slice_leaves_other_channels_untouched_and_handles_unsorted_ranks():
    """Caller-supplied, NOT voxel-sorted intervals take the order-agnostic path; only the slice is written."""
    from fusionocc_b200.bev_pool import build_plan, native_forward
    frames, (B, Z, Y, X) = _frames('small', 1, 1, (8,))
    d, ft, rd, rf, rb, st, ln = frames[0]
    perm = torch.randperm(st.numel(), generator=torch.Generator().manual_seed(0)).to(DEV)
    st2, ln2 = st[perm].contiguous(), ln[perm].contiguous()
    want = bev_pool_v2(d, ft, rd, rf, rb, (B, Z, Y, X, 8), st, ln)
    wide = torch.full((B, 20, Z, Y, X), 7.0, device=DEV)
    plan = build_plan(rb, st2, ln2, B, Z * Y * X)
    native_forward(d.contiguous(), ft.contiguous(), rd, rf, rb, st2, ln2, (B, Z, Y, X, 8), plan, out=wide,
                   c_total=20, c_offset=5)
    assert plan.flags() & 1, 'shuffled intervals must be detected as unsorted'
    assert torch.equal(_bits(wide[:, 5:13]), _bits(want))
    assert bool((wide[:, :5] == 7.0).all()) and bool((wide[:, 13:] == 7.0).all())


def test_cat_against_the_oracle_directly():
    """bev_pool_v2_cat against the C restatement of the reference kernels (oracle/kernels.py) itself —
    not against the repo's own op: every channel slice and both gradients of every frame, bit for bit."""
    from oracle import kernels as ok
    chans = (16, 8)
    frames, (B, Z, Y, X) = _frames('small', 2, len(chans), chans)
    leaves = [(f[0].clone().requires_grad_(), f[1].clone().requires_grad_()) for f in frames]
    got = bev_pool_v2_cat([(d, ft) + f[2:] for (d, ft), f in zip(leaves, frames)], (B, Z, Y, X, chans[0]))
    og = torch.randn(got.shape, generator=torch.Generator().manual_seed(9))
    got.backward(og.to(DEV))
    off = 0
    for (d, ft), f, c in zip(leaves, frames, chans):
        rd, rf, rb, st, ln = (t.cpu().numpy() for t in f[2:])
        dn, fn = f[0].cpu().numpy(), f[1].cpu().numpy()
        want = ok.bev_pool_v2(dn, fn, rd, rf, rb, (B, Z, Y, X, c), st, ln)
        assert np.array_equal(got[:, off:off + c].detach().cpu().contiguous().numpy().view(np.uint32),
                              np.ascontiguousarray(want).view(np.uint32)), f'slice [{off},{off + c})'
        wdg, wfg = ok.bev_pool_v2_backward(np.ascontiguousarray(og[:, off:off + c].numpy()), dn, fn, rd, rf, rb)
        assert np.array_equal(d.grad.cpu().numpy().view(np.uint32), np.ascontiguousarray(wdg).view(np.uint32))
        assert np.array_equal(ft.grad.cpu().numpy().view(np.uint32), np.ascontiguousarray(wfg).view(np.uint32))
        off += c
