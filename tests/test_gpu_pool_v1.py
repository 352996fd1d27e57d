"""bev_pool v1 / occ_pool on the native pipeline.  Oracles: (a) a numpy restatement of the
reference kernel's sequential sum in stable-sorted order (bev_pool.py:85-99 + bev_pool_cuda.cu:21-45) — bit-exact;
(b) the reference's own pure-PyTorch occ_pool (OCC_Pool.py:39-71: index_add_, order not defined) — rtol=atol=1e-5."""
import numpy as np
import pytest
import torch

from fusionocc_b200.pool_v1 import bev_pool, occ_pool, rank_from_keys

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
RTOL = ATOL = 1e-5


from oracle.pool_v1 import pool_v1 as oracle_v1  # noqa: E402  (numpy restatement, pinned by tests/golden/occ_pool_ref.npz)


def occ_pool_pure_pytorch(feats, coords, B, D, H, W):
    """The reference's CPU/GPU-agnostic implementation, restated op for op (OCC_Pool.py:39-71)."""
    C = feats.shape[1]
    out = torch.zeros((B, D, H, W, C), dtype=feats.dtype, device=feats.device)
    flat_idx = (coords[:, 3].long() * (D * H * W) + coords[:, 2].long() * (H * W) + coords[:, 0].long() * W +
                coords[:, 1].long())
    sorted_idx = flat_idx.argsort()
    out.view(-1, C).index_add_(0, flat_idx[sorted_idx], feats[sorted_idx])
    return out.permute(0, 4, 1, 2, 3).contiguous()


def _case(n, B, D, H, W, C, seed, clustered=False):
    g = torch.Generator().manual_seed(seed)
    hi = torch.tensor([H, W, D, B])
    if clustered:                       # many points per voxel: long intervals
        coords = (torch.rand(n, 4, generator=g) ** 3 * hi * 0.3).long()
    else:
        coords = (torch.rand(n, 4, generator=g) * hi).long()
    coords = torch.minimum(coords, hi - 1).int()
    feats = torch.randn(n, C, generator=g)
    return feats, coords


@pytest.mark.parametrize('n,B,D,H,W,C,clustered', [(5000, 2, 2, 16, 12, 8, False), (20000, 1, 1, 24, 24, 80, True),
                                                   (3000, 3, 4, 10, 7, 33, False)])
def test_bev_pool_v1_bit_exact_vs_sequential_oracle(n, B, D, H, W, C, clustered):
    feats, coords = _case(n, B, D, H, W, C, 1, clustered)
    want = oracle_v1(feats.numpy(), coords.numpy(), B, D, H, W)
    f = feats.to(DEV).requires_grad_()
    got = bev_pool(f, coords.to(DEV), B, D, H, W)
    assert got.shape == (B, C, D, H, W) and got.is_contiguous()
    assert np.array_equal(got.detach().cpu().numpy().view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
    og = torch.randn(got.shape, generator=torch.Generator().manual_seed(2)).to(DEV)
    got.backward(og)
    x, y, z, b = (coords[:, i].long() for i in range(4))
    want_grad = og.cpu()[b, :, z, x, y]                 # bev_pool_cuda.cu:66-91: x_grad[i] = out_grad[voxel(i)]
    assert torch.equal(f.grad.cpu().view(torch.int32), want_grad.contiguous().view(torch.int32))


def test_occ_pool_vs_reference_pure_pytorch_and_dropped_points():
    B, D, H, W, C = 2, 8, 32, 32, 16
    feats, coords = _case(40000, B, D, H, W, C, 3)
    fd, cd = feats.to(DEV), coords.to(DEV)
    want = occ_pool_pure_pytorch(fd, cd, B, D, H, W)
    got = occ_pool(fd, cd, B, D, H, W)
    torch.testing.assert_close(got, want, rtol=RTOL, atol=ATOL)
    # out-of-grid coordinates are dropped instead of corrupting memory
    bad = cd.clone()
    bad[:100, 0] = H + 3
    bad[100:200, 3] = -1
    got2 = occ_pool(fd, bad, B, D, H, W)
    want2 = occ_pool_pure_pytorch(fd[200:], cd[200:], B, D, H, W)
    torch.testing.assert_close(got2, want2, rtol=RTOL, atol=ATOL)


def test_rank_from_keys_is_the_stable_sort():
    g = torch.Generator().manual_seed(7)
    keys = torch.randint(-2, 500, (30000,), generator=g).int()
    keys[:4000] = 17                                     # one very long bucket: the CTA-wide ordering path
    rb, order, st, ln, counts = rank_from_keys(keys.to(DEV), 500)
    nk, ni = (int(v) for v in counts[:2].tolist())
    k = keys.numpy().astype(np.int64)
    valid = np.flatnonzero((k >= 0) & (k < 500))
    o = valid[np.argsort(k[valid], kind='stable')]
    assert nk == len(o)
    assert np.array_equal(order[:nk].cpu().numpy(), o)
    assert np.array_equal(rb[:nk].cpu().numpy(), k[o])
    starts = np.flatnonzero(np.r_[True, k[o][1:] != k[o][:-1]])
    assert ni == len(starts) and np.array_equal(st[:ni].cpu().numpy(), starts)
    assert np.array_equal(ln[:ni].cpu().numpy(), np.diff(np.r_[starts, nk]))


@pytest.mark.parametrize('n,B,D,H,W,C,clustered', [(120000, 4, 1, 180, 180, 80, False), (60000, 2, 8, 64, 64, 32, True)])
def test_bev_pool_v1_vs_unmodified_bevfusion_extension(n, B, D, H, W, C, clustered):
    """f-4 parity pin: projects/BEVFusion/bevfusion/ops/bev_pool/src/* compiled unmodified by
    oracle/build_ref.py --which v1, wrapped with the reference's own op sequence (bev_pool.py:37-99, restated in
    oracle/ref_ext.py).  Forward and x_grad bit-identical (same sequential sum in stable-sorted order) at a
    BEVFusion-sized grid (180 x 180 x 1, C = 80) and a clustered 3-D case."""
    from oracle import ref_ext
    if not ref_ext.available('v1'):
        pytest.skip('oracle/_ref_v1 not built (reference tree absent at build time)')
    feats, coords = _case(n, B, D, H, W, C, 5, clustered)
    fa = feats.to(DEV).requires_grad_()
    fb = feats.to(DEV).requires_grad_()
    cd = coords.to(DEV)
    got = bev_pool(fa, cd, B, D, H, W)
    want = ref_ext.bev_pool_v1(fb, cd, B, D, H, W)
    assert got.shape == want.shape == (B, C, D, H, W)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32)), 'forward differs from bev_pool_ext'
    og = torch.randn(got.shape, generator=torch.Generator().manual_seed(6)).to(DEV)
    got.backward(og)
    want.backward(og)
    assert torch.equal(fa.grad.view(torch.int32), fb.grad.view(torch.int32)), 'x_grad differs from bev_pool_ext'


def test_occ_pool_vs_reference_generated_golden(golden_dir):
    """tests/golden/occ_pool_ref.npz was produced by EXECUTING the reference's own occ_pool_pure_pytorch
    (projects/CONet/mmdet3d_plugin/ops/occ_pooling/OCC_Pool.py:39-71) on CPU (tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(golden_dir, 'occ_pool_ref.npz'))
    B, D, H, W = (int(v) for v in g['dims'])
    got = occ_pool(torch.from_numpy(g['feats']).to(DEV), torch.from_numpy(g['coords']).to(DEV), B, D, H, W)
    torch.testing.assert_close(got.cpu(), torch.from_numpy(g['out']), rtol=RTOL, atol=ATOL)
