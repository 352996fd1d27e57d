"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the sibling pooling ops.

``bev_pool`` v1 (projects/BEVFusion/bevfusion/ops/bev_pool/bev_pool.py:85-99 + src/bev_pool_cuda.cu:21-45) and
``occ_pool`` (projects/CONet/mmdet3d_plugin/ops/occ_pooling/OCC_Pool.py:39-71) sum pre-multiplied point features
``feats[N, C]`` into the voxel named by ``coords[N, 4] = (x, y, z, b)``: points are sorted by voxel (ties in ascending
point index — what the device argsort yields), every voxel's sum is the sequential fp32 ``psum += x`` of the
reference kernel, and the result is returned as ``(B, C, D, H, W)``.

Parity status: PINNED — tests/golden/occ_pool_ref.npz is the output of the reference's own
``occ_pool_pure_pytorch`` executed in the build container (tests/golden/make_golden.py); on the GPU box the product is
additionally compared bit for bit with the UNMODIFIED BEVFusion extension (oracle/_ref_v1).
"""
import numpy as np


def pool_v1(feats: np.ndarray, coords: np.ndarray, B: int, D: int, H: int, W: int) -> np.ndarray:
    x, y, z, b = (coords[:, i].astype(np.int64) for i in range(4))
    key = ((b * D + z) * H + x) * W + y
    order = np.argsort(key, kind='stable')
    out = np.zeros((B * D * H * W, feats.shape[1]), np.float32)
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]]) if len(ks) else np.zeros(0, np.int64)
    ends = np.r_[starts[1:], len(ks)]
    for s, e in zip(starts, ends):
        acc = np.zeros(feats.shape[1], np.float32)
        for i in order[s:e]:
            acc = (acc + feats[i]).astype(np.float32)
        out[ks[s]] = acc
    return out.reshape(B, D, H, W, -1).transpose(0, 4, 1, 2, 3)
