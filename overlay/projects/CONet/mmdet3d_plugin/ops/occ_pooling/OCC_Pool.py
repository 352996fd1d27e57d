"""Drop-in overlay for ``projects/CONet/mmdet3d_plugin/ops/occ_pooling/OCC_Pool.py`` of the reference tree
(the same file is vendored by ``projects/SparseOcc_cvpr``): ``occ_pool(feats, coords, B, D, H, W)`` (:74-104)."""
from fusionocc_b200.pool_v1 import occ_pool  # noqa: F401

__all__ = ['occ_pool']
