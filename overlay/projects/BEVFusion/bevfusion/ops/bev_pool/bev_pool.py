"""Drop-in overlay for ``projects/BEVFusion/bevfusion/ops/bev_pool/bev_pool.py`` of the reference tree:
``bev_pool(feats, coords, B, D, H, W)`` (:85-99) on the B200-native sort + splat.  The ``bev_pool_ext``
pybind extension is no longer needed."""
from fusionocc_b200.pool_v1 import bev_pool  # noqa: F401

__all__ = ['bev_pool']
