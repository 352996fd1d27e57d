"""Drop-in overlay for ``projects/LiCROcc/projects/mmdet3d_plugin/ops/bev_pool_v2/bev_pool.py`` (the vendored copy).

Copy this directory over ``mmdet3d/ops/bev_pool_v2/`` of a FusionOcc / BEVDet-family checkout (or put
``overlay/`` ahead of it on ``sys.path``): every view transformer in the tree imports
``from mmdet3d.ops.bev_pool_v2.bev_pool import bev_pool_v2`` (e.g.
``projects/FusionOcc/fusionocc/necks/view_transformer.py:11``) and picks up the B200-native op with
the unchanged signature.  The reference's pybind extension ``bev_pool_v2_ext`` is no longer needed.
"""
from fusionocc_b200.bev_pool import QuickCumsumCuda, TRTBEVPoolv2, bev_pool_v2  # noqa: F401

__all__ = ['bev_pool_v2', 'TRTBEVPoolv2']
