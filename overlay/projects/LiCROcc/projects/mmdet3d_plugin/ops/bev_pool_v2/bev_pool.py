"""Drop-in overlay for ``projects/LiCROcc/projects/mmdet3d_plugin/ops/bev_pool_v2/bev_pool.py`` — LiCROcc's
VENDORED copy of the op, which differs from ``mmdet3d/ops/bev_pool_v2/bev_pool.py`` in one place: its
``TRTBEVPoolv2`` (:108-159) takes ``(output_height, output_width, output_z)``, exports the ONNX attributes
``output_height_i / output_width_i / output_z_i`` and squeezes Z only when ``output_z == 1``.  This module hands
out that signature under the vendored name; ``bev_pool_v2`` / ``QuickCumsumCuda`` are the common op.
"""
from fusionocc_b200.bev_pool import QuickCumsumCuda, bev_pool_v2  # noqa: F401
from fusionocc_b200.bev_pool import TRTBEVPoolv2Z as TRTBEVPoolv2  # noqa: F401

__all__ = ['bev_pool_v2', 'TRTBEVPoolv2']
