# Drop-in overlay for mmdet3d/ops/bev_pool_v2/__init__.py of the reference tree (same exports, :3-5).
from .bev_pool import bev_pool_v2, TRTBEVPoolv2

__all__ = ['bev_pool_v2', 'TRTBEVPoolv2']
