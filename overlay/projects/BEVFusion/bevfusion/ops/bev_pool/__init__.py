# Drop-in overlay for projects/BEVFusion/bevfusion/ops/bev_pool/__init__.py (same export).
from .bev_pool import bev_pool

__all__ = ['bev_pool']
