"""fusionocc_b200 — B200-native (sm_100a) camera->voxel view transformation.

Drop-in for the reference's ``bev_pool_v2`` autograd op and ``LSSViewTransformer`` method
surface (bykinok/FusionOcc), backed by hand-written CUDA behind a C ABI
(``include/fusionocc_b200.h``).  No Triton, no multi-backend dispatch, no CPU fallback.

Importing the package does not load the native library; the first op call does, and raises if
``fusionocc_b200/lib/libfusionocc_b200.so`` is missing and cannot be built.
"""
from .bev_pool import (QuickCumsumCuda, TRTBEVPoolv2, VoxelPoolPlan, bev_pool_v2, bev_pool_v2_cat,
                       bev_pool_v2_with_plan, build_plan, clear_plan_cache)
from .lift import lift_prepare
from .view_transformer import LSSViewTransformer, pack_calibration, rank_prepare, rank_prepare_calib

__version__ = '0.1.0'
__all__ = ['bev_pool_v2', 'bev_pool_v2_cat', 'bev_pool_v2_with_plan', 'TRTBEVPoolv2', 'QuickCumsumCuda',
           'VoxelPoolPlan', 'build_plan', 'clear_plan_cache', 'LSSViewTransformer', 'rank_prepare',
           'rank_prepare_calib', 'pack_calibration', 'lift_prepare']
