# Drop-in overlay for projects/CONet/mmdet3d_plugin/ops/occ_pooling/__init__.py (same export).
from .OCC_Pool import occ_pool
