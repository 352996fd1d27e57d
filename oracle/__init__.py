"""ORACLE package — test infrastructure, NOT product code.

CPU restatement of the reference's algorithm for the bev_pool_v2 view-transform
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import anything
from here, and only as the checker or the timed CPU baseline — never as the
shipped path.  ``fusionocc_b200`` does not import this package.

Contents
  rank_oracle.py      numpy restatement of create_grid_infos / create_frustum /
                      voxel_pooling_prepare_v2 / backward re-sort; torch
                      op-for-op get_lidar_coor            (parity: PINNED)
  bevpool_oracle.c    plain-C fmaf restatement of the two CUDA kernels
                      (parity: PINNED by the reference KAT; by the reference
                      extension itself on the GPU box)
  kernels.py          ctypes binding of bevpool_oracle.c + the autograd-level
                      glue of bev_pool.py (casts, zero-init, re-sort, permute)
  torch_cpu_path.py   the reference-style pure-PyTorch scatter path that is
                      timed as the CPU baseline
  build_ref.py        compiles the UNMODIFIED reference CUDA extension from
                      /root/reference into oracle/_ref/ (git-ignored)
  ref_ext.py          loads oracle/_ref for GPU-side cross-checks
"""
