"""Shared test helpers (CPU side): build oracle inputs for a rig shape."""
import hashlib

import numpy as np
import torch

from fusionocc_b200.rig import SHAPES, make_calibration, make_values
from oracle import rank_oracle as ro


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rig_case(name: str, B: int, frame_shift: bool = False, with_ranks: bool = True):
    """Returns dict with grid, frustum, calib (torch CPU), coor (torch CPU), oracle ranks (int64-exact)."""
    sh = SHAPES[name]
    lb, itv, gs = ro.create_grid_infos(**sh.grid_cfg())
    fr = ro.create_frustum(sh.depth_cfg, sh.input_size, sh.downsample)
    cal = make_calibration(sh, B, frame_shift=frame_shift)
    coor = ro.get_lidar_coor(fr, *cal)
    ranks = ro.voxel_pooling_prepare_v2(coor.numpy(), lb, itv, gs, 'int64') if with_ranks else None
    return dict(shape=sh, lb=lb, itv=itv, gs=gs, frustum=fr, calib=cal, coor=coor, ranks=ranks)


def canon(rb, rd, rf):
    order = np.lexsort((rd, rb))
    return rb[order], rd[order], rf[order]
