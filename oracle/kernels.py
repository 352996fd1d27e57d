"""ORACLE (test infrastructure, NOT product code) — ctypes binding of
``bevpool_oracle.c`` plus the autograd-level glue of the reference op
(``/root/reference/mmdet3d/ops/bev_pool_v2/bev_pool.py:17-92``) restated on
numpy arrays.  Parity status: PINNED (see ``bevpool_oracle.c`` header).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Tuple

import numpy as np

from . import rank_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'liboracle_bevpool.so')
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, 'bevpool_oracle.c')
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_bev_pool_v2_fwd.argtypes = [ctypes.c_int, ctypes.c_int, _f32p, _f32p, _i32p, _i32p,
                                                _i32p, _i32p, _i32p, _f32p]
        _lib.oracle_bev_pool_v2_fwd.restype = None
        _lib.oracle_bev_pool_v2_bwd.argtypes = [ctypes.c_int, ctypes.c_int, _f32p, _f32p, _f32p, _i32p, _i32p,
                                                _i32p, _i32p, _i32p, _f32p, _f32p]
        _lib.oracle_bev_pool_v2_bwd.restype = None
        for fn in (_lib.oracle_permute_to_bczyx, _lib.oracle_permute_to_bzyxc):
            fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, _f32p, _f32p]
            fn.restype = None
    return _lib


def _f(a):
    return a.ctypes.data_as(_f32p)


def _i(a):
    return a.ctypes.data_as(_i32p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def bev_pool_v2_forward_bzyxc(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                              interval_starts, interval_lengths) -> np.ndarray:
    """QuickCumsumCuda.forward (bev_pool.py:17-41): returns fp32 (B,Z,Y,X,C)."""
    depth = _c(depth, np.float32)
    feat = _c(feat, np.float32)                 # logical (B,N,H,W,C)
    rd, rf, rb = _c(ranks_depth, np.int32), _c(ranks_feat, np.int32), _c(ranks_bev, np.int32)
    st, ln = _c(interval_starts, np.int32), _c(interval_lengths, np.int32)
    out = np.zeros(tuple(int(s) for s in bev_feat_shape), dtype=np.float32)      # :27 new_zeros
    c = feat.shape[-1]
    lib().oracle_bev_pool_v2_fwd(c, ln.shape[0], _f(depth), _f(feat), _i(rd), _i(rf), _i(rb), _i(st), _i(ln), _f(out))
    return out


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                interval_starts, interval_lengths) -> np.ndarray:
    """bev_pool.py:86-92 — forward + permute(0,4,1,2,3).contiguous() -> (B,C,Z,Y,X)."""
    o = bev_pool_v2_forward_bzyxc(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                                  interval_starts, interval_lengths)
    B, Z, Y, X, C = o.shape
    res = np.empty((B, C, Z, Y, X), dtype=np.float32)
    lib().oracle_permute_to_bczyx(B, C, Z * Y * X, _f(o), _f(res))
    return res


def bev_pool_v2_backward(out_grad_bczyx, depth, feat, ranks_depth, ranks_feat, ranks_bev
                         ) -> Tuple[np.ndarray, np.ndarray]:
    """QuickCumsumCuda.backward (bev_pool.py:44-83) fed with the gradient of the
    permuted output.  Returns (depth_grad like depth, feat_grad (B,N,H,W,C))."""
    g = _c(out_grad_bczyx, np.float32)
    B, C, Z, Y, X = g.shape
    g_bzyxc = np.empty((B, Z, Y, X, C), dtype=np.float32)                         # :69 out_grad.contiguous()
    lib().oracle_permute_to_bzyxc(B, C, Z * Y * X, _f(g), _f(g_bzyxc))
    depth = _c(depth, np.float32)
    feat = _c(feat, np.float32)
    rb, rd, rf, st, ln = rank_oracle.backward_resort(_c(ranks_bev, np.int32), _c(ranks_depth, np.int32),
                                                     _c(ranks_feat, np.int32))   # :47-57
    rb, rd, rf = _c(rb, np.int32), _c(rd, np.int32), _c(rf, np.int32)
    depth_grad = np.zeros_like(depth)                                             # :67
    feat_grad = np.zeros_like(feat)                                               # :68
    lib().oracle_bev_pool_v2_bwd(C, ln.shape[0], _f(g_bzyxc), _f(depth), _f(feat), _i(rd), _i(rf), _i(rb),
                                 _i(st), _i(ln), _f(depth_grad), _f(feat_grad))
    return depth_grad, feat_grad
