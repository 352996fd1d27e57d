"""ORACLE (test infrastructure) — load the reference CUDA extension built by ``build_ref.py`` and
wrap it with the autograd-level glue of ``mmdet3d/ops/bev_pool_v2/bev_pool.py:17-92`` restated in
torch (the reference's .py is not copied; its op sequence is: casts, ``new_zeros``, ext forward,
``permute(0,4,1,2,3).contiguous()``; backward: argsort by ranks_feat, interval rebuild,
``out_grad.contiguous()``, zero grads, ext backward).  GPU only.
"""
from __future__ import annotations

import importlib.util
import os

import torch

from .build_ref import NAME, built_path

_ext = None


def available() -> bool:
    return built_path() is not None and torch.cuda.is_available()


def ext():
    global _ext
    if _ext is None:
        p = built_path()
        if p is None:
            raise RuntimeError('oracle/_ref is not built (python oracle/build_ref.py in the build container)')
        spec = importlib.util.spec_from_file_location(NAME, p)
        _ext = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_ext)
    return _ext


def _intervals(ranks: torch.Tensor):
    kept = torch.ones(ranks.shape[0], device=ranks.device, dtype=torch.bool)
    kept[1:] = ranks[1:] != ranks[:-1]
    starts = torch.where(kept)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = ranks.shape[0] - starts[-1]
    return starts.contiguous(), lengths.contiguous()


def forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts, interval_lengths):
    """Reference forward incl. zero-init and the final permute copy -> (B,C,Z,Y,X)."""
    depth = depth.contiguous().float()
    feat = feat.contiguous().float()
    out = feat.new_zeros(bev_feat_shape)
    ext().bev_pool_v2_forward(depth, feat, out, ranks_depth.contiguous().int(), ranks_feat.contiguous().int(),
                              ranks_bev.int().contiguous(), interval_lengths.contiguous().int(),
                              interval_starts.contiguous().int())
    return out.permute(0, 4, 1, 2, 3).contiguous()


def backward(out_grad_bczyx, depth, feat, ranks_depth, ranks_feat, ranks_bev):
    """Reference backward incl. the argsort re-sort and the out_grad un-permute copy."""
    depth = depth.contiguous().float()
    feat = feat.contiguous().float()
    order = ranks_feat.argsort()
    rf, rd, rb = ranks_feat[order], ranks_depth[order], ranks_bev[order]
    st, ln = _intervals(rf)
    depth_grad = depth.new_zeros(depth.shape)
    feat_grad = feat.new_zeros(feat.shape)
    og = out_grad_bczyx.permute(0, 2, 3, 4, 1).contiguous()
    ext().bev_pool_v2_backward(og, depth_grad, feat_grad, depth, feat, rd.contiguous(), rf.contiguous(),
                               rb.contiguous(), ln, st)
    return depth_grad, feat_grad
