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

from .build_ref import TARGETS, built_path

_mods = {}


def available(which: str = 'v2') -> bool:
    return built_path(which) is not None and torch.cuda.is_available()


def ext(which: str = 'v2'):
    """``v2``: the unmodified reference extension; ``shim``: the reference's unmodified bev_pool.cpp linked against
    libfusionocc_b200.so through oracle/shim/bev_pool_shim.cpp (L0 drop-in proof); ``v1``: BEVFusion's bev_pool_ext."""
    if which not in _mods:
        p = built_path(which)
        if p is None:
            raise RuntimeError(f'oracle build "{which}" is missing (python oracle/build_ref.py --which {which} in the '
                               'build container)')
        if which == 'shim':                 # make sure the product library is the one already loaded in this process
            from fusionocc_b200 import _cabi
            _cabi.load()
        spec = importlib.util.spec_from_file_location(TARGETS[which]['name'], p)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mods[which] = m
    return _mods[which]


def _intervals(ranks: torch.Tensor):
    kept = torch.ones(ranks.shape[0], device=ranks.device, dtype=torch.bool)
    kept[1:] = ranks[1:] != ranks[:-1]
    starts = torch.where(kept)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = ranks.shape[0] - starts[-1]
    return starts.contiguous(), lengths.contiguous()


def forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts, interval_lengths,
            which: str = 'v2'):
    """Reference forward incl. zero-init and the final permute copy -> (B,C,Z,Y,X)."""
    depth = depth.contiguous().float()
    feat = feat.contiguous().float()
    out = feat.new_zeros(bev_feat_shape)
    ext(which).bev_pool_v2_forward(depth, feat, out, ranks_depth.contiguous().int(), ranks_feat.contiguous().int(),
                              ranks_bev.int().contiguous(), interval_lengths.contiguous().int(),
                              interval_starts.contiguous().int())
    return out.permute(0, 4, 1, 2, 3).contiguous()


def backward(out_grad_bczyx, depth, feat, ranks_depth, ranks_feat, ranks_bev, which: str = 'v2'):
    """Reference backward incl. the argsort re-sort and the out_grad un-permute copy."""
    depth = depth.contiguous().float()
    feat = feat.contiguous().float()
    order = ranks_feat.argsort()
    rf, rd, rb = ranks_feat[order], ranks_depth[order], ranks_bev[order]
    st, ln = _intervals(rf)
    depth_grad = depth.new_zeros(depth.shape)
    feat_grad = feat.new_zeros(feat.shape)
    og = out_grad_bczyx.permute(0, 2, 3, 4, 1).contiguous()
    ext(which).bev_pool_v2_backward(og, depth_grad, feat_grad, depth, feat, rd.contiguous(), rf.contiguous(),
                               rb.contiguous(), ln, st)
    return depth_grad, feat_grad


# ------------------------------------------------------------------------------------------------
# bev_pool v1 (projects/BEVFusion/bevfusion/ops/bev_pool/bev_pool.py:37-99), glue restated in torch around the
# UNMODIFIED extension oracle/_ref_v1: ranks = x*(W*D*B) + y*(D*B) + z*B + b, argsort, interval rebuild,
# ext forward -> (B,D,H,W,C) -> permute(0,4,1,2,3); backward: ext backward on out_grad.contiguous().
# ------------------------------------------------------------------------------------------------
class _V1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, geom_feats, ranks, B, D, H, W):
        st, ln = _intervals(ranks)
        geom_feats = geom_feats.int()
        out = ext('v1').bev_pool_forward(x, geom_feats, ln, st, B, D, H, W)
        ctx.save_for_backward(st, ln, geom_feats)
        ctx.saved_shapes = B, D, H, W
        return out

    @staticmethod
    def backward(ctx, out_grad):
        st, ln, geom_feats = ctx.saved_tensors
        B, D, H, W = ctx.saved_shapes
        x_grad = ext('v1').bev_pool_backward(out_grad.contiguous(), geom_feats, ln, st, B, D, H, W)
        return x_grad, None, None, None, None, None, None


def bev_pool_v1(feats, coords, B, D, H, W):
    ranks = coords[:, 0] * (W * D * B) + coords[:, 1] * (D * B) + coords[:, 2] * B + coords[:, 3]
    indices = ranks.argsort()
    feats, coords, ranks = feats[indices], coords[indices], ranks[indices]
    x = _V1.apply(feats, coords, ranks, B, D, H, W)
    return x.permute(0, 4, 1, 2, 3).contiguous()
