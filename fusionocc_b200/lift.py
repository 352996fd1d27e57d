"""``lift_prepare`` — the step before the splat, fused.

The reference view transformers end their ``forward`` with
(``/root/reference/projects/FusionOcc/fusionocc/necks/view_transformer.py:329-336``)::

    x = self.depth_net(x)
    depth_digit = x[:, :self.D, ...]
    tran_feat = x[:, self.D:self.D + self.out_channels, ...]
    depth = depth_digit.softmax(dim=1)
    return self.view_transform(input, depth, tran_feat)

and the op then makes ``feat.permute(0,1,3,4,2).contiguous().float()`` (``bev_pool.py:20-21``).  ``lift_prepare``
does the softmax, the channel split, the NCHW -> NHWC transpose and the fp16/bf16 -> fp32 conversion in ONE native
pass (``fo_lift_prepare_forward``) and its backward applies the softmax Jacobian and the inverse transpose in one
pass (``fo_lift_prepare_backward``).  fp32 arithmetic like the reference (``exp(x - max) / sum``); the softmax is
within rtol 1e-5 of torch's (different summation order), the features are exact.  No CPU path.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _cabi
from .bev_pool import _p, _require_cuda, _stream

__all__ = ['lift_prepare']

_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


class _LiftPrepare(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, D, C):
        _require_cuda(x)
        if x.dtype not in _DTYPES:
            raise TypeError(f'lift_prepare supports float32 / float16 / bfloat16 inputs, got {x.dtype}')
        if x.dim() != 4 or x.shape[1] < D + C:
            raise ValueError(f'x must be (B*N, >= D + C, H, W); got {tuple(x.shape)} with D={D}, C={C}')
        x = x.contiguous()
        BN, c_in, H, W = x.shape
        depth = torch.empty((BN, D, H, W), dtype=torch.float32, device=x.device)
        feat = torch.empty((BN, H, W, C), dtype=torch.float32, device=x.device)
        lib = _cabi.load()
        with torch.cuda.device(x.device):
            _cabi.check(lib.fo_lift_prepare_forward(_stream(x.device), _p(x), _DTYPES[x.dtype], BN, c_in, D, C, H * W,
                                                    _p(depth), _p(feat)), 'fo_lift_prepare_forward')
        ctx.save_for_backward(depth)
        ctx.meta = (BN, c_in, D, C, H, W, x.dtype)
        ctx.mark_non_differentiable()
        return depth, feat

    @staticmethod
    def backward(ctx, depth_grad, feat_grad):
        (depth,) = ctx.saved_tensors
        BN, c_in, D, C, H, W, dtype = ctx.meta
        dev = depth.device
        dg = torch.zeros_like(depth) if depth_grad is None else depth_grad.contiguous().float()
        fg = (torch.zeros((BN, H, W, C), dtype=torch.float32, device=dev) if feat_grad is None
              else feat_grad.contiguous().float())
        x_grad = torch.empty((BN, c_in, H, W), dtype=dtype, device=dev)
        lib = _cabi.load()
        with torch.cuda.device(dev):
            _cabi.check(lib.fo_lift_prepare_backward(_stream(dev), _p(depth), _p(dg), _p(fg), BN, c_in, D, C, H * W,
                                                     _p(x_grad), _DTYPES[dtype]), 'fo_lift_prepare_backward')
        return x_grad, None, None


def lift_prepare(x: torch.Tensor, D: int, C: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """``x``: depth-net output ``(B*N, D + C [+ extra], H, W)``, fp32 / fp16 / bf16.

    Returns ``(depth, feat_nhwc)``: ``depth = x[:, :D].float().softmax(dim=1)`` as fp32 ``(B*N, D, H, W)`` and
    ``feat_nhwc = x[:, D:D+C].float().permute(0, 2, 3, 1)`` as fp32 contiguous ``(B*N, H, W, C)`` — the layout
    ``bev_pool_v2`` gathers feature rows from, so its ``feat.contiguous().float()`` becomes a no-op.
    ``feat_nhwc.permute(0, 3, 1, 2)`` is the NCHW-shaped view the reference's ``view_transform`` expects.
    """
    return _LiftPrepare.apply(x, int(D), int(C))
