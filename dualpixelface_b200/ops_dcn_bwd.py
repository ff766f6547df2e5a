"""Backward of the D3D deformable convolution (dpf_dcn3d_bwd_data / dpf_dcn3d_bwd_weight) and its autograd Function.

Reference: DeformConvFunction.backward, src/module/dcn3d/functions/deform_conv_func.py:42-60.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib, ops


def dcn3d_bwd_data(x: torch.Tensor, offset: torch.Tensor, dy: torch.Tensor, weight: torch.Tensor, dx_channels: int = 64):
    """x [B,D,H,W,Cs] bf16 (Cs >= 64, channels >= Cin zero), offset [B,D,H,W,>=81] fp32, dy [B,D,H,W,64] bf16,
    weight [64,Cin,3,3,3] -> (dx [B,D,H,W,Cs] fp32, doffset like offset, fp32).  dx_channels = 32 leaves dx[..., 32:] zero
    (for callers that need no gradient there, e.g. the constant coordinate channels of the ANM volume)."""
    ops._req(x, torch.bfloat16, "x"); ops._req(offset, torch.float32, "offset"); ops._req(dy, torch.bfloat16, "dy")
    b, d, h, w, cs = x.shape
    assert dy.shape == (b, d, h, w, 64) and weight.shape[0] == 64 and weight.shape[1] <= 64
    w_t = ops.pack_conv_weight(weight.transpose(0, 1), cin_pad=64)          # [27][o/8][c -> 64][8]
    assert w_t.shape == (27, 8, 64, 8)
    dx = torch.zeros(b, d, h, w, cs, device=x.device, dtype=torch.float32)
    ocs = offset.shape[-1]
    doff = torch.empty(b, d, h, w, ocs, device=x.device, dtype=torch.float32)
    if ocs > 81:
        doff[..., 81:] = 0.0                                       # pad channels are not written by the kernel
    _lib.check(ops.lib().dpf_dcn3d_bwd_data(ops._p(x), ops._p(offset), ops._p(dy), ops._p(w_t), ops._p(dx), ops._p(doff), b, d, h, w,
                                            cs, ocs, dx_channels, ops._stream()), "dpf_dcn3d_bwd_data")
    return dx, doff


def dcn3d_bwd_weight(x: torch.Tensor, offset: torch.Tensor, dy: torch.Tensor, cin: int) -> torch.Tensor:
    """-> dW [64, cin, 3, 3, 3] fp32."""
    ops._req(x, torch.bfloat16, "x"); ops._req(offset, torch.float32, "offset"); ops._req(dy, torch.bfloat16, "dy")
    b, d, h, w, cs = x.shape
    dw = torch.zeros(27, 64, 64, device=x.device, dtype=torch.float32)
    _lib.check(ops.lib().dpf_dcn3d_bwd_weight(ops._p(x), ops._p(offset), ops._p(dy), ops._p(dw), b, d, h, w, cs, offset.shape[-1],
                                              ops._stream()), "dpf_dcn3d_bwd_weight")
    return dw[:, :cin].permute(2, 1, 0).reshape(64, cin, 3, 3, 3).contiguous()


class DCNFn(Function):
    """z = D3D(x, offset; W) (raw bf16, bias-free: the bias is folded into the following BatchNorm shift)."""

    @staticmethod
    def forward(ctx, x, offset, weight, dx_channels=64):
        wp = ops.pack_conv_weight(weight.detach(), cin_pad=64)
        ctx.save_for_backward(x, offset, weight)
        ctx.dx_channels = dx_channels
        return ops.dcn3d(x, offset, wp, 64)

    @staticmethod
    def backward(ctx, dz):
        x, offset, weight = ctx.saved_tensors
        dz = dz.to(torch.bfloat16).contiguous()
        dx, doff = dcn3d_bwd_data(x, offset, dz, weight, ctx.dx_channels)
        dw = dcn3d_bwd_weight(x, offset, dz, weight.shape[1]).to(weight.dtype)
        return dx.to(torch.bfloat16), doff, dw, None
