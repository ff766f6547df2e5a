"""Weight gradient of the stride-1 convolution kinds on the tcgen05 wgrad kernel (dpf_conv3d_wgrad)."""
from __future__ import annotations

import torch

from . import _lib, ops
from .layers import KIND_1x1x1, KIND_1x3x3, KIND_3x3x3

_TAPS = {KIND_3x3x3: (3, 3, 3), KIND_1x3x3: (1, 3, 3), KIND_1x1x1: (1, 1, 1)}


def conv3d_wgrad(x: torch.Tensor, dz: torch.Tensor, kind: int, cin: int | None = None) -> torch.Tensor:
    """x [B,D,H,W,Cx] bf16, dz [B,D,H,W,Cout] bf16 (or fp32/bf16 [...,1] for the heads) -> dW [Cout,Cin,kd,kh,kw] fp32."""
    ops._req(x, torch.bfloat16, "x")
    b, d, h, w, cx = x.shape
    cin = cin or cx
    cout = dz.shape[-1]
    if cout % 8 != 0:                                   # heads: pad the gradient channels to 8
        pad = torch.zeros(*dz.shape[:-1], 8 * ((cout + 7) // 8), device=dz.device, dtype=torch.bfloat16)
        pad[..., :cout] = dz
        dz = pad
    dz = dz.to(torch.bfloat16).contiguous()
    kd, kh, kw = _TAPS[kind]
    ntaps = kd * kh * kw
    dw = torch.zeros(ntaps, cin, cout, device=x.device, dtype=torch.float32)
    for co in range(0, cout, 32):
        n = min(32, cout - co)
        part = dw if cout <= 32 else torch.zeros(ntaps, cin, n, device=x.device, dtype=torch.float32)
        _lib.check(ops.lib().dpf_conv3d_wgrad(kind, ops._p(x), ops._p(dz), ops._p(part), b, d, h, w, cin, cx, 0, n, dz.shape[-1], co,
                                              ops._stream()), "dpf_conv3d_wgrad")
        if part is not dw:
            dw[:, :, co:co + n] = part
    return dw.permute(2, 1, 0).reshape(cout, cin, kd, kh, kw).contiguous()
