"""Weight gradient of the 3-D convolution kinds on the tcgen05 wgrad kernel (dpf_conv3d_wgrad)."""
from __future__ import annotations

import torch

from . import _lib, ops
from .layers import KIND_1x1x1, KIND_1x3x3, KIND_3x3x3, KIND_S2, KIND_T2

_TAPS = {KIND_3x3x3: (3, 3, 3), KIND_1x3x3: (1, 3, 3), KIND_1x1x1: (1, 1, 1), KIND_S2: (3, 3, 3)}


def conv3d_wgrad(x: torch.Tensor, dz: torch.Tensor, kind: int, cin: int | None = None) -> torch.Tensor:
    """x [B,D,H,W,Cx] bf16, dz [B,Dz,Hz,Wz,Cout] bf16 (or fp32/bf16 [...,1] for the heads) -> dW fp32 in the layout of the
    layer's nn.Parameter: [Cout,Cin,kd,kh,kw] for the conv kinds, [Cin,Cout,3,3,3] for the transposed kind.

    Stride 2 (KIND_S2): dz lives on the ceil(x/2) grid; 32 input channels per launch.  Transposed (KIND_T2): x is the coarse
    input and dz the fine output gradient; dW[ci][co][k] = sum_i x[i,ci] * dz[2i+k-1,co] is the stride-2 weight gradient with
    the two tensors swapped, which directly yields the ConvTranspose3d layout."""
    if kind == KIND_T2:
        return conv3d_wgrad(dz.to(torch.bfloat16).contiguous(), x, KIND_S2)
    ops._req(x, torch.bfloat16, "x")
    b, d, h, w, cx = x.shape
    cin = cin or cx
    cout = dz.shape[-1]
    if cout % 8 != 0:                                   # heads: pad the gradient channels to 8
        pad = torch.zeros(*dz.shape[:-1], 8 * ((cout + 7) // 8), device=dz.device, dtype=torch.bfloat16)
        pad[..., :cout] = dz
        dz = pad
    dz = dz.to(torch.bfloat16).contiguous()
    if kind == KIND_S2:
        assert dz.shape[1:4] == ((d + 1) // 2, (h + 1) // 2, (w + 1) // 2) and cin % 32 == 0, (x.shape, dz.shape)
    else:
        assert dz.shape[:4] == x.shape[:4], (x.shape, dz.shape)
    kd, kh, kw = _TAPS[kind]
    ntaps = kd * kh * kw
    # input-channel window per launch (the kernel takes 32 | 64): wider layers (gwcnet's 96-channel first layer) are split
    kwin = 32 if kind == KIND_S2 else (cin if cin <= 64 else (64 if cin % 64 == 0 else 32))
    dw = torch.zeros(ntaps, cin, cout, device=x.device, dtype=torch.float32)
    for ci in range(0, cin, kwin):
        for co in range(0, cout, 32):
            n = min(32, cout - co)
            whole = kwin == cin and cout <= 32
            part = dw if whole else torch.zeros(ntaps, kwin, n, device=x.device, dtype=torch.float32)
            _lib.check(ops.lib().dpf_conv3d_wgrad(kind, ops._p(x), ops._p(dz), ops._p(part), b, d, h, w, kwin, cx, ci, n, dz.shape[-1], co,
                                                  ops._stream()), "dpf_conv3d_wgrad")
            if not whole:
                dw[:, ci:ci + kwin, co:co + n] = part
    return dw.permute(2, 1, 0).reshape(cout, cin, kd, kh, kw).contiguous()
