"""Launch planning for the tensor-core convolution engine (host logic, no CUDA needed to plan).

One reference layer (nn.Conv3d / nn.ConvTranspose3d [+ BatchNorm3d] [+ ReLU] [+ add]) becomes one or a few launches of
``dpf_conv3d_fwd``.  The per-launch limits of the kernel (include/dpf_sm100.h) are met by splitting:
  * output channels into chunks written at ``y_coff`` into the same output tensor;
  * (stride-2, and stride-1 layers wider than 64 channels) input channels into 32/64-wide windows read at ``x_coff``, chained through an fp32 partial sum
    (``res_pre``) so that the affine / ReLU is applied once, by the last launch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

KIND_3x3x3, KIND_S2, KIND_T2, KIND_1x3x3, KIND_1x1x1 = 0, 1, 2, 3, 4


@dataclass(frozen=True)
class Launch:
    x_coff: int
    cin: int
    y_coff: int
    cout: int
    first_k: bool      # first input-channel window of this output chunk
    last_k: bool       # last one (applies scale/shift/residual/relu)


def plan_launches(kind: int, cin: int, cout: int) -> List[Launch]:
    """Split a (cin -> cout) layer of the given kind into launches the kernel accepts."""
    if cin % 32 != 0 or cin > 64 * 4:
        raise ValueError(f"input channels must be a multiple of 32 (pad on the host), got {cin}")
    if kind in (KIND_3x3x3, KIND_1x3x3, KIND_1x1x1):
        kwin = cin if cin in (32, 64) else (64 if cin % 64 == 0 else 32)    # wider layers: input-channel windows (K-split)
        # 3x3x3: 32-wide output chunks keep every launch on the kd-fused kernel (N = 3*32), 2.2x faster than one N = 64 launch
        cchunk = 64 if (kwin == 32 and kind != KIND_3x3x3) else 32
    elif kind == KIND_S2:
        kwin, cchunk = 32, 32
    elif kind == KIND_T2:
        if cin not in (32, 64):
            raise ValueError(f"transposed kind takes Cin in (32, 64), got {cin}")
        kwin, cchunk = cin, 32
    else:
        raise ValueError(f"unknown kind {kind}")
    out: List[Launch] = []
    for co in range(0, cout, cchunk):
        n = min(cchunk, cout - co)
        nk = cin // kwin
        for ki in range(nk):
            out.append(Launch(ki * kwin, kwin, co, n, ki == 0, ki == nk - 1))
    return out


def fold_bn(weight: torch.Tensor, bias: torch.Tensor, mean: torch.Tensor, var: torch.Tensor, eps: float = 1e-5,
            conv_bias: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode BatchNorm (running statistics) as a per-channel affine applied to the raw convolution output."""
    scale = weight.float() / torch.sqrt(var.float() + eps)
    shift = bias.float() - mean.float() * scale
    if conv_bias is not None:
        shift = shift + conv_bias.float() * scale
    return scale.contiguous(), shift.contiguous()


class TCConv3d:
    """Packed weights + launch plan of one layer; call it on CUDA tensors."""

    def __init__(self, weight: torch.Tensor, kind: int, transposed: bool = False, cin_pad: Optional[int] = None):
        from . import ops
        self.kind = kind
        w = weight.detach()
        if transposed:
            w = w.transpose(0, 1)
        self.cout, cin = int(w.shape[0]), int(w.shape[1])
        self.cin = cin_pad or cin
        # stride 2: the plane-streamed kernel (dpf_conv3d_s2_fwd) takes all input channels and up to 64 (Cin 32) / 32 (Cin 64) output
        # channels per launch -- 1 launch for conv1 (32 -> 64), 2 for conv3 (64 -> 64), no fp32 partial-sum chain
        self.s2 = None
        if kind == KIND_S2 and self.cin in (32, 64) and cin == self.cin and self.cout % 8 == 0:
            step = 64 if self.cin == 32 else 32
            self.s2 = [(ops.pack_conv_weight(w[co:co + step].float()), co, min(step, self.cout - co)) for co in range(0, self.cout, step)]
        # 32 -> 1 heads: the bandwidth-bound kernel with the taps as the GEMM's N dimension (dpf_conv3d_head_fwd)
        self.head = ops.pack_head_weight(w) if (kind == KIND_3x3x3 and self.cout == 1 and cin == 32 and self.cin == 32) else None
        self.plan = plan_launches(kind, self.cin, self.cout)
        self.packed = []
        for ln in self.plan:
            wk = torch.zeros(ln.cout, ln.cin, *w.shape[2:], device=w.device, dtype=torch.float32)
            hi = min(cin, ln.x_coff + ln.cin)
            if hi > ln.x_coff:
                wk[:, : hi - ln.x_coff] = w[ln.y_coff: ln.y_coff + ln.cout, ln.x_coff: hi].float()
            self.packed.append(ops.pack_conv_weight(wk, kind=kind))

    def out_shape(self, x: torch.Tensor) -> Tuple[int, ...]:
        b, d, h, w, _ = x.shape
        if self.kind == KIND_S2:
            return b, (d + 1) // 2, (h + 1) // 2, (w + 1) // 2
        if self.kind == KIND_T2:
            return b, 2 * d, 2 * h, 2 * w
        return b, d, h, w

    def __call__(self, x: torch.Tensor, scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
                 residual: Optional[torch.Tensor] = None, relu: bool = False, out_f32: bool = False,
                 out: Optional[torch.Tensor] = None, y_coff: int = 0, slope: float = 0.0) -> torch.Tensor:
        from . import ops
        assert x.shape[-1] == self.cin, (x.shape, self.cin)
        if out is None:
            out = torch.empty(*self.out_shape(x), self.cout, device=x.device,
                              dtype=torch.float32 if out_f32 else torch.bfloat16)
        if (self.head is not None and out.dtype == torch.float32 and out.shape[-1] == 1 and y_coff == 0 and scale is None and not relu
                and (shift is None or shift.numel() == 1)):
            return ops.conv3d_head(x, self.head, residual, float(shift) if shift is not None else 0.0, out=out)
        if slope != 0.0 and not (self.kind == KIND_3x3x3 and all(ln.first_k and ln.last_k for ln in self.plan)):
            raise NotImplementedError("a LeakyReLU slope is built for single-window 3x3x3 stride-1 layers only")
        if self.s2 is not None and residual is None and out.dtype == torch.bfloat16:
            for wp, co, n in self.s2:
                ops.conv3d_s2(x, wp, n, scale[co:co + n].contiguous() if scale is not None else None,
                              shift[co:co + n].contiguous() if shift is not None else None, relu, out=out, y_coff=y_coff + co)
            return out
        partial = None
        for ln, wp in zip(self.plan, self.packed):
            sc = scale[ln.y_coff: ln.y_coff + ln.cout] if scale is not None else None
            sh = shift[ln.y_coff: ln.y_coff + ln.cout] if shift is not None else None
            if ln.first_k and ln.last_k:
                ops.conv3d(x, wp, self.kind, ln.cout, sc, sh, residual, relu, out=out, y_coff=y_coff + ln.y_coff,
                           cin=ln.cin, x_coff=ln.x_coff, slope=slope)
                continue
            # input-channel split: chain fp32 partial sums, finish with the affine
            if residual is not None:
                raise NotImplementedError("a K-split layer cannot also take a residual")
            if ln.first_k:
                partial = torch.empty(*self.out_shape(x), ln.cout, device=x.device, dtype=torch.float32)
                ops.conv3d(x, wp, self.kind, ln.cout, out=partial, cin=ln.cin, x_coff=ln.x_coff)
            elif not ln.last_k:
                ops.conv3d(x, wp, self.kind, ln.cout, residual=partial, out=partial, cin=ln.cin, x_coff=ln.x_coff,
                           res_pre=True)
            else:
                if out.dtype == torch.float32:
                    res = partial if out.shape[-1] == ln.cout else None
                    if res is None:
                        raise NotImplementedError("fp32 K-split output must be dense")
                    ops.conv3d(x, wp, self.kind, ln.cout, sc, sh, residual=partial, relu=relu, out=out, cin=ln.cin,
                               x_coff=ln.x_coff, res_pre=True)
                else:
                    tmp = torch.empty_like(partial)
                    ops.conv3d(x, wp, self.kind, ln.cout, sc, sh, residual=partial, relu=relu, out=tmp, cin=ln.cin,
                               x_coff=ln.x_coff, res_pre=True)
                    out[..., y_coff + ln.y_coff: y_coff + ln.y_coff + ln.cout] = tmp.to(out.dtype)
        return out
