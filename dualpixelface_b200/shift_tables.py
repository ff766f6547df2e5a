"""Host-side sampling tables for the ASM sub-pixel shift (bit-exact coordinates by construction).

The reference resamples a feature map with ``F.grid_sample`` on a grid whose rows are displaced by a scalar and
whose columns are not (src/module/asm/asm.py:21-49, 87-127).  Both axes are therefore separable: the source row
depends only on the output row, the source column only on the output column.  This module evaluates the
reference's own fp32 op sequence (grid normalisation in asm.py:32-41, then ATen's grid_sampler un-normalisation,
GridSampler.h ``grid_sampler_unnormalize``) once per row and once per column on the host, with torch CPU fp32
ops, and hands the kernel (index, weight) pairs.  The CUDA kernel never recomputes a coordinate.

Sample order is the reference's: nearest, bilinear, phase (asm.py:92-125).
  nearest : grid_sample(mode='nearest') with the DEFAULT align_corners=False on a grid normalised for
            align_corners=True (asm.py:96) -> source index nearbyint(((g+1)*n-1)/2), zero padding.
  bilinear: align_corners=True (asm.py:101-102) -> src = ((g+1)/2)*(n-1), two taps per axis, zero padding.
  phase   : Fourier shift along rows (asm.py:59-75,112-125).  For an INTEGER displacement this is exactly a
            circular roll of the rows (wrap-around, not zero padding); that is the only case the reference ever
            executes, because make_grid caches the first level (shift -1/+1, asm.py:29-30,56-57).  A fractional
            phase shift is a dense length-H Dirichlet interpolation and is not expressible as two taps.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch


def _normalised(n: int, delta: float) -> torch.Tensor:
    """asm.py:32-41: (arange(0.0, n) + delta) / (n - 1) * 2.0 - 1.0 in fp32."""
    v = torch.arange(0.0, n) + torch.tensor(float(delta))
    return v / (n - 1) * 2.0 - 1.0


def _nearest_axis(g: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    src = ((g + 1.0) * n - 1.0) / 2.0                      # align_corners=False un-normalisation
    idx = torch.round(src)                                 # round-half-to-even == nearbyint
    ok = (idx >= 0) & (idx <= n - 1)
    i0 = torch.where(ok, idx, torch.full_like(idx, -1.0)).to(torch.int32)
    idx2 = torch.stack([i0, torch.full_like(i0, -1)], dim=1)
    wts = torch.stack([torch.ones(n), torch.zeros(n)], dim=1)
    return idx2, wts


def _bilinear_axis(g: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    src = ((g + 1.0) / 2.0) * (n - 1)                      # align_corners=True un-normalisation
    lo = torch.floor(src)
    hi = lo + 1.0
    w_lo = hi - src                                        # (ix_se - ix) of ATen's bilinear weights
    w_hi = src - lo
    def valid(i):
        return torch.where((i >= 0) & (i <= n - 1), i, torch.full_like(i, -1.0)).to(torch.int32)
    return torch.stack([valid(lo), valid(hi)], dim=1), torch.stack([w_lo, w_hi], dim=1)


def _roll_axis(n: int, delta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    if float(delta) != float(int(delta)):
        raise NotImplementedError(
            "fractional phase (Fourier) shift is a dense Dirichlet interpolation along H; only the integer shifts the "
            "reference actually executes (cached first level, +-1 row) are built")
    i0 = ((torch.arange(n) + int(delta)) % n).to(torch.int32)
    return torch.stack([i0, torch.full_like(i0, -1)], dim=1), torch.stack([torch.ones(n), torch.zeros(n)], dim=1)


def build_tables(h: int, w: int, disp: float, direction: str, modes: Sequence[bool] = (True, True, True)) -> Dict[str, torch.Tensor]:
    """Tables for one (shift, direction): ri/rw [S,H,2], ci/cw [S,W,2] (int32 / fp32, CPU)."""
    sign = 1.0 if direction == "forward" else -1.0
    delta = float(sign * disp)
    gy, gx = _normalised(h, delta), _normalised(w, 0.0)
    ri, rw, ci, cw = [], [], [], []
    nearest, bilinear, phase = modes
    if nearest:
        a, b = _nearest_axis(gy, h); ri.append(a); rw.append(b)
        a, b = _nearest_axis(gx, w); ci.append(a); cw.append(b)
    if bilinear:
        a, b = _bilinear_axis(gy, h); ri.append(a); rw.append(b)
        a, b = _bilinear_axis(gx, w); ci.append(a); cw.append(b)
    if phase:
        a, b = _roll_axis(h, delta); ri.append(a); rw.append(b)
        a, b = _roll_axis(w, 0.0); ci.append(a); cw.append(b)
    return {"ri": torch.stack(ri).contiguous(), "rw": torch.stack(rw).float().contiguous(),
            "ci": torch.stack(ci).contiguous(), "cw": torch.stack(cw).float().contiguous()}
