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
            executes, because make_grid caches the first level (shift -1/+1, asm.py:29-30,56-57).  A FRACTIONAL
            phase shift (only reachable with ``cached_first_level=False``, the evidently intended behaviour) is not a
            table sample: ``fourier_row_shift`` evaluates the reference's rfft2 -> rotate -> irfft2 sequence itself
            (legacy C2R semantics included) with the rotation of ``phase_rotation``.
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


def phase_rotation(h: int, delta: float) -> torch.Tensor:
    """Per-row-frequency rotation of the reference's Fourier row shift by `delta` rows (asm.py:59-75; even H): complex64 [H],
    exp(2*pi*i * (delta/H) * N_r) with N_r = [0..H/2-1, -H/2..-1], evaluated with the reference's fp32 op sequence."""
    import math
    if h % 2:
        raise ValueError(f"the reference builds its frequency grid for even H only (asm.py:67), got {h}")
    deltar = torch.tensor(float(delta)) / h
    nr = torch.cat([torch.arange(0.0, math.ceil(h // 2)), torch.arange(-float(h // 2), 0.0)])
    arg = torch.tensor(2.0 * math.pi) * (deltar * nr)
    return torch.complex(torch.cos(arg), torch.sin(arg))


def fourier_row_shift(x: torch.Tensor, rot: torch.Tensor) -> torch.Tensor:
    """The phase sample for a FRACTIONAL shift on a channels-last map x [B,H,W,C]: rfft2 over (H, W) -> rotate every row
    frequency -> irfft2, i.e. asm.py:112-125 with its legacy ``torch.irfft(.., 2, onesided=False)`` semantics (a C2R transform
    over the first W/2+1 columns of the rotated spectrum; because the Nyquist row is rotated by a non-real factor this is NOT a
    pure row interpolation, so it is evaluated exactly as the reference does, with the FFT library).  Differentiable (torch ops);
    only reachable with cached_first_level=False -- the shipped reference never executes a fractional phase shift."""
    b, h, w, c = x.shape
    spec = torch.fft.rfft2(x.float(), dim=(1, 2))                                   # == fft2(x)[..., :W/2+1] for real x
    out = torch.fft.irfft2(spec * rot.view(1, h, 1, 1), s=(h, w), dim=(1, 2))
    return out.to(x.dtype)


def is_fractional(delta: float) -> bool:
    return float(delta) != float(int(delta))


def _roll_axis(n: int, delta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    assert not is_fractional(delta)
    i0 = ((torch.arange(n) + int(delta)) % n).to(torch.int32)
    return torch.stack([i0, torch.full_like(i0, -1)], dim=1), torch.stack([torch.ones(n), torch.zeros(n)], dim=1)


def build_tables(h: int, w: int, disp: float, direction: str, modes: Sequence[bool] = (True, True, True)) -> Dict[str, torch.Tensor]:
    """Tables for one (shift, direction): ri/rw [S,H,2], ci/cw [S,W,2] (int32 / fp32, CPU).  When the phase sample is requested
    for a FRACTIONAL shift it is not a table sample: the tables then cover nearest / bilinear only (S <= 2) and the extra entry
    "rot" holds phase_rotation(h, delta) for fourier_row_shift (the phase sample stays last in the sample order)."""
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
    extra = {}
    if phase and is_fractional(delta):
        extra["rot"] = phase_rotation(h, delta)
    elif phase:
        a, b = _roll_axis(h, delta); ri.append(a); rw.append(b)
        a, b = _roll_axis(w, 0.0); ci.append(a); cw.append(b)
    out = {"ri": torch.stack(ri).contiguous(), "rw": torch.stack(rw).float().contiguous(),
           "ci": torch.stack(ci).contiguous(), "cw": torch.stack(cw).float().contiguous()} if ri else {}
    out.update(extra)
    return out
