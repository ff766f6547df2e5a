"""ANM tail wrapper (kept apart from ops.py only to keep that file small)."""
from __future__ import annotations

import torch

from . import _lib, ops


def anm_tail(x_nhwc: torch.Tensor, b: int, k: int) -> torch.Tensor:
    """x [B*K,H4,W4,3] bf16 contiguous -> normals [B,3,H,W] fp32 = mean_k(sigmoid(bilinear x4)) * 2 - 1."""
    ops._req(x_nhwc, torch.bfloat16, "x")
    bk, h4, w4, c = x_nhwc.shape
    assert bk == b * k and c == 3
    out = torch.empty(b, 3, 4 * h4, 4 * w4, device=x_nhwc.device, dtype=torch.float32)
    _lib.check(ops.lib().dpf_anm_tail(ops._p(x_nhwc), ops._p(out), b, k, h4, w4, ops._stream()), "dpf_anm_tail")
    return out
