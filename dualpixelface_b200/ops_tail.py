"""ANM tail wrapper (kept apart from ops.py only to keep that file small)."""
from __future__ import annotations

import torch

from . import _lib, ops


def anm_tail(x_nhwc: torch.Tensor, b: int, k: int, h4_global: int = 0, q_row0: int = 0, out_rows: int = 0, y_row0: int = 0) -> torch.Tensor:
    """x [B*K,H4,W4,Cs] bf16 contiguous (3 real channels, channel pitch Cs = 3 or 8) -> normals [B,3,H,W] fp32 =
    mean_k(sigmoid(bilinear x4)) * 2 - 1.  Row tiles (config 5): x holds quarter-res rows q_row0.. of an image h4_global rows
    tall and the call produces the full-res rows y_row0 .. y_row0+out_rows-1 with the global align_corners coordinates."""
    ops._req(x_nhwc, torch.bfloat16, "x")
    bk, h4, w4, c = x_nhwc.shape
    assert bk == b * k and c >= 3
    h4g = h4_global or h4
    rows = out_rows or 4 * h4
    out = torch.empty(b, 3, rows, 4 * w4, device=x_nhwc.device, dtype=torch.float32)
    _lib.check(ops.lib().dpf_anm_tail_tile(ops._p(x_nhwc), ops._p(out), b, k, h4, w4, c, h4g, q_row0, rows, y_row0, ops._stream()),
               "dpf_anm_tail")
    return out
