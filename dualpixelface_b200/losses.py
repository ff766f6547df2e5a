"""Losses that consume the path (kept in PyTorch): the reference's SMOOTHL1Loss and COSINELoss with its
selector's output names (src/loss/loss_selector.py:29-42 -> '<name>_loss', 'final_loss')."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _masked_mean(val, mask):
    """mean of val over mask > 0 without boolean-mask gathers (their backward is an index_put scatter)."""
    while mask.dim() < val.dim():
        mask = mask.unsqueeze(-1)
    m = (mask > 0).to(val.dtype)
    return (val * m).sum() / (m.sum() * (val.numel() // m.numel())).clamp_min(1.0)


def smooth_l1(pred, batch, weights):
    """src/loss/depth/smoothL1.py:15-49, 'given' conversion with a disparity target: weighted sum over the heads."""
    n = pred.shape[1]
    ws = [1.0] if n == 1 else list(weights)
    assert len(ws) == n
    gt = batch["disp"]
    if "mask" in batch:
        return sum(ws[i] * _masked_mean(F.smooth_l1_loss(pred[:, i], gt, reduction="none"), batch["mask"]) for i in range(n))
    return sum(ws[i] * F.smooth_l1_loss(pred[:, i], gt) for i in range(n))


def cosine(pred, batch):
    """src/loss/normal/cosine.py:35-55 (masked branch, one prediction); note the element-wise similarity of :18-26."""
    p = pred.permute(0, 3, 4, 1, 2)                                   # [B,H,W,1,3]
    g = batch["normal"].permute(0, 2, 3, 1)                           # [B,H,W,3]
    p = p / torch.norm(p, p=2, dim=-1, keepdim=True).clamp_min(1e-6)
    g = g / torch.norm(g, p=2, dim=-1, keepdim=True).clamp_min(1e-6)
    a = p[..., 0, :]
    den = (torch.norm(a, p=2, dim=-1, keepdim=True) * torch.norm(g, p=2, dim=-1, keepdim=True)).clamp_min(1e-6)
    return _masked_mean(1.0 - ((a * g) / den).clamp(-1.0, 1.0), batch["mask"])


class LossModel:
    def __init__(self, option):
        self.types = list(option.model.loss_type)
        self.lambdas = list(option.model.lambdas)
        self.weights = list(option.model.loss_weight)

    def forward(self, results, batch):
        out, total = {}, 0.0
        for name, lam in zip(self.types, self.lambdas):
            if name == "smoothL1":
                val = smooth_l1(results["pred_depth"], batch, self.weights)
            elif name == "cosine":
                if results.get("pred_normal") is None:
                    continue
                val = cosine(results["pred_normal"], batch)
            else:
                raise NotImplementedError(f"wrong loss type : {name}")
            out[f"{name}_loss"] = val
            total = total + lam * val
        out["final_loss"] = total
        return out
