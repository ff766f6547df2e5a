"""Losses that consume the path (kept in PyTorch): the reference's SMOOTHL1Loss and COSINELoss with its
selector's output names (src/loss/loss_selector.py:29-42 -> '<name>_loss', 'final_loss')."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def smooth_l1(pred, batch, weights):
    """src/loss/depth/smoothL1.py:15-49, 'given' conversion with a disparity target: weighted sum over the heads."""
    n = pred.shape[1]
    ws = [1.0] if n == 1 else list(weights)
    assert len(ws) == n
    gt = batch["disp"]
    if "mask" in batch:
        m = batch["mask"] > 0
        return sum(ws[i] * F.smooth_l1_loss(pred[:, i][m], gt[m]) for i in range(n))
    return sum(ws[i] * F.smooth_l1_loss(pred[:, i], gt) for i in range(n))


def cosine(pred, batch):
    """src/loss/normal/cosine.py:35-55 (masked branch, one prediction); note the element-wise similarity of :18-26."""
    m = batch["mask"] > 0
    p = pred.permute(0, 3, 4, 1, 2)[m]
    g = batch["normal"].permute(0, 2, 3, 1)[m]
    p = p / torch.norm(p, p=2, dim=-1, keepdim=True).clamp_min(1e-6)
    g = g / torch.norm(g, p=2, dim=-1, keepdim=True).clamp_min(1e-6)
    a = p[:, 0]
    den = (torch.norm(a, p=2, dim=-1, keepdim=True) * torch.norm(g, p=2, dim=-1, keepdim=True)).clamp_min(1e-6)
    return torch.mean(1.0 - ((a * g) / den).clamp(-1.0, 1.0))


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
