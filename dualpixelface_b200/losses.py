"""Losses that consume the path: the reference's SMOOTHL1Loss and COSINELoss with its selector's output names
(src/loss/loss_selector.py:29-42 -> '<name>_loss', 'final_loss').

On CUDA tensors both losses -- value AND gradient -- come from ONE pass of ``dpf_fused_losses`` (csrc/losses.cu, SURVEY.md 8f-2:
"losses fused into the epilogue"): the kernel reads mask, targets and predictions once, writes the unnormalised gradients and
deterministic partial sums; ``FusedLossFn`` applies the scalars (1 / mask count, head weights, lambdas).  The plain PyTorch
formulation below (dense masked means, no boolean-mask gathers) is the CPU path of the host-logic tests and the statement the
kernel is checked against (tests/test_gpu_kernels.py::test_fused_losses)."""
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


class FusedLossFn(torch.autograd.Function):
    """(smoothL1, cosine | None) from dpf_fused_losses; backward = saved unnormalised gradients x scalars."""

    @staticmethod
    def forward(ctx, pred_depth, pred_normal, disp, mask, normal, weights):
        from . import _lib, ops
        b, n, h, w = pred_depth.shape
        pd = pred_depth.float().contiguous()
        pn = pred_normal.float().contiguous().view(b, 3, h, w) if pred_normal is not None else None
        dev = pd.device
        g_depth = torch.empty_like(pd)
        g_normal = torch.empty_like(pn) if pn is not None else None
        ws = torch.empty(int(ops.lib().dpf_fused_losses_ws_floats(b * h * w)), device=dev, dtype=torch.float32)
        sums = torch.empty(n + 2, device=dev, dtype=torch.float32)
        m = mask.float().contiguous() if mask is not None else None
        gn = normal.float().contiguous() if pn is not None else None
        _lib.check(ops.lib().dpf_fused_losses(ops._p(pd), n, ops._p(disp.float().contiguous()), ops._p(m), ops._p(pn), ops._p(gn),
                                              ops._p(g_depth), ops._p(g_normal), ops._p(ws), ops._p(sums), b, h, w, ops._stream()),
                   "dpf_fused_losses")
        wts = torch.tensor([1.0] if n == 1 else list(weights), device=dev, dtype=torch.float32)
        assert wts.numel() == n
        count = sums[n].clamp_min(1.0)
        l1 = (wts * sums[:n]).sum() / count
        lc = sums[n + 1] / (3.0 * count) if pn is not None else None
        ctx.save_for_backward(g_depth, g_normal if g_normal is not None else g_depth.new_zeros(0), wts / count, 1.0 / (3.0 * count))
        ctx.has_normal, ctx.shapes = pn is not None, (pred_depth.shape, pred_normal.shape if pred_normal is not None else None)
        return (l1, lc) if pn is not None else (l1, l1.new_zeros(()))

    @staticmethod
    def backward(ctx, d_l1, d_lc):
        g_depth, g_normal, wc, nc = ctx.saved_tensors
        gd = g_depth * (d_l1 * wc).view(1, -1, 1, 1)
        gn = (g_normal * (d_lc * nc)).view(ctx.shapes[1]) if ctx.has_normal else None
        return gd, gn, None, None, None, None


class LossModel:
    def __init__(self, option):
        self.types = list(option.model.loss_type)
        self.lambdas = list(option.model.lambdas)
        self.weights = list(option.model.loss_weight)

    def forward(self, results, batch):
        pd = results["pred_depth"]
        if (pd.is_cuda and pd.shape[1] <= 4 and set(self.types) <= {"smoothL1", "cosine"} and "smoothL1" in self.types
                and ("mask" in batch or "cosine" not in self.types)):     # (the reference's unmasked cosine branch is a different formula)
            return self._forward_fused(results, batch)
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

    def _forward_fused(self, results, batch):
        """One kernel pass for every loss of the configuration (smoothL1 [+ cosine])."""
        pn = results.get("pred_normal") if "cosine" in self.types else None
        l1, lc = FusedLossFn.apply(results["pred_depth"], pn, batch["disp"], batch.get("mask"), batch["normal"] if pn is not None else None,
                                   self.weights)
        out, total = {}, 0.0
        for name, lam in zip(self.types, self.lambdas):
            if name == "cosine" and pn is None:
                continue
            val = l1 if name == "smoothL1" else lc
            out[f"{name}_loss"] = val
            total = total + lam * val
        out["final_loss"] = total
        return out
