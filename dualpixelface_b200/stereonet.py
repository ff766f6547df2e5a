"""STEREONET with the reference's model-class contract on the sm_100a kernels.

Reference: src/model/stereonet/mainmodel.py:31-152 (model), src/model/stereonet/modules.py (FeatureExtraction, BasicBlock,
EdgeAwareRefinement, disp_regression).  Sub-module names are the reference's, so the state_dict layout is identical (186
entries, pinned by tests/golden/state_keys_stereonet.json).  What runs where, in eval mode:

  difference volume (mainmodel.py:100-114)                      dpf_costvol_fwd, mode "diff"            [B,8,h,w,32] bf16
  4 x convbn_3d + LeakyReLU(0.2), 32 -> 1 conv (:43-51,117-120) dpf_conv3d_fwd (kd-fused tcgen05 kernel, slope in the epilogue)
  soft-argmin over the 2^k levels, no up-sampling (:123)        dpf_softargmin_fwd
  BasicBlocks of the encoder / the refinement (3x3, dil 1..8)   dpf_conv2d_tc_fwd with y = LeakyReLU(BN(conv x)) + x fused
  5x5 stride-2 stem, the 4 -> 32 conv, bilinear resizes         cuDNN / ATen (adjacent 2-D ops with 3-4 input channels)

Training: the 3-D path (difference volume, the four convbn_3d + LeakyReLU layers, the head) runs forward AND backward on the same
kernels through the autograd Functions of train_ops.py (CostVolumeFn, ConvBNAct with slope 0.2, HeadConv); the 2-D encoder and
refinement and the 1/8-resolution soft-argmin (a [B,8,h,w] tensor) go through PyTorch autograd, as the encoders of the other models do.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .layers import KIND_3x3x3, TCConv3d, fold_bn
from .models import _StereoBase
from .modules import _cb2, _cb3, cost_range


class BasicBlock(nn.Module):
    """src/model/stereonet/modules.py:10-29 -- conv2 is constructed (and lives in the state dict) but never applied (:23)."""

    def __init__(self, c, dilation):
        super().__init__()
        self.conv1 = nn.Sequential(_cb2(c, c, 3, 1, 1, dilation), nn.LeakyReLU(0.2, inplace=True))
        self.conv2 = _cb2(c, c, 3, 1, 1, dilation)
        self.dilation = dilation

    def forward(self, x):                        # training path (PyTorch autograd over cuDNN); eval runs _run_block
        return x + self.conv1(x)


class FeatureExtraction(nn.Module):
    """src/model/stereonet/modules.py:32-61."""

    def __init__(self, k, cin):
        super().__init__()
        self.k = k
        self.downsample = nn.ModuleList([nn.Conv2d(cin if i == 0 else 32, 32, 5, 2, 2) for i in range(k)])
        self.residual_blocks = nn.ModuleList([BasicBlock(32, 1) for _ in range(6)])
        self.conv_alone = nn.Conv2d(32, 32, 3, 1, 1)

    def forward(self, x):                        # training path
        for conv in self.downsample:
            x = conv(x)
        for blk in self.residual_blocks:
            x = blk(x)
        return self.conv_alone(x)


class EdgeAwareRefinement(nn.Module):
    """src/model/stereonet/modules.py:64-96."""

    def __init__(self, cin):
        super().__init__()
        self.conv2d_feature = nn.Sequential(_cb2(cin, 32, 3, 1, 1, 1), nn.LeakyReLU(0.2, inplace=True))
        self.residual_astrous_blocks = nn.ModuleList([BasicBlock(32, d) for d in (1, 2, 4, 8, 1, 1)])
        self.conv2d_out = nn.Conv2d(32, 1, 3, 1, 1)

    def forward(self, low_disparity, rgb):       # training path
        up = F.interpolate(low_disparity.unsqueeze(1), size=rgb.shape[-2:], mode="bilinear", align_corners=False)
        if rgb.shape[-1] / low_disparity.shape[-1] >= 1.5:
            up = up * 8
        x = self.conv2d_feature(torch.cat([up, rgb], 1))
        for blk in self.residual_astrous_blocks:
            x = blk(x)
        return torch.relu((up + self.conv2d_out(x).float()).squeeze(1))


def _pack_block(blk: BasicBlock):
    conv, bn = blk.conv1[0][0], blk.conv1[0][1]
    sc, sh = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
    return ops.pack_conv2d_tc_weight(conv.weight.detach().float()), sc, sh, blk.dilation


def _run_block(x, pk):
    """x + LeakyReLU_0.2(BN(conv(x))) in one launch (residual added after the activation)."""
    wp, sc, sh, dil = pk
    return ops.conv2d_tc(x, wp, 32, dil, sc, sh, residual=x, relu=True, slope=0.2, res_post=True)


class STEREONET(_StereoBase):
    def __init__(self, option):
        super().__init__()
        self._common_init(option)
        self.k = int(option.model.k)
        self.level = int(math.pow(2, self.k))
        self.costrange = cost_range(self.mindisp, self.maxdisp, self.level)       # keeps the /4 of the other models (mainmodel.py:39-40)
        self.shifts = [int(d) for d in self.costrange]
        self.feature_extraction = FeatureExtraction(self.k, option.model.input_channel)
        self.filter = nn.ModuleList([nn.Sequential(_cb3(32, 32), nn.LeakyReLU(0.2, inplace=True)) for _ in range(4)])
        self.conv3d_alone = nn.Conv3d(32, 1, 3, 1, 1)
        self.edge_aware_refinements = nn.ModuleList([EdgeAwareRefinement(4)])
        self.want_prob = True                    # the reference always returns prob_depth [B,1,2^k,h,w] (tiny at 1/8 resolution)
        self._plan = None
        self.reference_init()

    def refresh(self):
        self._plan = None
        super().refresh()

    def check_input_size(self, h, w):
        m = 2 ** self.k
        if h % m or w % m:
            raise ValueError(f"input size {h}x{w}: height and width must be multiples of {m}")

    def _build(self):
        if self._plan is None:
            fe, rf = self.feature_extraction, self.edge_aware_refinements[0]
            bf = lambda t: t.detach().to(torch.bfloat16)
            cl = lambda conv: bf(conv.weight).contiguous(memory_format=torch.channels_last)
            p = {"stem": [(cl(c), c.bias.detach().float().contiguous()) for c in fe.downsample],
                 "enc_blocks": [_pack_block(b) for b in fe.residual_blocks],
                 "enc_out": (ops.pack_conv2d_tc_weight(fe.conv_alone.weight.detach().float()), fe.conv_alone.bias.detach().float().contiguous()),
                 "filter": [], "ref_blocks": [_pack_block(b) for b in rf.residual_astrous_blocks]}
            for seq in self.filter:
                conv, bn = seq[0][0], seq[0][1]
                p["filter"].append((TCConv3d(conv.weight, KIND_3x3x3), fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)))
            p["head"] = (TCConv3d(self.conv3d_alone.weight, KIND_3x3x3), self.conv3d_alone.bias.detach().float().contiguous())
            conv, bn = rf.conv2d_feature[0][0], rf.conv2d_feature[0][1]
            sc, sh = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
            p["ref_in"] = (bf(conv.weight.detach().float() * sc.view(-1, 1, 1, 1)).contiguous(memory_format=torch.channels_last), sh)
            p["ref_out"] = (ops.pack_conv2d_tc_weight(rf.conv2d_out.weight.detach().float()), rf.conv2d_out.bias.detach().float().contiguous())
            self._plan = p
        return self._plan

    def _encode(self, img, p):
        """[N,3,H,W] -> [N,h,w,32] bf16 channels-last."""
        x = img.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        for w, b in p["stem"]:                                           # 5x5 stride-2 convs with bias, no activation (modules.py:37-46)
            x = ops.bias_act(F.conv2d(x, w, None, 2, 2), b, 1.0)         # (cuDNN's own bias is a non-vectorised broadcast add in ATen)
        x = x.permute(0, 2, 3, 1)
        x = x if x.is_contiguous() else x.contiguous()
        for pk in p["enc_blocks"]:
            x = _run_block(x, pk)
        wp, b = p["enc_out"]
        return ops.conv2d_tc(x, wp, 32, 1, None, b)

    def _refine(self, disp, rgb, p):
        """EdgeAwareRefinement.forward: disp [B,h,w] fp32, rgb [B,3,H,W] -> [B,H,W] fp32."""
        up = F.interpolate(disp.unsqueeze(1), size=rgb.shape[-2:], mode="bilinear", align_corners=False)
        if rgb.shape[-1] / disp.shape[-1] >= 1.5:
            up = up * 8
        x = torch.cat([up, rgb.float()], 1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w, sh = p["ref_in"]
        x = ops.bias_act(F.conv2d(x, w, None, 1, 1), sh, 0.2).permute(0, 2, 3, 1)
        x = x if x.is_contiguous() else x.contiguous()
        for pk in p["ref_blocks"]:
            x = _run_block(x, pk)
        wp, b = p["ref_out"]
        res = ops.conv2d_tc(x, wp, 1, 1, None, b)[..., 0].float()       # [B,H,W]; channels 1..7 of the 16-byte piece are zeros
        return torch.relu(up.squeeze(1) + res)

    def _forward_train(self, batch, ref_img, tgt_img):
        from .train_ops import ConvBNAct, CostVolumeFn, HeadConv, LayerCfg
        cl = torch.channels_last
        if not self.__dict__.get("_enc_channels_last", False):
            self.feature_extraction.to(memory_format=cl)
            self.edge_aware_refinements.to(memory_format=cl)
            self.__dict__["_enc_channels_last"] = True
        to_cl = lambda t: t.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.encoder_autocast):      # one encoder call per view (:84-98)
            ref_fea = self.feature_extraction(ref_img.float().contiguous(memory_format=cl))
            tgt_fea = self.feature_extraction(tgt_img.float().contiguous(memory_format=cl))
        x = CostVolumeFn.apply(to_cl(ref_fea), to_cl(tgt_fea), self.shifts, "diff", 0)
        for seq in self.filter:
            conv, bn = seq[0][0], seq[0][1]
            x = ConvBNAct.apply(x, conv.weight, bn.weight, bn.bias, None, LayerCfg(KIND_3x3x3, True, bn, 0.2))
        cost = HeadConv.apply(x, self.conv3d_alone.weight, None).squeeze(-1) + self.conv3d_alone.bias.float().view(1, 1, 1, 1)
        prob = F.softmax(cost, dim=1)
        bins = torch.arange(self.level, device=cost.device, dtype=torch.float32) * ((self.maxdisp - self.mindisp) / float(self.level)) + self.mindisp
        disp = (prob * bins.view(1, -1, 1, 1)).sum(1)
        right = batch["right"].float()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.encoder_autocast):
            refined = self.edge_aware_refinements[0](disp, right.contiguous(memory_format=cl))
        coarse = F.interpolate((disp * (right.shape[-1] / disp.shape[-1])).unsqueeze(1), size=right.shape[-2:], mode="bilinear",
                               align_corners=False).squeeze(1)
        results = {"pred_depth": torch.stack([coarse, refined.float()], 1), "prob_depth": prob.unsqueeze(1), "pred_normal": None,
                   "ref_feature": ref_fea.detach().amax(1).float()}
        if "disp" in batch:
            results.update(self.loss_model.forward(results, batch))
        return results

    def forward(self, batch):
        if not batch["left"].is_cuda:
            raise RuntimeError("the sm_100a hot path needs CUDA tensors; there is no CPU implementation")
        self.check_input_size(*batch["left"].shape[-2:])
        ref_img, tgt_img = self._select_views(batch)
        if self.training:
            return self._forward_train(batch, ref_img, tgt_img)
        p = self._build()
        b = ref_img.shape[0]
        self._mark("start")
        f = self._encode(torch.cat([ref_img, tgt_img], 0), p)
        ref_fea, tgt_fea = f[:b], f[b:]
        self._mark("encoder")
        x = ops.costvol_fwd(ref_fea, tgt_fea, self.shifts, "diff")      # [B,D,h,w,32]
        self._mark("cost_volume")
        for conv, (sc, sh) in p["filter"]:
            x = conv(x, sc, sh, relu=True, slope=0.2)
        head, bias = p["head"]
        cost = head(x, shift=bias, out_f32=True).squeeze(-1)            # [B,D,h,w] fp32
        self._mark("aggregation")
        disp, prob = ops.softargmin(cost, float(self.mindisp), (self.maxdisp - self.mindisp) / float(self.level), self.want_prob)
        self._mark("regression")
        right = batch["right"]
        refined = self._refine(disp, right, p)
        coarse = F.interpolate((disp * (right.shape[-1] / disp.shape[-1])).unsqueeze(1), size=right.shape[-2:], mode="bilinear",
                               align_corners=False).squeeze(1)
        self._mark("refinement")
        return {"pred_depth": torch.stack([coarse, refined], 1), "prob_depth": prob.unsqueeze(1) if prob is not None else None,
                "pred_normal": None, "ref_feature": ops.channel_max(ref_fea.contiguous())}
