"""Host-side mirror of the reference's stage modules for StereoDPNet and PSMNet.

Every class keeps the reference's constructor arguments, sub-module attribute names and therefore its
``state_dict`` key layout (checkpoints load unchanged), but ``forward`` runs the hot path through libdpf_sm100.so:

  CostVolume            src/model/stereodpnet/modules.py:137-200, src/model/psmnet/modules.py:174-275
  PSMNetHGAggregation   src/model/stereodpnet/modules.py:204-337 (PSMNet twin: psmnet/modules.py:279-416)
  disp_regression       src/model/stereodpnet/modules.py:341-362
  ANM                   src/model/stereodpnet/normal_module.py:32-194

The 2-D encoders (feature_extraction) are adjacent to the hot path and stay PyTorch / cuDNN.
Hot-path activations are bf16, channels-last ([B,D,H,W,C]); parameters stay fp32 nn.Parameters and are folded / packed
into kernel layout on first use (``refresh()`` after a weight update).  There is no CPU path: calling ``forward`` on CPU
tensors raises.  In train mode the same classes route through the autograd Functions of train_ops.py / train_asm.py /
train_anm.py (batch-statistics BatchNorm, backward kernels); configurations that are not built raise.
"""
from __future__ import annotations

import os

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import layers, ops, shift_tables
from .layers import KIND_1x1x1, KIND_1x3x3, KIND_3x3x3, KIND_S2, KIND_T2, TCConv3d, fold_bn


class _DenseGrad(torch.autograd.Function):
    """Identity whose backward makes the incoming gradient dense in the forward tensor's memory format.  torch.cat's backward
    hands channel slices (strided views) to its inputs, which sends the following BatchNorm backward down the slow generic
    (non channels-last) kernel: 10.5 ms per training step at config 3."""

    @staticmethod
    def forward(ctx, x):
        ctx.cl = x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.contiguous(memory_format=torch.channels_last) if ctx.cl else g.contiguous()


class EncoderBatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d of the 2-D encoders (same parameters / buffers / state-dict keys).  In TRAINING on bf16 channels-last CUDA
    activations (the autocast encoder) the batch statistics, the normalisation and the backward run on the repository's own
    bandwidth-bound kernels (dpf_channel_stats, dpf_affine_act, dpf_bn_bwd_reduce / _apply -- the train-mode BatchNorm of the 3-D
    path, which is shape-agnostic over [pixels, C]); ATen's channels-last BatchNorm kernels took 22 % of a StereoDPNet training
    step at C = 32..128.  Everything else (eval, fp32, NCHW) is nn.BatchNorm2d."""

    fused_training = True
    # size threshold below which ATen's kernels are used: with the coefficient arithmetic in one launch each way (dpf_bn_fwd_coefs /
    # dpf_bn_bwd_coefs) the fused path wins at every layer size (StereoDPNet config 3: 287 -> 267 ms, PSMNet config 4: 77 -> 69 ms)
    min_pixels = int(os.environ.get("DPF_ENC_BN_MIN_PIXELS", "0"))

    def forward(self, x):
        if (self.training and self.fused_training and x.is_cuda and x.numel() // max(x.shape[1], 1) >= self.min_pixels and x.dtype == torch.bfloat16 and x.dim() == 4 and self.affine
                and self.track_running_stats and x.shape[1] % 8 == 0 and x.shape[1] <= 256 and 256 % (x.shape[1] // 8) == 0
                and x.is_contiguous(memory_format=torch.channels_last)):
            from .train_ops import BN2dTrainFn
            return BN2dTrainFn.apply(x, self.weight, self.bias, self)
        return super().forward(x)


def _cb2(cin, cout, k, stride, pad, dil):
    """conv + BN pair with the reference's padding rule (src/module/asm/basics.py:17-22)."""
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, dil if dil > 1 else pad, dil, bias=False), EncoderBatchNorm2d(cout))


def _cb3(cin, cout, stride=1):
    return nn.Sequential(nn.Conv3d(cin, cout, 3, stride, 1, bias=False), nn.BatchNorm3d(cout))


def _tb3(cin, cout):
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False), nn.BatchNorm3d(cout))


def cost_range(mindisp, maxdisp, level) -> np.ndarray:
    return np.arange(int(level), dtype=np.float64) * ((maxdisp / 4.0 - mindisp / 4.0) / float(level)) + mindisp / 4.0


# ======================================================================================================
# 2-D encoders (PyTorch / cuDNN; adjacent to the hot path)
# ======================================================================================================
class _SepConv(nn.Module):
    """depthwise 3x3 + pointwise + BN + PReLU (reference: src/module/asm/basics.py:39-60)."""

    def __init__(self, c):
        super().__init__()
        self.depthwise = nn.Conv2d(c, c, 3, padding=1, groups=c, bias=False)
        self.pointwise = nn.Conv2d(c, c, 1, bias=False)
        self.bn = EncoderBatchNorm2d(c)
        self.prelu = nn.PReLU(init=0.05)

    def forward(self, x):
        return self.prelu(self.bn(self.pointwise(self.depthwise(x))))


class DPBlock(nn.Module):
    """src/model/stereodpnet/modules.py:21-54."""

    def __init__(self, c, ratio_s, ratio_t, reluw=0.05):
        super().__init__()
        self.conv1 = nn.Sequential(_cb2(c, c, 3, 1, 1, 1), nn.PReLU(init=reluw))
        self.conv2 = nn.Sequential(_cb2(c, c, 3, 1, 1, 1), nn.PReLU(init=reluw))
        self.conv_dilate = nn.ModuleList([_cb2(c, c, 3, 1, 2 * i + 1, 2 * i + 1) for i in range(3)])
        self.conv3 = _cb2(3 * c, c, 3, 1, 1, 1)
        self.conv4 = nn.Sequential(_cb2(c, ratio_t * c, 3, ratio_s, ratio_s, 2), nn.PReLU(init=reluw))
        self.conv5 = _SepConv(ratio_t * c)
        self.conv_skip = nn.Conv2d(c, ratio_t * c, 1, ratio_s)
        self.prelu = nn.PReLU(init=reluw)

    def forward(self, x):
        a = self.conv1(x)
        y = self.conv2(a)
        y = self.conv3(torch.cat([_DenseGrad.apply(m(y)) if self.training else m(y) for m in self.conv_dilate], 1))
        y = self.conv5(self.conv4(self.prelu(y + a)))
        return y + self.conv_skip(x)


class SDPFeatureExtraction(nn.Module):
    """feature_extraction of StereoDPNet, src/model/stereodpnet/modules.py:58-134 -> [B,32,H/4,W/4]."""

    def __init__(self, option):
        super().__init__()
        from torchvision.ops import FeaturePyramidNetwork
        c, n = option.model.inplanes, option.model.block_stack
        self.blockstack = n
        self.firstconv = nn.Sequential(_cb2(option.model.input_channel, c, 3, 2, 1, 1), nn.ReLU(inplace=True),
                                       _cb2(c, c, 3, 1, 1, 1), nn.ReLU(inplace=True),
                                       _cb2(c, c, 3, 1, 1, 1), nn.ReLU(inplace=True))
        self.block1 = DPBlock(c, 2, 1)
        self.interblock1 = nn.ModuleList([DPBlock(c, 1, 1) for _ in range(n)])
        self.block2 = DPBlock(c, 2, 2)
        self.interblock2 = nn.ModuleList([DPBlock(2 * c, 1, 1) for _ in range(n)])
        self.block3 = DPBlock(2 * c, 2, 2)
        self.fpn = FeaturePyramidNetwork([c, 2 * c, 4 * c], c, extra_blocks=None)
        self.lastconv = nn.Sequential(_cb2(3 * c, 2 * c, 3, 1, 1, 1), nn.ReLU(inplace=True),
                                      _cb2(2 * c, c, 3, 1, 1, 1), nn.ReLU(inplace=True))

    def forward(self, x):
        o1 = self.block1(self.firstconv(x))
        o2 = o1
        for m in self.interblock1:
            o2 = m(o2)
        o2 = self.block2(o2)
        o3 = o2
        for m in self.interblock2:
            o3 = m(o3)
        o3 = self.block3(o3)
        f = self.fpn(OrderedDict(layer1=o1, layer2=o2, layer3=o3))
        up = lambda t, s: F.interpolate(t, scale_factor=s, mode="bilinear", align_corners=True)
        return self.lastconv(torch.cat([f["layer1"], up(f["layer2"], 2), up(f["layer3"], 4)], 1))


class _ResBlock(nn.Module):
    """BasicBlock of the PSMNet encoder, src/model/psmnet/modules.py:14-34."""

    def __init__(self, cin, c, stride, downsample, pad, dil):
        super().__init__()
        self.conv1 = nn.Sequential(_cb2(cin, c, 3, stride, pad, dil), nn.ReLU(inplace=True))
        self.conv2 = _cb2(c, c, 3, 1, pad, dil)
        self.downsample = downsample

    def forward(self, x):
        head = self.conv2[0]
        if getattr(head, "fuses_residual", False):            # eval plan (models.route_convs_to_tc): the add rides in conv2's epilogue
            return head(self.conv1(x), residual=self.downsample(x) if self.downsample is not None else x)
        y = self.conv2(self.conv1(x))
        return y + (self.downsample(x) if self.downsample is not None else x)


class PSMFeatureExtraction(nn.Module):
    """feature_extraction of PSMNet (SPP), src/model/psmnet/modules.py:64-171 -> [B,32,H/4,W/4]."""
    branch_align_corners = True          # NNet's copy of this encoder up-samples its pooled branches with False (nnet/modules.py:115-124)

    def __init__(self, option):
        super().__init__()
        c = option.model.inplanes
        self._in = c
        self.firstconv = nn.Sequential(_cb2(3, c, 3, 2, 1, 1), nn.ReLU(inplace=True), _cb2(c, c, 3, 1, 1, 1),
                                       nn.ReLU(inplace=True), _cb2(c, c, 3, 1, 1, 1), nn.ReLU(inplace=True))
        self.layer1 = self._stack(c, 3, 1, 1, 1)
        self.layer2 = self._stack(2 * c, c // 2, 2, 1, 1)
        self.layer3 = self._stack(4 * c, 3, 1, 1, 1)
        self.layer4 = self._stack(4 * c, 3, 1, 1, 2)
        for i, k in enumerate((2 * c, c, c // 2, c // 4), start=1):
            setattr(self, f"branch{i}", nn.Sequential(nn.AvgPool2d((k, k), stride=(k, k)), _cb2(4 * c, c, 1, 1, 0, 1),
                                                       nn.ReLU(inplace=True)))
        self.lastconv = nn.Sequential(_cb2(10 * c, 4 * c, 3, 1, 1, 1), nn.ReLU(inplace=True),
                                      nn.Conv2d(4 * c, c, 1, bias=False))

    def _stack(self, planes, blocks, stride, pad, dil):
        down = None
        if stride != 1 or self._in != planes:
            down = nn.Sequential(nn.Conv2d(self._in, planes, 1, stride, bias=False), EncoderBatchNorm2d(planes))
        mods = [_ResBlock(self._in, planes, stride, down, pad, dil)]
        self._in = planes
        mods += [_ResBlock(planes, planes, 1, None, pad, dil) for _ in range(1, blocks)]
        return nn.Sequential(*mods)

    def forward(self, x):
        raw = self.layer2(self.layer1(self.firstconv(x)))
        skip = self.layer4(self.layer3(raw))
        size = skip.shape[-2:]
        br = [F.interpolate(getattr(self, f"branch{i}")(skip), size=size, mode="bilinear", align_corners=self.branch_align_corners) for i in (1, 2, 3, 4)]
        return self.lastconv(torch.cat([raw, skip, br[3], br[2], br[1], br[0]], 1))


# ======================================================================================================
# cost volumes
# ======================================================================================================
class MaskingAttention(nn.Module):
    """Parameter container with the reference's names (src/module/asm/asm.py:131-156); evaluated by CostVolumeSDP."""

    def __init__(self, nin, act="sigmoid", feature_fetch=False):
        super().__init__()
        self.normalize = nn.InstanceNorm3d(nin, affine=True)
        self.mask_convs = nn.Sequential(nn.Conv3d(nin, nin, (1, 3, 3), 1, (0, 1, 1), bias=False), nn.BatchNorm3d(nin),
                                        nn.ReLU(inplace=True),
                                        nn.Sequential(nn.Conv3d(nin, nin, 1, bias=False), self.normalize))
        if act != "sigmoid" or feature_fetch:
            raise NotImplementedError("only asm_activation='sigmoid', feature_fetch=false (the shipped config) is built")
        self.activation = nn.Sigmoid()


class _ShiftLayer(nn.Module):
    """Stands in for subpixel_shift (no parameters); keeps the attribute name `shifting_layer`."""

    def __init__(self, option):
        super().__init__()
        self.modes = (bool(option.model.nearest), bool(option.model.bilinear), bool(option.model.phase))


class CostVolumeSDP(nn.Module):
    """CostVolume of StereoDPNet (ASM), src/model/stereodpnet/modules.py:137-200 -> [B,D,H4,W4,2C] bf16.

    ``cached_first_level=True`` (default) reproduces the reference as shipped: subpixel_shift.make_grid caches the grids
    of its first call (src/module/asm/asm.py:29-30,56-57), so every level is sampled with costrange[0]; all D slices
    are identical and are produced by ONE sample / attention / blend pass that writes all D slices.
    ``cached_first_level=False`` is the evidently intended behaviour -- one shift per level (-1, -0.5, ..., 2.5 rows), including
    the fractional Fourier (phase) shifts of asm.py:63-75,112-125 (shift_tables.fourier_row_shift).
    """

    def __init__(self, option, mindisp, maxdisp):
        super().__init__()
        self.level = int(option.model.level)
        self.costrange = cost_range(mindisp, maxdisp, self.level)
        self.shifting_layer = _ShiftLayer(option)
        self.attention_layer = MaskingAttention(option.model.inplanes, act=option.model.asm_activation,
                                                feature_fetch=option.model.feature_fetch)
        self.cached_first_level = True
        self._tables: Dict[tuple, dict] = {}
        self._packed = None

    def refresh(self):
        self._packed = None

    def _tab(self, h, w, disp, direction, device):
        key = (h, w, float(disp), direction, str(device))
        if key not in self._tables:
            t = shift_tables.build_tables(h, w, disp, direction, self.shifting_layer.modes)
            self._tables[key] = {k: v.to(device) for k, v in t.items()}
        return self._tables[key]

    def sample(self, feat: torch.Tensor, tab: dict, train: bool = False) -> torch.Tensor:
        """feat [B,H4,W4,C] -> the (up to) three resampled copies [B,S,H4,W4,C]: table samples (nearest, bilinear, integer phase
        roll) from dpf_asm_sample_fwd; a FRACTIONAL phase shift (cached_first_level=False only) from shift_tables.fourier_row_shift."""
        s = None
        if "ri" in tab:
            if train:
                from .train_asm import AsmSampleFn
                s = AsmSampleFn.apply(feat, tab)
            else:
                s = ops.asm_sample(feat, tab)
        if "rot" in tab:
            ph = shift_tables.fourier_row_shift(feat, tab["rot"]).unsqueeze(1)
            s = ph if s is None else torch.cat([s, ph], 1)
        return s

    def _pack(self):
        if self._packed is None:
            mc = self.attention_layer.mask_convs
            bn = mc[1]
            self._packed = dict(conv1=TCConv3d(mc[0].weight, KIND_1x3x3), conv2=TCConv3d(mc[3][0].weight, KIND_1x1x1),
                                bn=fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps))
        return self._packed

    def _attend(self, samples: torch.Tensor):
        """samples [N,S,H,W,C] -> (logits [N,S,H,W,C] bf16, IN affine a,d [N,C])."""
        pk = self._pack()
        m = pk["conv1"](samples, pk["bn"][0], pk["bn"][1], relu=True)
        logits = pk["conv2"](m)
        st = ops.channel_stats(logits)                                  # [N,C,2]
        n = float(logits.shape[1] * logits.shape[2] * logits.shape[3])
        mean = st[..., 0] / n
        var = (st[..., 1] / n - mean * mean).clamp_min(0.0)
        inorm = self.attention_layer.normalize
        a = inorm.weight.float().unsqueeze(0) / torch.sqrt(var + inorm.eps)
        d = inorm.bias.float().unsqueeze(0) - mean * a
        return logits, a.contiguous(), d.contiguous()

    def forward(self, ref_feat: torch.Tensor, tar_feat: torch.Tensor) -> torch.Tensor:
        """ref/tar [B,H4,W4,C] bf16 (channels-last)."""
        if self.training:
            from .train_asm import asm_volume_train
            return asm_volume_train(self, ref_feat, tar_feat)
        b, h, w, c = ref_feat.shape
        vol = torch.empty(b, self.level, h, w, 2 * c, device=ref_feat.device, dtype=torch.bfloat16)
        levels = [(0, self.level, self.costrange[0])] if self.cached_first_level else \
            [(i, 1, d) for i, d in enumerate(self.costrange)]
        for d0, rep, disp in levels:
            sf = self.sample(ref_feat, self._tab(h, w, disp, "forward", ref_feat.device))
            sb = self.sample(tar_feat, self._tab(h, w, disp, "backward", ref_feat.device))
            smp = torch.cat([sf, sb], 0)                                # [2B,S,H,W,C]
            logits, a, dd = self._attend(smp)
            ops.asm_blend(smp[:b], logits[:b], a[:b].contiguous(), dd[:b].contiguous(), vol, d0, rep, 0)
            ops.asm_blend(smp[b:], logits[b:], a[b:].contiguous(), dd[b:].contiguous(), vol, d0, rep, c)
        return vol


class CostVolumePSM(nn.Module):
    """CostVolume of PSMNet, src/model/psmnet/modules.py:174-275: integer shifts int(costrange) -> bf16 NDHWC volume."""

    def __init__(self, option, mindisp, maxdisp):
        super().__init__()
        self.style = option.model.cost_volume
        self.level = int(option.model.level)
        self.group_num = option.model.group_num
        self.costrange = cost_range(mindisp, maxdisp, self.level)
        self.shifts = [int(d) for d in self.costrange]                   # truncation toward zero, modules.py:229
        if self.style not in ("psmnet", "gwcnet", "difference"):
            raise NotImplementedError(f"cost volume style is not defined : {self.style}")

    def forward(self, ref_feat, tar_feat):
        if self.style in ("psmnet", "difference"):
            mode = "concat" if self.style == "psmnet" else "diff"
            if ref_feat.requires_grad or tar_feat.requires_grad:
                from .train_ops import CostVolumeFn
                return CostVolumeFn.apply(ref_feat, tar_feat, self.shifts, mode, 0)
            return ops.costvol_fwd(ref_feat, tar_feat, self.shifts, mode)
        # gwcnet (psmnet/modules.py:268-271): cat(concat volume [2C], group-wise correlation volume [G]) along the channels, zero
        # padded to the next multiple of 32 (the conv engine's input-channel window); PSMNetHGAggregation pads its first layer alike
        c, g = ref_feat.shape[-1], int(self.group_num)
        if c % g != 0:
            raise ValueError(f"group_num {g} must divide the {c} feature channels (assert of psmnet/modules.py:217; the shipped "
                             f"group_num = 40 fails that assert in the reference too)")
        if g % 8 != 0:
            raise NotImplementedError(f"group_num {g}: the correlation kernel writes 16-byte pieces (group_num must be a multiple of 8)")
        if ref_feat.requires_grad or tar_feat.requires_grad:
            from .train_ops import CostVolumeFn
            vc = CostVolumeFn.apply(ref_feat, tar_feat, self.shifts, "concat", 0)
            vg = CostVolumeFn.apply(ref_feat, tar_feat, self.shifts, "gwc", g)
        else:
            vc = ops.costvol_fwd(ref_feat, tar_feat, self.shifts, "concat")
            vg = ops.costvol_fwd(ref_feat, tar_feat, self.shifts, "gwc", g)
        pad = (-(2 * c + g)) % 32
        parts = [vc, vg] + ([vc.new_zeros(*vc.shape[:-1], pad)] if pad else [])
        return torch.cat(parts, -1)


# ======================================================================================================
# 3-D aggregation
# ======================================================================================================
class PSMNetHourglass(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Sequential(_cb3(c, 2 * c, 2), nn.ReLU(inplace=True))
        self.conv2 = _cb3(2 * c, 2 * c, 1)
        self.conv3 = nn.Sequential(_cb3(2 * c, 2 * c, 2), nn.ReLU(inplace=True))
        self.conv4 = nn.Sequential(_cb3(2 * c, 2 * c, 1), nn.ReLU(inplace=True))
        self.conv5 = _tb3(2 * c, 2 * c)
        self.conv6 = _tb3(2 * c, c)


class PSMNetHGAggregation(nn.Module):
    """22 Conv3d + 6 ConvTranspose3d on the tcgen05 engine, BN/ReLU/residual fused into each epilogue.

    Accepts the reference's two constructor forms: StereoDPNet passes the channel count
    (src/model/stereodpnet/modules.py:267), PSMNet the option object (src/model/psmnet/modules.py:342-349).
    forward(volume [B,D,H4,W4,2C] bf16) -> ([cost3(,cost2,cost1)] each [B,D,H4,W4] fp32 at QUARTER resolution,
    [out3(,out2,out1)] each [B,D,H4,W4,C] bf16); the x4 trilinear upsample is fused into disp_regression.
    """

    def __init__(self, option_or_channels):
        super().__init__()
        if isinstance(option_or_channels, int):
            c, first = option_or_channels, 2 * option_or_channels
        else:
            o = option_or_channels
            c = o.model.inplanes
            first = 2 * c if o.model.cost_volume == "psmnet" else 2 * c + o.model.group_num
        self.first_pad = (-first) % 32          # gwcnet: 2C+G input channels, zero-padded to the conv engine's 32-channel windows
        self.multiplier = 4
        self.dres0 = nn.Sequential(_cb3(first, c), nn.ReLU(inplace=True), _cb3(c, c), nn.ReLU(inplace=True))
        self.dres1 = nn.Sequential(_cb3(c, c), nn.ReLU(inplace=True), _cb3(c, c))
        self.dres2, self.dres3, self.dres4 = PSMNetHourglass(c), PSMNetHourglass(c), PSMNetHourglass(c)
        for k in (1, 2, 3):
            setattr(self, f"classif{k}", nn.Sequential(_cb3(c, c), nn.ReLU(inplace=True), nn.Conv3d(c, 1, 3, 1, 1, bias=False)))
        self._plan = None

    def refresh(self):
        self._plan = None

    def _layer(self, seq: nn.Sequential, kind, transposed=False, cin_pad=None):
        conv, bn = seq[0], seq[1]
        return TCConv3d(conv.weight, kind, transposed, cin_pad=cin_pad), fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)

    def _build(self):
        if self._plan is not None:
            return self._plan
        p = {}
        p["dres0.0"] = self._layer(self.dres0[0], KIND_3x3x3, cin_pad=self.dres0[0][0].in_channels + self.first_pad)
        p["dres0.2"] = self._layer(self.dres0[2], KIND_3x3x3)
        p["dres1.0"] = self._layer(self.dres1[0], KIND_3x3x3)
        p["dres1.2"] = self._layer(self.dres1[2], KIND_3x3x3)
        for name in ("dres2", "dres3", "dres4"):
            hg = getattr(self, name)
            p[name + ".conv1"] = self._layer(hg.conv1[0], KIND_S2)
            p[name + ".conv2"] = self._layer(hg.conv2, KIND_3x3x3)
            p[name + ".conv3"] = self._layer(hg.conv3[0], KIND_S2)
            p[name + ".conv4"] = self._layer(hg.conv4[0], KIND_3x3x3)
            p[name + ".conv5"] = self._layer(hg.conv5, KIND_T2, transposed=True)
            p[name + ".conv6"] = self._layer(hg.conv6, KIND_T2, transposed=True)
        for k in (1, 2, 3):
            cl = getattr(self, f"classif{k}")
            p[f"classif{k}.0"] = self._layer(cl[0], KIND_3x3x3)
            p[f"classif{k}.2"] = (TCConv3d(cl[2].weight, KIND_3x3x3), None)
        self._plan = p
        return p

    def _run(self, name, x, residual=None, relu=True):
        conv, (sc, sh) = self._plan[name]
        return conv(x, sc, sh, residual=residual, relu=relu)

    def _hourglass(self, name, x, presqu, postsqu, cost0):
        o = self._run(name + ".conv1", x)
        pre = self._run(name + ".conv2", o, residual=postsqu, relu=True)
        o = self._run(name + ".conv3", pre)
        o = self._run(name + ".conv4", o)
        post = self._run(name + ".conv5", o, residual=presqu if presqu is not None else pre, relu=True)
        out = self._run(name + ".conv6", post, residual=cost0, relu=False)     # "+ cost0" of modules.py:315,318,321 fused
        return out, pre, post

    # ---- training: batch-statistics BatchNorm, autograd Functions over the same kernels (train_ops.py) ----------
    def _tl(self, seq, kind, x, residual=None, relu=True):
        from .train_ops import ConvBNAct, LayerCfg
        conv, bn = seq[0], seq[1]
        w = conv.weight
        if kind != KIND_T2 and x.shape[-1] > w.shape[1]:            # zero-padded input channels (gwcnet first layer)
            w = F.pad(w, (0, 0, 0, 0, 0, 0, 0, x.shape[-1] - w.shape[1]))
        return ConvBNAct.apply(x, w, bn.weight, bn.bias, residual, LayerCfg(kind, relu, bn))

    def _hourglass_train(self, hg, x, presqu, postsqu, cost0):
        o = self._tl(hg.conv1[0], KIND_S2, x)
        pre = self._tl(hg.conv2, KIND_3x3x3, o, residual=postsqu, relu=True)
        o = self._tl(hg.conv3[0], KIND_S2, pre)
        o = self._tl(hg.conv4[0], KIND_3x3x3, o)
        post = self._tl(hg.conv5, KIND_T2, o, residual=presqu if presqu is not None else pre, relu=True)
        out = self._tl(hg.conv6, KIND_T2, post, residual=cost0, relu=False)
        return out, pre, post

    def _forward_train(self, cost):
        from .train_ops import HeadConv
        c0 = self._tl(self.dres0[0], KIND_3x3x3, cost)
        c0 = self._tl(self.dres0[2], KIND_3x3x3, c0)
        r = self._tl(self.dres1[0], KIND_3x3x3, c0)
        cost0 = self._tl(self.dres1[2], KIND_3x3x3, r, residual=c0, relu=False)
        out1, pre1, post1 = self._hourglass_train(self.dres2, cost0, None, None, cost0)
        out2, _p2, post2 = self._hourglass_train(self.dres3, out1, pre1, post1, cost0)
        out3, _p3, _q3 = self._hourglass_train(self.dres4, out2, pre1, post2, cost0)
        costs, prev = [], None
        for k, o in ((1, out1), (2, out2), (3, out3)):
            cl = getattr(self, f"classif{k}")
            prev = HeadConv.apply(self._tl(cl[0], KIND_3x3x3, o), cl[2].weight, prev)
            costs.append(prev)
        costs = [c.squeeze(-1) for c in costs]
        return [costs[2], costs[1], costs[0]], [out3, out2, out1]

    def forward(self, cost: torch.Tensor, all_heads: Optional[bool] = None):
        if self.training:
            return self._forward_train(cost)
        self._build()
        c0 = self._run("dres0.0", cost)
        c0 = self._run("dres0.2", c0)
        r = self._run("dres1.0", c0)
        cost0 = self._run("dres1.2", r, residual=c0, relu=False)
        out1, pre1, post1 = self._hourglass("dres2", cost0, None, None, cost0)
        out2, _pre2, post2 = self._hourglass("dres3", out1, pre1, post1, cost0)
        out3, _pre3, _post3 = self._hourglass("dres4", out2, pre1, post2, cost0)
        costs, prev = [], None
        for k, o in ((1, out1), (2, out2), (3, out3)):
            hfeat = self._run(f"classif{k}.0", o)
            head, _ = self._plan[f"classif{k}.2"]
            prev = head(hfeat, residual=prev, relu=False, out_f32=True)      # [B,D,H4,W4,1] fp32, cumulative adds
            costs.append(prev)
        costs = [c.squeeze(-1) for c in costs]
        if all_heads if all_heads is not None else self.training:
            return [costs[2], costs[1], costs[0]], [out3, out2, out1]
        return [costs[2]], [out3]


class disp_regression(nn.Module):
    """Fused x4 trilinear upsample + softmax + soft-argmin (src/model/stereodpnet/modules.py:327-334,341-362).

    forward(list of quarter-res costs [B,D,H4,W4] fp32) -> (disparities [B,H,W], probabilities [B,4D,H,W] or None).
    `prob_depth` is consumed by no loss or metric of the reference; it is materialised only when want_prob is set.
    """

    def __init__(self, mindisp, maxdisp, level):
        super().__init__()
        self.mindisp, self.step = float(mindisp), (maxdisp - mindisp) / float(4 * level)
        self.want_prob = False

    def forward(self, x: Sequence[torch.Tensor]):
        disps, probs = [], []
        for cost in x:
            assert cost.dim() == 4
            if cost.requires_grad:
                from .train_ops import RegressFn
                disps.append(RegressFn.apply(cost, self.mindisp, self.step))
                probs.append(None)
                continue
            d, p = ops.regress_fwd(cost.contiguous(), self.mindisp, self.step, self.want_prob)
            disps.append(d)
            probs.append(p)
        return disps, probs


# ======================================================================================================
# normal branch
# ======================================================================================================
class DeformConvPack(nn.Module):
    """Parameter container of DeformConvPack_dv2 (src/module/dcn3d/modules/deform_conv.py:295-321)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, 3, 3, 3))
        self.bias = nn.Parameter(torch.zeros(cout))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self.conv_offset = nn.Conv3d(cin, 81, 3, 1, 1, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)


def _convtext(cin, cout, dil):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 1, dil, dil, bias=False), nn.LeakyReLU(0.1, inplace=True))


class ANM(nn.Module):
    """Normal branch, src/model/stereodpnet/normal_module.py:32-194 (use_sampling, use_deform: the shipped config)."""

    def __init__(self, option, mindisp, maxdisp):
        super().__init__()
        c = option.model.inplanes
        self.use_deform, self.use_sampling = bool(option.model.use_deform), bool(option.model.use_sampling)
        # use_sampling=false (normal_module.py:159-163): all `level` cost slices with their own disparities instead of the k nearest
        # ones -- exactly what the select / gather kernels produce for k = level (top-k of all, sorted ascending = every level)
        self.k = int(option.model.dsample_num) if self.use_sampling else int(option.model.level)
        self.levels = [float(v) for v in cost_range(mindisp, maxdisp, option.model.level).astype(np.float32)]
        if self.use_deform:
            self.deform_conv1 = DeformConvPack(c + 3, 2 * c)
            self.act1 = nn.Sequential(nn.BatchNorm3d(2 * c), nn.ReLU(inplace=True))
            self.deform_conv2 = DeformConvPack(2 * c, 2 * c)
            self.act2 = nn.Sequential(nn.BatchNorm3d(2 * c), nn.ReLU(inplace=True))
        else:                                  # use_deform=false (normal_module.py:52-56): two plain convbn_3d + ReLU
            self.original_conv = nn.Sequential(_cb3(c + 3, 2 * c), nn.ReLU(inplace=True), _cb3(2 * c, 2 * c), nn.ReLU(inplace=True))
        self.n_convs = nn.Sequential(_convtext(2 * c, 3 * c, 1), _convtext(3 * c, 3 * c, 2), _convtext(3 * c, 2 * c, 4),
                                     _convtext(2 * c, 2 * c, 8), _convtext(2 * c, c, 1), _convtext(c, 3, 1))
        cr = torch.arange(option.model.level) * ((maxdisp / 4.0 - mindisp / 4.0) / float(option.model.level)) + mindisp / 4.0
        self.register_parameter("costrange", nn.Parameter(cr.view(1, -1, 1, 1), False))
        self._plan = None

    def refresh(self):
        self._plan = None

    def _build(self):
        if self._plan is None:
            p = {}
            if not self.use_deform:
                for i, seq in ((1, self.original_conv[0]), (2, self.original_conv[2])):
                    conv, bn = seq[0], seq[1]
                    p[f"oc{i}"] = (TCConv3d(conv.weight, KIND_3x3x3, cin_pad=64),
                                   fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps))
            for i, (dc, act) in enumerate(((self.deform_conv1, self.act1), (self.deform_conv2, self.act2)) if self.use_deform else (), start=1):
                bn = act[0]
                cin = dc.weight.shape[1]
                cpad = 64      # gathering 64 (zero-padded) channels measured faster than the 48-channel variant
                # 81 -> 96 zero-padded offset channels: a 384-byte voxel pitch keeps the fp32 epilogue on 128-bit stores
                p[f"off{i}"] = TCConv3d(F.pad(dc.conv_offset.weight.detach(), (0, 0, 0, 0, 0, 0, 0, 0, 0, 15)), KIND_3x3x3, cin_pad=64)
                p[f"offb{i}"] = F.pad(dc.conv_offset.bias.detach().float(), (0, 15)).contiguous()
                p[f"w{i}"] = ops.pack_conv_weight(dc.weight.detach(), cin_pad=cpad)
                p[f"cpad{i}"] = cpad
                p[f"aff{i}"] = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, conv_bias=dc.bias)
            # the six dilated 3x3 `convtext` layers (64->96->96->64->64->32->3, dilation 1,2,4,8,1,1) on the dedicated 2-D tcgen05
            # kernel (dpf_conv2d_tc_fwd): one launch per layer, LeakyReLU(0.1) fused, any dilation via residue-class sub-images
            p["nconv"] = [(ops.pack_conv2d_tc_weight(m[0].weight.detach().float()), m[0].out_channels, m[0].dilation[0])
                          for m in self.n_convs]
            self._plan = p
        return self._plan

    def _kinv(self, K: torch.Tensor) -> torch.Tensor:
        """inverse of the quarter-resolution intrinsics (normal_module.py:100-103).  torch.inverse synchronises the host (its error
        check reads back a status); linalg.inv_ex(check_errors=False) is the same factorisation without the read-back, and the result
        is cached per intrinsics tensor (storage, version, shape): steady-state inference is sync-free and CUDA-graph capturable."""
        key = (K.data_ptr(), K._version, tuple(K.shape), K.device)
        c = self.__dict__.get("_kinv_cache")
        if c is None or c[0] != key:
            kq = K.float().clone()
            kq[:, :2, :] = kq[:, :2, :] / 4.0
            c = (key, torch.linalg.inv_ex(kq, check_errors=False)[0].contiguous())
            self.__dict__["_kinv_cache"] = c
        return c[1]

    def forward(self, costs: Sequence[torch.Tensor], disp_maps: Sequence[torch.Tensor], batch: dict):
        """costs: [out3] as [B,D,H4,W4,C] bf16; disp_maps: [disparity [B,H,W] fp32] -> ([normal [B,3,H,W]], offsets, offsets)."""
        if self.training:                      # forward + backward through the autograd Functions of train_anm.py
            from .train_anm import anm_train
            outs = [anm_train(self, out3, disp, batch) for out3, disp in zip(costs, disp_maps)]
            return [o[0] for o in outs], [o[1] for o in outs], [o[2] for o in outs]
        p = self._build()
        normals, off1s, off2s = [], [], []
        for out3, disp in zip(costs, disp_maps):
            b = out3.shape[0]
            kinv = self._kinv(batch["K"])
            idx, coord, minmax = ops.anm_select(disp.contiguous(), kinv, batch["abvalue"].float().contiguous(), self.levels, self.k)
            fv = ops.anm_gather(out3, idx, coord, minmax, 64)                          # [B,K,H4,W4,64]
            if self.use_deform:
                off1 = p["off1"](fv, shift=p["offb1"], out_f32=True)
                f1 = ops.dcn3d(fv, off1, p["w1"], p["cpad1"], p["aff1"][0], p["aff1"][1], relu=True, cin_real=self.deform_conv1.weight.shape[1])
                off2 = p["off2"](f1, shift=p["offb2"], out_f32=True)
                f2 = ops.dcn3d(f1, off2, p["w2"], p["cpad2"], p["aff2"][0], p["aff2"][1], relu=True)
            else:
                off1 = off2 = None
                f1 = p["oc1"][0](fv, p["oc1"][1][0], p["oc1"][1][1], relu=True)
                f2 = p["oc2"][0](f1, p["oc2"][1][0], p["oc2"][1][1], relu=True)
            # shared 2-D normal convs on the (b*k) slices, channels-last, no layout change: [B,K,H4,W4,64] IS [B*K,H4,W4,64]
            x = f2.view(b * self.k, f2.shape[2], f2.shape[3], f2.shape[4])
            for wp, cout, dil in p["nconv"]:
                x = ops.conv2d_tc(x, wp, cout, dil, relu=True, slope=0.1)          # last layer: 3 real + 5 zero channels
            # fused x4 bilinear upsample + sigmoid + mean over the k sampled planes + rescale to [-1, 1]
            from .ops_tail import anm_tail
            normals.append(anm_tail(x, b, self.k))
            off1s.append(off1[..., :81] if off1 is not None else None)
            off2s.append(off2[..., :81] if off2 is not None else None)
        return normals, off1s, off2s
