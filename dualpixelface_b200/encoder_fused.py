"""Eval-mode execution plan of the StereoDPNet 2-D encoder (adjacent to the hot path).

The convolutions stay cuDNN (bf16, channels-last) but every BatchNorm2d is folded into its convolution and each
"+bias (+skip) -> PReLU/ReLU" tail runs as ONE pass of ``dpf_bias_act`` instead of the aten::add_ / aten::add /
aten::prelu kernels PyTorch would launch; the three dilated DPBlock branches are written straight into the channel
windows of one 96-channel buffer (no torch.cat).  Mirrors feature_extraction.forward / DPBlock.forward of the reference
(src/model/stereodpnet/modules.py:21-134) -- parameters are read from the module that owns them, nothing is registered.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.fusion import fuse_conv_bn_weights

from . import ops


# Row tiles (BASELINE config 5, tiled.py): when TILING is a tiled.RowTiling, every convolution of the plan runs on
# cat(top halo, this rank's rows, bottom halo) with the zero padding along H switched off (tiled.conv_halo); the row-streamed
# tcgen05 layers fall back to cuDNN + dpf_bias_act there (their kernel pads H itself), and the pyramid upsampling uses the
# global align_corners coordinates.  None = the untiled plan.
TILING = None
CROP = None       # (quarter-res rows of the whole image, first quarter-res row of this crop): overlap-recompute encoder of tiled.py


def _halo_input(x, f):
    """NCHW view of channels-last memory -> (input with halo rows, padding to pass to cuDNN)."""
    if TILING is None or f["w"].shape[2] == 1:
        return x, f["pad"]
    from .tiled import conv_halo
    top, bottom = conv_halo(f["w"].shape[2], f["stride"][0], f["dil"][0], f["pad"][0])
    return TILING.halo_cat(x, top, bottom, 2), (0, f["pad"][1])


# conv3 of a DPBlock as three chained 32-channel windows works (tests/test_gpu_kernels.py::test_conv2d_rows_channel_windows_chain)
# but measured slower than cuDNN + one dpf_bias_act pass (encoder 5.37 -> 5.70 ms): off.
CHAIN_CONV3 = False
import os
TC_CONV3 = os.environ.get("DPF_ENC_TC_CONV3", "1") != "0"
STEM_KERNEL = os.environ.get("DPF_ENC_STEM", "1") != "0"
TC_MODE = os.environ.get("DPF_ENC_TC", "all")          # none | dil (dilated layers only) | all


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d | None):
    """(weight bf16 channels_last, bias fp32 | None, conv hyper-parameters)."""
    w, b = conv.weight, conv.bias
    if bn is not None:
        w, b = fuse_conv_bn_weights(w, b, bn.running_mean, bn.running_var, bn.eps, bn.weight, bn.bias)
    wf = w.detach().float()
    w = wf.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = b.detach().float().contiguous() if b is not None else None
    f = dict(w=w, b=b, stride=conv.stride, pad=conv.padding, dil=conv.dilation, groups=conv.groups)
    # 3x3 / stride 1 / dilation 1, 3 or 5 / 32 -> <=32 channels: the row-streamed 2-D mode of the kd-fused tcgen05 kernel
    # (dpf_conv2d_fwd) with the bias + activation (+ skip) tail fused into its epilogue -- no cuDNN launch, no extra pass
    if (tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (1, 1) and tuple(conv.dilation) in ((1, 1), (3, 3), (5, 5)) and
            tuple(conv.padding) == tuple(conv.dilation) and conv.groups == 1 and conv.in_channels == 32 and conv.out_channels in (16, 32)):
        # (64-channel layers and multi-chunk outputs are supported by the kernel but measured slower than cuDNN here)
        f["wp"] = ops.conv2d_rows_plan(wf)
    # other 3x3 / stride 1 layers with 32, 64 or 96 input channels (any dilation: the dilation-2 conv4 of the inter-blocks, for which
    # cuDNN falls back to sm80-class kernels; the 64-channel blocks): the dedicated 2-D tcgen05 kernel, bias + activation (+ skip) fused
    elif (TC_MODE != "none" and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (1, 1) and conv.dilation[0] == conv.dilation[1]
          and tuple(conv.padding) == tuple(conv.dilation) and conv.groups == 1 and conv.in_channels in (32, 64, 96)
          and conv.out_channels % 8 == 0 and conv.out_channels <= 96 and (TC_MODE == "all" or conv.dilation[0] > 1)):
        f["tc"] = ops.pack_conv2d_tc_weight(wf)
        f["cout"] = conv.out_channels
    return f


def _conv(x, f):
    x, pad = _halo_input(x, f)
    return F.conv2d(x, f["w"], None, f["stride"], pad, f["dil"], f["groups"])


def _conv_act(x, f, slope, res=None, out=None, y_coff=0):
    """conv + bias (+ res) + activation on an NCHW view of channels-last memory: one dpf_conv2d_fwd launch when the layer is
    eligible, else cuDNN + one dpf_bias_act pass.  `out` (NHWC buffer) / y_coff select a channel window of a wider tensor."""
    if "wp" in f and TILING is None:
        xh = x.permute(0, 2, 3, 1)
        rh = res.permute(0, 2, 3, 1) if res is not None else None
        y = ops.conv2d_rows_multi(xh if xh.is_contiguous() else xh.contiguous(), f["wp"], f["b"],
                                  rh if rh is None or rh.is_contiguous() else rh.contiguous(), relu=slope != 1.0, slope=slope,
                                  out=out, y_coff=y_coff, dil=f["dil"][0])
        return y.permute(0, 3, 1, 2)
    if "tc" in f and TILING is None:
        xh = x.permute(0, 2, 3, 1)
        rh = res.permute(0, 2, 3, 1) if res is not None else None
        y = ops.conv2d_tc(xh if xh.is_contiguous() else xh.contiguous(), f["tc"], f["cout"], f["dil"][0], None, f["b"],
                          rh if rh is None or rh.is_contiguous() else rh.contiguous(), relu=slope != 1.0, slope=slope, out=out, y_coff=y_coff)
        return y.permute(0, 3, 1, 2)
    return ops.bias_act(_conv(x, f), f["b"], slope, res=res, out=out, y_coff=y_coff)


def _conv_bias_relu(x, f):
    """conv + bias + ReLU as ONE cuDNN fused op (no separate tail pass); used where the activation is a plain ReLU."""
    if "b16" not in f:
        f["b16"] = f["b"].to(torch.bfloat16)
    x, pad = _halo_input(x, f)
    return torch.cudnn_convolution_relu(x, f["w"], f["b16"], f["stride"], pad, f["dil"], f["groups"])


def pyramid_cat(f1, f2, f3, hglob=None, row0=0):
    """cat([f1, bilinear x2 (f2), bilinear x4 (f3)], 1) with align_corners=True (modules.py:128-133 of the reference); inputs
    and output are NCHW views of channels-last bf16 memory.  hglob / row0: the maps are row crops starting at level-1 row row0 of
    an image hglob level-1 rows tall (global coordinates)."""
    from . import _lib
    n, c, h, w = f1.shape
    a, b, d = (t.permute(0, 2, 3, 1).contiguous() for t in (f1, f2, f3))
    for t in (a, b, d):
        ops._req(t, torch.bfloat16, "pyramid level")
    out = torch.empty(n, h, w, 3 * c, device=f1.device, dtype=torch.bfloat16)
    _lib.check(ops.lib().dpf_pyramid_cat_tile(ops._p(a), ops._p(b), ops._p(d), ops._p(out), n, h, w, b.shape[1], b.shape[2], d.shape[1],
                                              d.shape[2], c, hglob or h, row0, ops._stream()), "dpf_pyramid_cat")
    return out.permute(0, 3, 1, 2)


def _slope(m) -> float:
    if isinstance(m, nn.PReLU):
        assert m.weight.numel() == 1
        return float(m.weight.detach())
    if isinstance(m, nn.ReLU):
        return 0.0
    raise TypeError(type(m))


class _Block:
    def __init__(self, blk):
        self.c1, self.s1 = _fold(blk.conv1[0][0], blk.conv1[0][1]), _slope(blk.conv1[1])
        self.c2, self.s2 = _fold(blk.conv2[0][0], blk.conv2[0][1]), _slope(blk.conv2[1])
        self.dil = [_fold(m[0], m[1]) for m in blk.conv_dilate]
        self.c3, self.s3 = _fold(blk.conv3[0], blk.conv3[1]), _slope(blk.prelu)
        # conv3 (96 -> 32 over the concatenated branches): three 32-channel input windows of the cat buffer, chained through the
        # residual input of dpf_conv2d_fwd -- bias and the skip tensor enter with the first window, the PReLU leaves with the last
        c3 = blk.conv3[0]
        if (CHAIN_CONV3 and tuple(c3.kernel_size) == (3, 3) and tuple(c3.stride) == (1, 1) and tuple(c3.dilation) == (1, 1) and
                tuple(c3.padding) == (1, 1) and c3.groups == 1 and c3.in_channels == 96 and c3.out_channels == 32):
            from torch.nn.utils.fusion import fuse_conv_bn_weights as _fuse
            bn = blk.conv3[1]
            wf, _ = _fuse(c3.weight, c3.bias, bn.running_mean, bn.running_var, bn.eps, bn.weight, bn.bias)
            self.c3_windows = [ops.pack_conv2d_weight(wf.detach().float()[:, k:k + 32]) for k in (0, 32, 64)]
        # conv3 on the dedicated 2-D tcgen05 kernel (dpf_conv2d_tc_fwd: all 96 input channels in one launch, bias + skip + PReLU fused)
        if (TC_CONV3 and not hasattr(self, "c3_windows") and tuple(c3.kernel_size) == (3, 3) and tuple(c3.stride) == (1, 1) and
                tuple(c3.dilation) == (1, 1) and tuple(c3.padding) == (1, 1) and c3.groups == 1 and c3.in_channels in (32, 64, 96)
                and c3.out_channels % 8 == 0 and c3.out_channels <= 96 and self.s3 != 1.0):
            from torch.nn.utils.fusion import fuse_conv_bn_weights as _fuse
            bn = blk.conv3[1]
            wf, _ = _fuse(c3.weight, c3.bias, bn.running_mean, bn.running_var, bn.eps, bn.weight, bn.bias)
            self.c3_tc = ops.pack_conv2d_tc_weight(wf.detach().float())
        self.c4, self.s4 = _fold(blk.conv4[0][0], blk.conv4[0][1]), _slope(blk.conv4[1])
        self.dw = _fold(blk.conv5.depthwise, None)
        self.pw, self.s5 = _fold(blk.conv5.pointwise, blk.conv5.bn), _slope(blk.conv5.prelu)
        self.skip = _fold(blk.conv_skip, None)

    def __call__(self, x):
        a = _conv_act(x, self.c1, self.s1)
        y = _conv_act(a, self.c2, self.s2)
        n, c, h, w = y.shape
        cat = torch.empty(n, h, w, 3 * c, device=y.device, dtype=torch.bfloat16)
        for i, f in enumerate(self.dil):
            _conv_act(y, f, 1.0, out=cat, y_coff=i * c)
        if hasattr(self, "c3_windows") and TILING is None:                                            # prelu(conv3 + a)
            ah = a.permute(0, 2, 3, 1)
            t = ops.conv2d_rows(cat, self.c3_windows[0], c, None, self.c3["b"], ah if ah.is_contiguous() else ah.contiguous(), x_coff=0)
            t = ops.conv2d_rows(cat, self.c3_windows[1], c, None, None, t, x_coff=c)
            t = ops.conv2d_rows(cat, self.c3_windows[2], c, None, None, t, relu=self.s3 != 1.0, slope=self.s3, x_coff=2 * c)
            t = t.permute(0, 3, 1, 2)
        elif hasattr(self, "c3_tc") and TILING is None:
            ah = a.permute(0, 2, 3, 1)
            t = ops.conv2d_tc(cat, self.c3_tc, c, 1, None, self.c3["b"], ah if ah.is_contiguous() else ah.contiguous(), relu=True,
                              slope=self.s3).permute(0, 3, 1, 2)
        else:
            t = ops.bias_act(_conv(cat.permute(0, 3, 1, 2), self.c3), self.c3["b"], self.s3, res=a)
        u = _conv_act(t, self.c4, self.s4)
        v = ops.bias_act(_conv(_conv(u, self.dw), self.pw), self.pw["b"], self.s5)
        return ops.bias_act(_conv(x, self.skip), self.skip["b"], 1.0, res=v)                         # + weighted skip


def fpn_merge(lateral, bias, top):
    """lateral + bias + nearest-upsampled top (one pass, dpf_fpn_merge); NCHW views of channels-last bf16 memory."""
    from . import _lib
    n, c, h, w = lateral.shape
    a, t = lateral.permute(0, 2, 3, 1).contiguous(), top.permute(0, 2, 3, 1).contiguous()
    ops._req(a, torch.bfloat16, "lateral"); ops._req(t, torch.bfloat16, "top")
    out = torch.empty_like(a)
    _lib.check(ops.lib().dpf_fpn_merge(ops._p(a), ops._p(bias), ops._p(t), ops._p(out), n, h, w, t.shape[1], t.shape[2], c,
                                       ops._stream()), "dpf_fpn_merge")
    return out.permute(0, 3, 1, 2)


class _FPN:
    """torchvision.ops.FeaturePyramidNetwork.forward (no extra blocks) over folded bf16 convs: lateral 1x1 convs, top-down
    nearest-upsample + add fused with the lateral bias (dpf_fpn_merge), 3x3 output convs."""

    def __init__(self, fpn):
        self.inner = [_fold(b[0], None) for b in fpn.inner_blocks]
        self.layer = [_fold(b[0], None) for b in fpn.layer_blocks]

    def __call__(self, feats):
        last = ops.bias_act(_conv(feats[-1], self.inner[-1]), self.inner[-1]["b"], 1.0)
        outs = [_conv_act(last, self.layer[-1], 1.0)]
        for i in range(len(feats) - 2, -1, -1):
            last = fpn_merge(_conv(feats[i], self.inner[i]), self.inner[i]["b"], last)
            outs.insert(0, _conv_act(last, self.layer[i], 1.0))
        return outs


class FusedSDPEncoder:
    def __init__(self, enc):
        fc = enc.firstconv
        self.first = [_fold(fc[i][0], fc[i][1]) for i in (0, 2, 4)]
        # 3 -> 8 zero-padded input channels (the model stages its images into an 8-channel channels-last buffer)
        self.in_channels = 8
        w0 = self.first[0]["w"]
        self.first[0]["w"] = F.pad(w0, (0, 0, 0, 0, 0, self.in_channels - w0.shape[1])).contiguous(memory_format=torch.channels_last)
        # the stem (3 -> 32, stride 2) on the repository's bandwidth-bound kernel (dpf_stem_conv_fwd) instead of cuDNN's sm80-class
        # kernel for K = 27 (0.40 -> 0.1 ms at 8 x 1120 x 1680)
        c0, bn0 = fc[0][0], fc[0][1]
        self.stem = None
        if (STEM_KERNEL and tuple(c0.kernel_size) == (3, 3) and tuple(c0.stride) == (2, 2) and tuple(c0.padding) == (1, 1)
                and tuple(c0.dilation) == (1, 1) and c0.groups == 1 and c0.out_channels == 32 and c0.in_channels <= 8):
            wf, bf = fuse_conv_bn_weights(c0.weight, c0.bias, bn0.running_mean, bn0.running_var, bn0.eps, bn0.weight, bn0.bias)
            self.stem = (ops.pack_stem_weight(wf.detach().float()), bf.detach().float().contiguous())
        self.block1 = _Block(enc.block1)
        self.inter1 = [_Block(b) for b in enc.interblock1]
        self.block2 = _Block(enc.block2)
        self.inter2 = [_Block(b) for b in enc.interblock2]
        self.block3 = _Block(enc.block3)
        self.fpn = _FPN(enc.fpn)
        self.last = [_fold(enc.lastconv[i][0], enc.lastconv[i][1]) for i in (0, 2)]

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """x [N,3,H,W] -> features [N,C,H/4,W/4] bf16, channels-last memory format."""
        x = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        if x.shape[1] < self.in_channels:
            x = F.pad(x, (0, 0, 0, 0, 0, self.in_channels - x.shape[1])).contiguous(memory_format=torch.channels_last)
        first = self.first
        if self.stem is not None and TILING is None and x.shape[1] == 8:
            xh = x.permute(0, 2, 3, 1)
            x = ops.stem_conv(xh if xh.is_contiguous() else xh.contiguous(), self.stem[0], self.stem[1], relu=True).permute(0, 3, 1, 2)
            first = self.first[1:]
        for f in first:
            x = _conv_act(x, f, 0.0) if ("wp" in f and TILING is None) else _conv_bias_relu(x, f)
        o1 = self.block1(x)
        o2 = o1
        for b in self.inter1:
            o2 = b(o2)
        o2 = self.block2(o2)
        o3 = o2
        for b in self.inter2:
            o3 = b(o3)
        o3 = self.block3(o3)
        f1, f2, f3 = self.fpn([o1, o2, o3])
        if TILING is None:
            y = pyramid_cat(f1, f2, f3, *(CROP or (None, 0)))        # upsample x2 / x4 + concat in one pass
        else:                                                        # row tile: bilinear rows from the GLOBAL coordinates (1 halo row)
            from .tiled import tiled_bilinear_rows
            t = TILING
            up2 = tiled_bilinear_rows(f2, 2, t.height // 8, t.y0 // 8, t).to(torch.bfloat16)
            up4 = tiled_bilinear_rows(f3, 4, t.height // 16, t.y0 // 16, t).to(torch.bfloat16)
            y = torch.cat([f1, up2, up4], 1).contiguous(memory_format=torch.channels_last)
        for fl in self.last:
            y = _conv_bias_relu(y, fl)
        return y
