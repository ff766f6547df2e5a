"""NNET with the reference's model-class contract on the sm_100a kernels.

Reference: src/model/nnet/mainmodel.py:31-177 (model), src/model/nnet/modules.py (encoder, CostVolume, disp_regression),
src/model/nnet/normal_module_.py (NormalModule).  Sub-module names are the reference's, so the state_dict layout is identical
(473 entries, pinned by tests/golden/state_keys_nnet.json).  What runs where, in eval mode:

  SPP encoder (modules.py:46-139; PSMNet's, branches up-sampled with align_corners=False)   the PSMNet plan: BN folded, 39 of its
                                                                                           3x3 convs on dpf_conv2d_tc_fwd
  concat volume over int(costrange) row shifts (modules.py:169-188)                         dpf_costvol_fwd, mode "concat"
  dres0-4 + classify: 11 convbn_3d (+ReLU / +residual), 32 -> 1 head (mainmodel.py:58-85)   dpf_conv3d_fwd (kd-fused tcgen05 kernel),
                                                                                           dpf_conv3d_head_fwd
  x4 trilinear (align_corners=False) + softmax + soft-argmin, both volumes (:149-152)       dpf_regress_fwd_halfpixel (never materialises
                                                                                           the [B,32,H,W] tensor)
  context refinement `convs` on cat(ref features, cost slice), 8 levels batched (:142-146)  33 -> 128 -> 128 -> 128 -> 96: cuDNN bf16 +
                                                                                           dpf_bias_act; 96 -> 64 -> 32 -> 1: dpf_conv2d_tc_fwd
  NormalModule: 67 -> 32 and 32 -> 32 convbn_3d, three depth-halving (2,3,3) convbn_3d,     dpf_conv3d_fwd (three 32-channel input windows),
  seven dilated `n_convs` (normal_module_.py:20-42,84-118)                                  dpf_conv2d_tc_fwd (depth pairs folded into channels;
                                                                                           LeakyReLU 0.1, dilation up to 16)

Training: cost volume, dres0-4, classify and the head run forward AND backward on the kernels through the autograd Functions of
train_ops.py (CostVolumeFn, ConvBNAct, HeadConv); the 2-D encoder, the context refinement, the NormalModule and the half-pixel
up-sampling + softmax go through PyTorch autograd (bf16 autocast over cuDNN), as the encoders of the other models do.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import modules as M
from . import ops
from .layers import KIND_3x3x3, TCConv3d, fold_bn
from .models import _StereoBase
from .modules import _cb3, cost_range

CONTEXT_DILATIONS = (1, 2, 4, 8, 16, 1, 1)       # `convs` (mainmodel.py:48-56) and `n_convs` (normal_module_.py:34-42)


def _convtext(cin, cout, dil):
    """convtext(), src/model/nnet/modules.py:37-43: bias-free 3x3 conv, padding = dilation, LeakyReLU(0.1)."""
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 1, dil, dil, bias=False), nn.LeakyReLU(0.1, inplace=True))


def _context_stack(widths):
    return nn.Sequential(*[_convtext(widths[i], widths[i + 1], d) for i, d in enumerate(CONTEXT_DILATIONS)])


class NNetFeatureExtraction(M.PSMFeatureExtraction):
    """src/model/nnet/modules.py:46-139: PSMNet's SPP encoder; the pooled branches are up-sampled with align_corners=False (:115-124)."""
    branch_align_corners = False


class CostVolume(nn.Module):
    """src/model/nnet/modules.py:142-188: concat volume over int(costrange) row shifts -> [B,D,h,w,2C] bf16."""

    def __init__(self, option, mindisp, maxdisp):
        super().__init__()
        self.level = int(option.model.level)
        self.costrange = cost_range(mindisp, maxdisp, self.level)
        self.shifts = [int(d) for d in self.costrange]

    def forward(self, ref_feat, tar_feat):
        if ref_feat.requires_grad or tar_feat.requires_grad:
            from .train_ops import CostVolumeFn
            return CostVolumeFn.apply(ref_feat, tar_feat, self.shifts, "concat", 0)
        return ops.costvol_fwd(ref_feat, tar_feat, self.shifts, "concat")


def _pool3(c):
    """convbn_3d(c, c, (2,3,3), (2,1,1), (0,1,1)) + ReLU, normal_module_.py:26-31."""
    return nn.Sequential(nn.Sequential(nn.Conv3d(c, c, (2, 3, 3), (2, 1, 1), (0, 1, 1), bias=False), nn.BatchNorm3d(c)), nn.ReLU(inplace=True))


class NormalModule(nn.Module):
    """src/model/nnet/normal_module_.py:14-118 (parameter container + the PyTorch training path; eval runs NNET._normal)."""

    def __init__(self, option, mindisp, maxdisp):
        super().__init__()
        c = option.model.inplanes
        self.wc0 = nn.Sequential(_cb3(2 * c + 3, c), nn.ReLU(inplace=True), _cb3(c, c), nn.ReLU(inplace=True))
        self.pool1, self.pool2, self.pool3 = _pool3(c), _pool3(c), _pool3(c)
        self.n_convs = _context_stack((c, 3 * c, 3 * c, 3 * c, 2 * c, 2 * c, c, 3))
        level = int(option.model.level)
        cr = torch.arange(level) * ((maxdisp / 4.0 - mindisp / 4.0) / float(level)) + mindisp / 4.0
        self.costrange = nn.Parameter(cr.view(1, -1, 1, 1), False)
        self._kinv_cache = None

    def _kinv(self, K):
        """(K with its first two rows / 4)^-1 (:67-70); one small inverse per distinct K (cached: the calibration is constant per camera)."""
        c = self._kinv_cache
        if c is not None and c[0].shape == K.shape and c[0].device == K.device and torch.equal(c[0], K):
            return c[1]
        kq = K.detach().float().clone()
        kq[:, :2, :] = kq[:, :2, :] / 4.0
        inv = torch.linalg.inv_ex(kq, check_errors=False)[0]
        self._kinv_cache = (K.detach().clone(), inv)
        return inv

    def coord_volume(self, K, abvalue, h, w):
        """grid_maker_3d (:46-82) for the 8 cost levels themselves: [B,3,D,h,w] fp32, per-sample min/max normalised."""
        b = K.shape[0]
        dev = K.device
        ys, xs = torch.meshgrid(torch.arange(h, device=dev, dtype=torch.float32), torch.arange(w, device=dev, dtype=torch.float32), indexing="ij")
        grid = torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(1, 3, -1).expand(b, -1, -1)
        rays = torch.bmm(self._kinv(K), grid).view(b, 3, 1, h, w)
        ab = abvalue.float()
        depth = ab[:, 1].view(-1, 1) / (self.costrange.detach().float().view(1, -1) - ab[:, 0].view(-1, 1))     # a / (d - b), geometry.py:35-40
        depth = torch.where(torch.isnan(depth) | torch.isinf(depth), torch.zeros_like(depth), depth)
        vol = rays * depth.view(b, 1, -1, 1, 1)
        vmin = vol.reshape(b, -1).amin(-1).view(b, 1, 1, 1, 1)
        vmax = vol.reshape(b, -1).amax(-1).view(b, 1, 1, 1, 1)
        return (vol - vmin) / (vmax - vmin + 1e-6)

    def forward(self, cost_in, batch):
        """Training path: cost_in [B,2C,D,h,w] (any strides) -> [B,3,H,W]."""
        b, _, d, h, w = cost_in.shape
        wc = self.coord_volume(batch["K"], batch["abvalue"], h, w).to(cost_in.dtype)
        x = self.pool3(self.pool2(self.pool1(self.wc0(torch.cat([wc, cost_in], 1)))))
        nmap = sum(self.n_convs(x[:, :, i]) for i in range(x.shape[2]))
        nmap = F.interpolate(nmap.float(), scale_factor=4, mode="bilinear", align_corners=True)
        return F.normalize(nmap, dim=1)


class NNET(_StereoBase):
    train_supported = True
    min_quarter_size = 64                       # the encoder's 64 x 64 average-pool branch (modules.py:67-69)
    predict_normal = True

    def __init__(self, option):
        super().__init__()
        self._common_init(option)
        c = option.model.inplanes
        if c != 32 or int(option.model.level) != 8:
            raise NotImplementedError(f"NNET on the sm_100a kernels is built for inplanes = 32 and level = 8 (the reference's shipped "
                                      f"setting, src/model/nnet/config.json), got inplanes = {c}, level = {option.model.level}")
        self.predict_normal = bool(option.model.predict_normal)
        self.feature_extraction = NNetFeatureExtraction(option)
        self.cost_volume = CostVolume(option, self.mindisp, self.maxdisp)
        self.convs = _context_stack((c + 1, 4 * c, 4 * c, 4 * c, 3 * c, 2 * c, c, 1))
        self.dres0 = nn.Sequential(_cb3(2 * c, c), nn.ReLU(inplace=True), _cb3(c, c), nn.ReLU(inplace=True))
        for k in (1, 2, 3, 4):
            setattr(self, f"dres{k}", nn.Sequential(_cb3(c, c), nn.ReLU(inplace=True), _cb3(c, c)))
        self.classify = nn.Sequential(_cb3(c, c), nn.ReLU(inplace=True), nn.Conv3d(c, 1, 3, 1, 1, bias=False))
        self.normal_module = NormalModule(option, self.mindisp, self.maxdisp) if self.predict_normal else None
        self.step = (self.maxdisp - self.mindisp) / float(4 * self.level)
        self.want_prob = False                  # prob_depth [B,2,32,H,W] is consumed by no loss or metric; materialised on request only
        self._plan = None
        self.reference_init()

    def refresh(self):
        self._plan = None
        super().refresh()

    # ---- eval plan: folded BatchNorm + packed weights ---------------------------------------------------------------------
    @staticmethod
    def _l3(seq, weight=None, cin_pad=None):
        conv, bn = seq[0], seq[1]
        return (TCConv3d(conv.weight if weight is None else weight, KIND_3x3x3, cin_pad=cin_pad),
                fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps))

    def _build(self):
        if self._plan is not None:
            return self._plan
        p = {"dres0.0": self._l3(self.dres0[0]), "dres0.2": self._l3(self.dres0[2]), "classify.0": self._l3(self.classify[0]),
             "classify.2": TCConv3d(self.classify[2].weight, KIND_3x3x3)}
        for k in (1, 2, 3, 4):
            seq = getattr(self, f"dres{k}")
            p[f"dres{k}.0"], p[f"dres{k}.2"] = self._l3(seq[0]), self._l3(seq[2])
        # context refinement: the four wide layers (33 -> 128 -> 128 -> 128 -> 96) on cuDNN, the rest on the 2-D tcgen05 kernel
        cud, tc = [], []
        for i, dil in enumerate(CONTEXT_DILATIONS):
            w = self.convs[i][0].weight.detach().float()
            if i < 4:
                if i == 0:
                    w = F.pad(w, (0, 0, 0, 0, 0, 40 - w.shape[1]))            # 33 -> 40 input channels (16-byte pixels)
                cud.append((w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last), dil))
            else:
                tc.append((ops.pack_conv2d_tc_weight(w), int(w.shape[0]), dil))
        p["ctx_cudnn"], p["ctx_tc"] = cud, tc
        if self.predict_normal:
            nm = self.normal_module
            w = nm.wc0[0][0].weight.detach()
            c2 = w.shape[1] - 3
            # input channels reordered to [cost_in (2C) | coordinates (3) | zero pad] so that the 32-channel windows stay aligned
            p["wc0.0"] = self._l3(nm.wc0[0], weight=torch.cat([w[:, 3:], w[:, :3]], 1), cin_pad=(c2 + 3 + 31) // 32 * 32)
            p["wc0.2"] = self._l3(nm.wc0[2])
            pools = []
            for pool in (nm.pool1, nm.pool2, nm.pool3):
                conv, bn = pool[0][0], pool[0][1]
                w2 = conv.weight.detach().float().permute(0, 2, 1, 3, 4).reshape(conv.out_channels, 2 * conv.in_channels, 3, 3)   # channel = kd * C + c
                pools.append((ops.pack_conv2d_tc_weight(w2), conv.out_channels, fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)))
            p["pools"] = pools
            p["n_convs"] = [(ops.pack_conv2d_tc_weight(nm.n_convs[i][0].weight.detach().float()), int(nm.n_convs[i][0].out_channels), dil)
                            for i, dil in enumerate(CONTEXT_DILATIONS)]
        self._plan = p
        return p

    def _run(self, name, x, residual=None, relu=True):
        conv, (sc, sh) = self._plan[name]
        return conv(x, sc, sh, residual=residual, relu=relu)

    def _context(self, ref_fea, costs, p):
        """`convs` on cat(ref features, cost slice) for the D levels at once (mainmodel.py:142-146): costs [B,D,h,w] fp32 -> refined."""
        b, d, h, w = costs.shape
        x = torch.zeros(b, d, h, w, 40, device=costs.device, dtype=torch.bfloat16)
        x[..., :32] = ref_fea.unsqueeze(1)
        x[..., 32] = costs
        x = x.view(b * d, h, w, 40).permute(0, 3, 1, 2)                           # NCHW-shaped view of channels-last memory
        for wt, dil in p["ctx_cudnn"]:
            x = ops.bias_act(F.conv2d(x, wt, None, 1, dil, dil), None, 0.1)
        x = x.permute(0, 2, 3, 1)
        x = x if x.is_contiguous() else x.contiguous()
        for wp, cout, dil in p["ctx_tc"]:
            x = ops.conv2d_tc(x, wp, cout, dil, relu=True, slope=0.1)
        return x[..., 0].float().view(b, d, h, w) + costs

    def _normal(self, cost_in0, c0, batch, p):
        """NormalModule.forward (normal_module_.py:84-118) on the kernels: [B,D,h,w,C] x 2 -> [B,3,H,W] fp32."""
        nm = self.normal_module
        b, d, h, w, c = c0.shape
        conv0 = p["wc0.0"][0]
        x = torch.zeros(b, d, h, w, conv0.cin, device=c0.device, dtype=torch.bfloat16)
        x[..., :c] = cost_in0
        x[..., c:2 * c] = c0
        x[..., 2 * c:2 * c + 3] = nm.coord_volume(batch["K"], batch["abvalue"], h, w).permute(0, 2, 3, 4, 1)
        x = self._run("wc0.0", x)
        x = self._run("wc0.2", x)
        for wp, cout, (sc, sh) in p["pools"]:
            bb, dd, hh, ww, cc = x.shape
            if dd % 2:
                x = x[:, :dd - 1]                                                   # stride-2 depth conv without padding drops an odd tail
            pairs = x.reshape(bb, dd // 2, 2, hh, ww, cc).permute(0, 1, 3, 4, 2, 5).reshape(bb * (dd // 2), hh, ww, 2 * cc)
            x = ops.conv2d_tc(pairs.contiguous(), wp, cout, 1, sc, sh, relu=True).view(bb, dd // 2, hh, ww, cout)
        nmap = None
        for i in range(x.shape[1]):
            y = x[:, i].contiguous()
            for wp, cout, dil in p["n_convs"]:
                y = ops.conv2d_tc(y, wp, cout, dil, relu=True, slope=0.1)
            y = y[..., :3].float()
            nmap = y if nmap is None else nmap + y
        nmap = F.interpolate(nmap.permute(0, 3, 1, 2), scale_factor=4, mode="bilinear", align_corners=True)
        return F.normalize(nmap, dim=1)

    # ---- training ---------------------------------------------------------------------------------------------------------
    def _forward_train(self, batch, ref_img, tgt_img):
        from .train_ops import ConvBNAct, HeadConv, LayerCfg
        ref_fea, tgt_fea = self._features(ref_img, tgt_img)                        # one encoder call per view (mainmodel.py:116-130)
        vol = self.cost_volume(ref_fea, tgt_fea)

        def tl(seq, x, residual=None, relu=True):
            conv, bn = seq[0], seq[1]
            return ConvBNAct.apply(x, conv.weight, bn.weight, bn.bias, residual, LayerCfg(KIND_3x3x3, relu, bn))

        c0 = tl(self.dres0[2], tl(self.dres0[0], vol))
        cost_in0 = c0
        for k in (1, 2, 3, 4):
            seq = getattr(self, f"dres{k}")
            c0 = tl(seq[2], tl(seq[0], c0), residual=c0, relu=False)
        costs = HeadConv.apply(tl(self.classify[0], c0), self.classify[2].weight, None).squeeze(-1)          # [B,D,h,w] fp32
        b, d, h, w = costs.shape
        cl = torch.channels_last
        if not self.__dict__.get("_ctx_channels_last", False):
            self.convs.to(memory_format=cl)
            if self.normal_module is not None:
                self.normal_module.n_convs.to(memory_format=cl)
            self.__dict__["_ctx_channels_last"] = True
        rf = ref_fea.permute(0, 3, 1, 2).float()                                   # [B,C,h,w]
        x = torch.cat([rf.unsqueeze(1).expand(-1, d, -1, -1, -1), costs.unsqueeze(2)], 2).reshape(b * d, -1, h, w)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.encoder_autocast):
            ctx = self.convs(x.contiguous(memory_format=cl))
        refined = ctx.float().reshape(b, d, h, w) + costs
        bins = torch.arange(4 * self.level, device=costs.device, dtype=torch.float32) * self.step + self.mindisp
        disps, probs = [], []
        for c in (costs, refined):
            up = F.interpolate(c.unsqueeze(1), scale_factor=4, mode="trilinear", align_corners=False).squeeze(1)
            prob = F.softmax(up, dim=1)
            disps.append((prob * bins.view(1, -1, 1, 1)).sum(1))
            probs.append(prob if self.want_prob else None)
        normal = None
        if self.predict_normal:
            cost_in = torch.cat([cost_in0, c0], -1).permute(0, 4, 1, 2, 3)       # [B,2C,D,h,w] view of the channels-last volume
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.encoder_autocast):
                normal = self.normal_module(cost_in.float() if not self.encoder_autocast else cost_in, batch).unsqueeze(1)
        results = {"pred_depth": torch.stack(disps, 1), "prob_depth": torch.stack(probs, 1) if probs[0] is not None else None,
                   "pred_normal": normal, "ref_feature": ref_fea.detach().amax(-1).float()}
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
        self._mark("start")
        ref_fea, tgt_fea = self._features(ref_img, tgt_img)
        self._mark("encoder")
        vol = self.cost_volume(ref_fea, tgt_fea)
        self._mark("cost_volume")
        c0 = self._run("dres0.2", self._run("dres0.0", vol))
        cost_in0 = c0
        for k in (1, 2, 3, 4):
            c0 = self._run(f"dres{k}.2", self._run(f"dres{k}.0", c0), residual=c0, relu=False)
        costs = p["classify.2"](self._run("classify.0", c0), relu=False, out_f32=True).squeeze(-1)           # [B,D,h,w] fp32
        self._mark("aggregation")
        refined = self._context(ref_fea, costs, p)
        self._mark("context_refinement")
        disps, probs = zip(*[ops.regress_fwd(c.contiguous(), self.mindisp, self.step, self.want_prob, align_corners=False)
                             for c in (costs, refined)])
        self._mark("regression")
        normal = None
        if self.predict_normal:
            normal = self._normal(cost_in0, c0, batch, p).unsqueeze(1)
            self._mark("normal_branch")
        return {"pred_depth": torch.stack(disps, 1), "prob_depth": torch.stack(probs, 1) if probs[0] is not None else None,
                "pred_normal": normal, "ref_feature": ops.channel_max(ref_fea.contiguous())}
