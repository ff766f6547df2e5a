"""STEREODPNET / PSMNET with the reference's model-class contract on the sm_100a hot path.

Contract kept (src/model/stereodpnet/mainmodel.py:21-177, src/model/psmnet/mainmodel.py:31-181 of the reference):
``CLS(option)``; sub-module names ``feature_extraction, cost_volume, aggregation, normal_estimator, regression_layer``
(=> identical state_dict keys); ``forward(batch: dict) -> dict`` with ``pred_depth [B,n,H,W]``, ``prob_depth``,
``pred_normal [B,1,3,H,W] | None``, ``ref_feature [B,H4,W4]``; the Lightning-style hooks used by main.py.
pytorch_lightning is not available here, so ``runner.LightningModule`` supplies the few hooks the repo uses.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import modules as M
from . import ops
from .runner import LightningModule, optimizer_selector, scheduler_selector
from .synthetic import synthetic_batch


import os as _os
_GENERIC_TC = _os.environ.get("DPF_ENC_GENERIC_TC", "1") != "0"


class TCConv2dEval(nn.Module):
    """Eval-mode stand-in for a BN-folded nn.Conv2d (3x3, stride 1, padding = dilation) on channels-last bf16 activations:
    one dpf_conv2d_tc_fwd launch with the bias (and an immediately following ReLU) in its epilogue."""

    def __init__(self, conv: nn.Conv2d, relu: bool):
        super().__init__()
        self.weight = nn.Parameter(conv.weight.detach().clone(), requires_grad=False)
        self.bias = nn.Parameter(conv.bias.detach().clone(), requires_grad=False) if conv.bias is not None else None
        self.cout, self.dil, self.relu = conv.out_channels, conv.dilation[0], relu
        self._packed = None

    min_pixels = 100_000          # below this a launch is latency-bound and cuDNN's small-problem kernels win (PSMNet 1 x 448 x 448)
    fuses_residual = True         # forward(x, residual): the skip connection of a residual block rides in the conv epilogue

    def forward(self, x, residual=None):
        assert residual is None or not self.relu          # y = conv + bias + residual (BasicBlock: no activation after the add)
        if x.shape[0] * x.shape[2] * x.shape[3] < self.min_pixels:
            y = torch.nn.functional.conv2d(x, self.weight, self.bias, 1, self.dil, self.dil)
            if residual is not None:
                y = y + residual
            return torch.relu_(y) if self.relu else y
        if self._packed is None or self._packed[0].device != x.device:
            self._packed = (ops.pack_conv2d_tc_weight(self.weight.detach().float().to(x.device)),
                            self.bias.detach().float().to(x.device).contiguous() if self.bias is not None else None)
        xh = x.permute(0, 2, 3, 1)
        rh = None
        if residual is not None:
            rh = residual.permute(0, 2, 3, 1)
            rh = rh if rh.is_contiguous() else rh.contiguous()
        y = ops.conv2d_tc(xh if xh.is_contiguous() else xh.contiguous(), self._packed[0], self.cout, self.dil, None, self._packed[1],
                          residual=rh, relu=self.relu)
        return y.permute(0, 3, 1, 2)


def _tc_eligible(m) -> bool:
    return (isinstance(m, nn.Conv2d) and tuple(m.kernel_size) == (3, 3) and tuple(m.stride) == (1, 1) and m.groups == 1
            and m.dilation[0] == m.dilation[1] and tuple(m.padding) == tuple(m.dilation) and m.in_channels in (32, 64, 96)
            and m.out_channels % 8 == 0 and m.out_channels <= 96)


class CudnnConvBiasAct(nn.Module):
    """Eval-mode stand-in for a BN-folded nn.Conv2d that stays on cuDNN (strided, 1x1, 128-channel layers): the convolution runs
    bias-free and ONE dpf_bias_act pass applies the folded bias, the ReLU that follows and the skip connection of a residual
    block.  (cuDNN's own bias is a separate broadcast add that ATen runs on its non-vectorised element-wise kernel: 21 such passes
    were 13 % of an NNet / PSMNet encoder forward, the ReLU and the skip add two more passes each.)"""
    fuses_residual = True

    def __init__(self, conv: nn.Conv2d, relu: bool):
        super().__init__()
        self.weight = nn.Parameter(conv.weight.detach().clone(), requires_grad=False)
        self.__dict__["_bias32"] = conv.bias.detach().float().clone()          # kept in fp32 whatever dtype the module is moved to
        self.stride, self.padding, self.dilation, self.relu = conv.stride, conv.padding, conv.dilation, relu
        self.cout = conv.out_channels

    def forward(self, x, residual=None):
        assert residual is None or not self.relu          # y = conv + bias + residual (BasicBlock: no activation after the add)
        y = torch.nn.functional.conv2d(x, self.weight, None, self.stride, self.padding, self.dilation)
        b = self.__dict__["_bias32"]
        if b.device != y.device:
            b = self.__dict__["_bias32"] = b.to(y.device)
        if not (y.is_cuda and y.dtype == torch.bfloat16):
            y = y + b.to(y.dtype).view(1, -1, 1, 1)
            if residual is not None:
                y = y + residual
            return torch.relu_(y) if self.relu else y
        cl = torch.channels_last
        y = y if y.is_contiguous(memory_format=cl) else y.contiguous(memory_format=cl)
        if residual is not None and not residual.is_contiguous(memory_format=cl):
            residual = residual.contiguous(memory_format=cl)
        return ops.bias_act(y, b, 0.0 if self.relu else 1.0, res=residual)


def _stand_in(conv, relu: bool):
    """(replacement module | None, 1 if it runs on the repository's 2-D tcgen05 kernel else 0)."""
    if _tc_eligible(conv):
        return TCConv2dEval(conv, relu), 1
    if isinstance(conv, nn.Conv2d) and conv.bias is not None and conv.groups == 1 and conv.out_channels % 8 == 0:
        return CudnnConvBiasAct(conv, relu), 0
    return None, 0


def route_convs_to_tc(mod: nn.Module) -> int:
    """Replace the convolutions of an eval-mode, BN-folded encoder copy (in place): the eligible 3x3 ones by TCConv2dEval (their
    number is returned), the other biased ones by CudnnConvBiasAct.
    Pattern handled: Sequential(conv, Identity[folded BN]) optionally followed by nn.ReLU in the parent Sequential."""
    n = 0
    for _, child in list(mod.named_children()):
        if isinstance(child, nn.Sequential):
            items = list(child.named_children())
            for i, (name, sub) in enumerate(items):
                if isinstance(sub, nn.Sequential) and len(sub) >= 1 and type(sub[0]) is nn.Conv2d and all(isinstance(t, nn.Identity) for t in list(sub)[1:]):
                    relu = i + 1 < len(items) and isinstance(items[i + 1][1], nn.ReLU)
                    rep, k = _stand_in(sub[0], relu)
                    if rep is None:
                        continue
                    sub[0] = rep
                    if relu:
                        setattr(child, items[i + 1][0], nn.Identity())
                    n += k
            if len(child) >= 1 and type(child[0]) is nn.Conv2d and all(isinstance(t, nn.Identity) for t in list(child)[1:]):
                rep, k = _stand_in(child[0], False)               # a bare (conv, folded BN) pair, e.g. _ResBlock.conv2 / downsample
                if rep is not None:
                    child[0] = rep
                    n += k
        n += route_convs_to_tc(child)
    return n


class _StereoBase(LightningModule):
    predict_normal = False
    train_supported = False
    data_seed_offset = 0              # Trainer.fit sets rank * const under torchrun: every rank draws its own synthetic pairs

    def _common_init(self, option):
        from .losses import LossModel
        self.loss_model = LossModel(option)
        self.option = option
        self.mindisp = option.model.mindisp
        self.maxdisp = option.model.maxdisp
        self.level = option.model.level
        self.encoder_autocast = True          # bf16 autocast for the cuDNN encoder; tests switch it off to isolate the hot path

    # ---- the hot path -------------------------------------------------------------------------------------
    def _select_views(self, batch):
        """reference / target image selection, mainmodel.py:70-83."""
        if "groupname" in batch and not self.training:
            swap = batch["groupname"][0] == "2020-2-9_group20"
        else:
            swap = bool(self.option.dataset.flip_lr)
        return (batch["right"], batch["left"]) if swap else (batch["left"], batch["right"])

    def _fused_encoder(self):
        """Eval-mode copy of the cuDNN encoder with every BatchNorm2d folded into its convolution, bf16, channels-last
        (kept outside the module tree so that the state_dict layout is untouched; rebuilt by refresh())."""
        enc = self.__dict__.get("_enc_fused")
        if enc is None and isinstance(self.feature_extraction, M.SDPFeatureExtraction):
            from .encoder_fused import FusedSDPEncoder
            enc = FusedSDPEncoder(self.feature_extraction)      # cuDNN convs + dpf_bias_act tails, no torch.cat
            self.__dict__["_enc_fused"] = enc
        if enc is None:
            import copy
            from torch.nn.utils.fusion import fuse_conv_bn_eval
            enc = copy.deepcopy(self.feature_extraction).eval()

            def walk(mod):
                for _, child in list(mod.named_children()):
                    if isinstance(child, nn.Sequential) and len(child) >= 2 and isinstance(child[0], nn.Conv2d) \
                            and isinstance(child[1], nn.BatchNorm2d):
                        child[0] = fuse_conv_bn_eval(child[0], child[1])
                        child[1] = nn.Identity()
                    if isinstance(child, M._SepConv):
                        child.pointwise = fuse_conv_bn_eval(child.pointwise, child.bn)
                        child.bn = nn.Identity()
                    walk(child)

            walk(enc)
            # every BN-folded 3x3 / stride-1 convolution with 32, 64 or 96 input and <= 96 output channels (PSMNet: firstconv 2-3,
            # layer1, layer2 = 39 of its 3x3 convs) runs on the repository's 2-D tcgen05 kernel, a directly following ReLU fused
            if _GENERIC_TC:
                route_convs_to_tc(enc)
            enc = enc.to(device=next(self.parameters()).device, dtype=torch.bfloat16, memory_format=torch.channels_last)
            self.__dict__["_enc_fused"] = enc
        return enc

    def _features(self, ref_img, tgt_img):
        b = ref_img.shape[0]
        if self.training:
            # the reference runs the encoder once per view (mainmodel.py:72-83): BatchNorm batch statistics are per call
            # channels-last end to end: cuDNN then needs no NCHW<->NHWC transposes and BatchNorm takes the NHWC kernels
            cl = torch.channels_last
            if not self.__dict__.get("_enc_channels_last", False):
                self.feature_extraction.to(memory_format=cl)
                self.__dict__["_enc_channels_last"] = True
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.encoder_autocast):
                fr = self.feature_extraction(ref_img.float().contiguous(memory_format=cl))
                ft = self.feature_extraction(tgt_img.float().contiguous(memory_format=cl))
            to_cl = lambda t: t.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
            return to_cl(fr), to_cl(ft)
        if self.encoder_autocast:
            # both views go straight into one bf16 channels-last batch (one conversion pass each, no fp32 cat)
            # (the SDP plan takes an 8-channel, zero-padded input: cuDNN then needs no channel-padding pass of its own; the
            # buffer is cached per shape so that the pad channels are zeroed once)
            enc = self._fused_encoder()
            cpad = getattr(enc, "in_channels", ref_img.shape[1])
            key = (2 * b, cpad, *ref_img.shape[2:], ref_img.device)
            buf = self.__dict__.get("_in_buf")
            if buf is None or buf[0] != key:
                x = torch.zeros(2 * b, cpad, *ref_img.shape[2:], device=ref_img.device, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
                self.__dict__["_in_buf"] = (key, x)
            else:
                x = buf[1]
            c = ref_img.shape[1]
            x[:b, :c].copy_(ref_img)
            x[b:, :c].copy_(tgt_img)
            f = enc(x)
        else:
            f = self.feature_extraction(torch.cat([ref_img, tgt_img], 0).float())
        f = f.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()      # [2B,H4,W4,C] channels-last bf16
        return f[:b], f[b:]

    def _mark(self, name):
        """Optional per-stage CUDA events (bench.py sets self.stage_events = [] to collect them)."""
        ev = getattr(self, "stage_events", None)
        if ev is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e))

    def forward(self, batch):
        if not batch["left"].is_cuda:
            raise RuntimeError("the sm_100a hot path needs CUDA tensors; there is no CPU implementation")
        self.check_input_size(*batch["left"].shape[-2:])
        ref_img, tgt_img = self._select_views(batch)
        self._mark("start")
        ref_fea, tgt_fea = self._features(ref_img, tgt_img)
        self._mark("encoder")
        cost = self.cost_volume(ref_fea, tgt_fea)
        self._mark("cost_volume")
        cost_i, cost = self.aggregation(cost)
        self._mark("aggregation")
        cost_f, cost_p = self.regression_layer(cost_i)
        self._mark("regression")
        normal = None
        if self.predict_normal:
            normals, _, _ = self.normal_estimator([cost[0]], [cost_f[0]], batch)
            normal = torch.stack(normals, 1)                             # 'n b c h w -> b n c h w'
            self._mark("normal_branch")
        results = {"pred_depth": torch.stack(cost_f, 1),
                   "prob_depth": torch.stack(cost_p, 1) if cost_p[0] is not None else None,
                   "pred_normal": normal,
                   "ref_feature": ops.channel_max(ref_fea) if ref_fea.shape[-1] % 8 == 0 else ref_fea.amax(-1).float()}
        if self.training and "disp" in batch:
            results.update(self.loss_model.forward(results, batch))
        return results

    min_quarter_size = 1          # PSMNet overrides: its 64x64 average-pool branch (psmnet/modules.py:88 of the reference)

    def check_input_size(self, h, w):
        """Same constraint as the reference: two stride-2 stages at quarter resolution followed by
        output_padding=1 transposed convs need H and W to be multiples of 16 (stereodpnet/modules.py:208-227)."""
        if h % 16 or w % 16:
            raise ValueError(f"input size {h}x{w}: height and width must be multiples of 16")
        if min(h, w) // 4 < self.min_quarter_size:
            raise ValueError(f"input size {h}x{w}: this model needs at least {4 * self.min_quarter_size} pixels per side")

    def refresh(self):
        """Re-pack kernel-layout weights after parameters changed.  The eval-mode plans (folded BatchNorm running statistics +
        packed weights of PSMNetHGAggregation / ANM / CostVolumeSDP, the fused encoder copy) are built lazily on the first eval
        forward; load_state_dict(), every train()/eval() transition and every .to()/.cuda()/.half() (``_apply``) drop them, so
        eval -> train N steps -> eval never runs on pre-training weights."""
        self.__dict__.pop("_enc_fused", None)
        self.__dict__.pop("_in_buf", None)
        for m in self.modules():
            if m is not self and hasattr(m, "refresh"):
                m.refresh()

    def train(self, mode: bool = True):
        self.refresh()                       # an optimizer step may have happened since the plans were packed
        return super().train(mode)

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self.refresh()
        return out

    def reference_init(self):
        """Weight initialisation of the reference constructors (src/model/stereodpnet/mainmodel.py:49-64 and the loops inside
        the hourglass / aggregation / PSMNet encoder, stereodpnet/modules.py:229-239,298-308, psmnet/modules.py:112-127):
        N(0, sqrt(2 / (prod(kernel) * C_out))) for every Conv2d / Conv3d / ConvTranspose3d -- this includes, as in the
        reference, the zero-initialised `conv_offset` convolutions of the deformable layers, whose offsets are therefore
        non-zero at random init -- BatchNorm weight 1 / bias 0, Linear bias 0.  Convolution biases, PReLU slopes, the
        InstanceNorm affine and the DeformConvPack weight Parameter keep their module defaults, as they do in the reference."""
        import math
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.ConvTranspose3d)):
                n = m.out_channels
                for k in m.kernel_size:
                    n *= k
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.bias.data.zero_()
        self.refresh()

    @staticmethod
    def convert_checkpoint_keys(state_dict):
        """Key layout of the released checkpoints (README.md:95-101 of the reference) -> this torch / torchvision:
        * `normal_estimator.grid` (NNet: `normal_module.grid`) is registered lazily by the reference (normal_module.py:91-99,
          nnet/normal_module_.py:57-65) and rebuilt here;
        * torchvision <= 0.14 FeaturePyramidNetwork stored plain convs (`fpn.inner_blocks.N.weight`), current torchvision wraps
          them in Conv2dNormActivation (`fpn.inner_blocks.N.0.weight`)."""
        import re
        out = {}
        for k, v in state_dict.items():
            if k.endswith(("normal_estimator.grid", "normal_module.grid")):
                continue
            k = re.sub(r"(\.fpn\.(?:inner|layer)_blocks\.\d+)\.(weight|bias)$", r"\1.0.\2", k)
            out[k] = v
        return out

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = self.convert_checkpoint_keys(state_dict)
        out = super().load_state_dict(sd, strict=strict, **kw)
        self.refresh()
        return out

    # ---- Lightning-style hooks used by main.py ------------------------------------------------------------
    def _synthetic_loader(self, training, batch_size):
        h, w = getattr(self.option, "synthetic_size", (448, 448))
        n = int(getattr(self.option, "synthetic_batches", 2))
        off = int(self.data_seed_offset)
        return [synthetic_batch(batch_size, h, w, training=training, seed=off + i) for i in range(n)]

    def train_dataloader(self):
        return self._synthetic_loader(True, self.option.batch_size)

    def val_dataloader(self):
        return self._synthetic_loader(False, 1)

    def test_dataloader(self):
        return self._synthetic_loader(False, self.option.batch_size)

    def training_step(self, batch, batch_idx):
        results = self.forward(batch)
        losses = {k: v for k, v in results.items() if "loss" in k and k != "final_loss"}
        for k, v in losses.items():
            self.log(k, v, prog_bar=True)
        return {"loss": results["final_loss"], "log": losses}

    def validation_step(self, batch, batch_idx):
        return self.forward(batch)

    def test_step(self, batch, batch_idx):
        return self.forward(batch)

    def configure_optimizers(self):
        opt = optimizer_selector(self.parameters(), self.option)
        sch = scheduler_selector(opt, self.option)
        return [opt], ([sch] if sch is not None else [])


class STEREODPNET(_StereoBase):
    def __init__(self, option):
        super().__init__()
        self._common_init(option)
        self.predict_normal = bool(option.model.predict_normal)
        self.feature_extraction = M.SDPFeatureExtraction(option)
        self.cost_volume = M.CostVolumeSDP(option, self.mindisp, self.maxdisp)
        self.aggregation = M.PSMNetHGAggregation(option.model.inplanes)
        self.normal_estimator = M.ANM(option, self.mindisp, self.maxdisp) if self.predict_normal else None
        self.regression_layer = M.disp_regression(self.mindisp, self.maxdisp, self.level)
        self.reference_init()


class PSMNET(_StereoBase):
    train_supported = True
    min_quarter_size = 64

    def __init__(self, option):
        super().__init__()
        self._common_init(option)
        self.feature_extraction = M.PSMFeatureExtraction(option)
        self.cost_volume = M.CostVolumePSM(option, self.mindisp, self.maxdisp)
        self.aggregation = M.PSMNetHGAggregation(option)
        self.regression_layer = M.disp_regression(self.mindisp, self.maxdisp, self.level)
        self.reference_init()
