"""Training path of the StereoDPNet ASM cost volume: autograd Functions over the sm_100a kernels.

Mirrors CostVolume.build_concat_volume + MaskingAttention.forward + subpixel_shift.forward of the reference
(src/model/stereodpnet/modules.py:181-197, src/module/asm/asm.py:87-173) in train mode: the attention's BatchNorm3d runs on
batch statistics *per call* (once for the forward-shifted reference features, once for the backward-shifted target
features), InstanceNorm3d on per-(b,c) statistics, and every piece has a hand-written backward.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib, ops
from .layers import KIND_1x1x1, KIND_1x3x3, TCConv3d
from .train_ops import ConvBNAct, LayerCfg, _dgrad, _npix, _wgrad


class AsmSampleFn(Function):
    """samples [B,S,H,W,C] = table-driven resampling of feat [B,H,W,C] (dpf_asm_sample_fwd / dpf_asm_sample_bwd)."""

    @staticmethod
    def forward(ctx, feat, tables):
        ctx.tables = tables
        return ops.asm_sample(feat.contiguous(), tables)

    @staticmethod
    def backward(ctx, ds):
        t = ctx.tables
        ds = ds.to(torch.bfloat16).contiguous()
        b, s, h, w, c = ds.shape
        dfeat = torch.empty(b, h, w, c, device=ds.device, dtype=torch.float32)
        _lib.check(ops.lib().dpf_asm_sample_bwd(ops._p(ds), ops._p(dfeat), b, h, w, c, s, ops._p(t["ri"]), ops._p(t["rw"]),
                                                ops._p(t["ci"]), ops._p(t["cw"]), ops._stream()), "dpf_asm_sample_bwd")
        return dfeat.to(torch.bfloat16), None


class ConvOnly(Function):
    """z = conv(x, w) on the tcgen05 engine, no normalisation; backward = dgrad on the same engine + cuDNN wgrad."""

    @staticmethod
    def forward(ctx, x, weight, kind):
        ctx.save_for_backward(x, weight)
        ctx.kind = kind
        return TCConv3d(weight, kind)(x)

    @staticmethod
    def backward(ctx, dz):
        x, weight = ctx.saved_tensors
        dz = dz.to(torch.bfloat16).contiguous()
        return _dgrad(dz, weight, ctx.kind), _wgrad(x, dz, weight, ctx.kind), None


class AsmBlendFn(Function):
    """y [B,H,W,C] = mean_s( x_s * softmax_s( sigmoid( InstanceNorm(l)_s ) ) ); InstanceNorm3d(affine) over (S,H,W) per (b,c)."""

    @staticmethod
    def forward(ctx, samples, logits, gamma, beta, eps):
        b, s, h, w, c = samples.shape
        n = float(s * h * w)
        st = ops.channel_stats(logits)
        mean = st[..., 0] / n
        var = (st[..., 1] / n - mean * mean).clamp_min(0.0)
        inv_std = torch.rsqrt(var + eps)
        a = (gamma.float().unsqueeze(0) * inv_std).contiguous()
        d = (beta.float().unsqueeze(0) - mean * a).contiguous()
        y = torch.empty(b, 1, h, w, c, device=samples.device, dtype=torch.bfloat16)
        ops.asm_blend(samples, logits, a, d, y, 0, 1, 0)
        ctx.save_for_backward(samples, logits, a, d, mean, inv_std)
        return y[:, 0]

    @staticmethod
    def backward(ctx, dy):
        samples, logits, a, d, mean, inv_std = ctx.saved_tensors
        b, s, h, w, c = samples.shape
        dy = dy.to(torch.bfloat16).contiguous()
        dsamples, dlhat = torch.empty_like(samples), torch.empty_like(samples)
        _lib.check(ops.lib().dpf_asm_blend_bwd(ops._p(samples), ops._p(logits), ops._p(a), ops._p(d), ops._p(dy), ops._p(dsamples),
                                               ops._p(dlhat), b, h, w, c, s, 1, 0, 1, 0, c, ops._stream()), "dpf_asm_blend_bwd")
        # InstanceNorm backward, one (reduce, apply) pair per sample: statistics are per (b, c)
        n = s * h * w
        dlogits = torch.empty_like(logits)
        dgamma = torch.zeros(c, device=samples.device, dtype=torch.float32)
        dbeta = torch.zeros(c, device=samples.device, dtype=torch.float32)
        sums = torch.empty(2 * c, device=samples.device, dtype=torch.float32)
        for i in range(b):
            _lib.check(ops.lib().dpf_bn_bwd_reduce(ops._p(dlhat[i]), None, ops._p(logits[i]), ops._p(sums), n, c, 0, 0.0, ops._stream()),
                       "dpf_bn_bwd_reduce")
            s1, s2 = sums[:c], sums[c:]
            centred = s2 - mean[i] * s1
            dgamma += inv_std[i] * centred
            dbeta += s1
            coef = torch.cat([a[i], s1 / n, inv_std[i] * inv_std[i] * centred / n, mean[i]]).contiguous()
            _lib.check(ops.lib().dpf_bn_bwd_apply(ops._p(dlhat[i]), None, ops._p(logits[i]), ops._p(coef), ops._p(dlogits[i]), None, n, c,
                                                  0, 0.0, ops._stream()), "dpf_bn_bwd_apply")
        return dsamples, dlogits, dgamma, dbeta, None


def asm_volume_train(cv, ref_feat, tar_feat):
    """Train-mode CostVolumeSDP.forward: [B,D,H4,W4,2C] bf16 with autograd through every stage."""
    b, h, w, c = ref_feat.shape
    att = cv.attention_layer
    conv1, bn1, conv2 = att.mask_convs[0], att.mask_convs[1], att.mask_convs[3][0]
    inorm = att.normalize
    levels = [(cv.level, cv.costrange[0])] if cv.cached_first_level else [(1, d) for d in cv.costrange]
    slices = []
    for rep, disp in levels:
        halves, stats = [], []
        for feat, direction in ((ref_feat, "forward"), (tar_feat, "backward")):
            smp = cv.sample(feat, cv._tab(h, w, disp, direction, feat.device), train=True)
            m = ConvBNAct.apply(smp, conv1.weight, bn1.weight, bn1.bias, None, LayerCfg(KIND_1x3x3, True, bn1))
            last = bn1.__dict__.get("_dpf_last_fwd")                          # (a | b | mean | inv_std, n) of the call above
            if last is not None:
                fwd_c, n_el = last
                var_b = (1.0 / (fwd_c[3] * fwd_c[3]) - bn1.eps).clamp_min(0.0)
                stats.append((fwd_c[2].clone(), var_b * (n_el / max(n_el - 1, 1))))
            else:
                stats.append(None)
            logits = ConvOnly.apply(m, conv2.weight, KIND_1x1x1)
            halves.append(AsmBlendFn.apply(smp, logits, inorm.weight, inorm.bias, inorm.eps))
        y = torch.cat(halves, -1).unsqueeze(1)                           # [B,1,H,W,2C]
        if rep > 1:
            # the reference evaluates the shared attention module once per level and view (16 calls) on inputs that are
            # identical across levels (cached grids): same batch statistics every time, so the outputs are shared and only
            # the running statistics see the repeated momentum updates, in the reference's call order (fwd, bwd, fwd, ...)
            with torch.no_grad():
                for _ in range(rep - 1):
                    for st in stats:
                        if st is not None and bn1.track_running_stats:
                            bn1.running_mean.mul_(1 - bn1.momentum).add_(st[0], alpha=bn1.momentum)
                            bn1.running_var.mul_(1 - bn1.momentum).add_(st[1], alpha=bn1.momentum)
                            bn1.num_batches_tracked += 1
            y = y.expand(b, rep, h, w, 2 * c)
        slices.append(y)
    return torch.cat(slices, 1).contiguous()
