"""Training path of the ANM normal branch (src/model/stereodpnet/normal_module.py:140-194): autograd Functions over the
sm_100a kernels, forward AND backward.

    fv      = gather(out3 at the k sampled levels) ++ normalised coordinates     dpf_anm_gather / dpf_anm_gather_bwd
    off_i   = conv3x3x3(x; W_off) + b_off                                        tcgen05 conv engine (fwd, dgrad, wgrad)
    z_i     = D3D(x, off_i; W)                                                   dpf_dcn3d_fwd / dpf_dcn3d_bwd_data / _bwd_weight
    f_i     = ReLU(BN_train(z_i + bias))                                         dpf_channel_stats / dpf_affine_act / dpf_bn_bwd_*
    n_convs : six dilated 2-D convs + LeakyReLU(0.1) on the (b*k) slices         cuDNN through torch autograd (adjacent op)
    normal  = mean_k sigmoid(bilinear x4) * 2 - 1                                dpf_anm_tail / dpf_anm_tail_bwd
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import _lib, ops
from .layers import KIND_3x3x3, TCConv3d
from .ops_dcn_bwd import DCNFn
from .ops_tail import anm_tail
from .ops_wgrad import conv3d_wgrad
from .train_ops import _affine_act, _bn_bwd, _npix, _teacher, batch_stats


class GatherFn(Function):
    """out3 [B,D,H4,W4,C] bf16 -> fv [B,K,H4,W4,64] bf16 (C cost channels, 3 coordinates, zero pad)."""

    @staticmethod
    def forward(ctx, out3, idx, coord, minmax):
        ctx.save_for_backward(idx)
        ctx.shape = out3.shape
        return ops.anm_gather(out3, idx, coord, minmax, 64)

    @staticmethod
    def backward(ctx, dfv):
        (idx,) = ctx.saved_tensors
        b, d, h4, w4, c = ctx.shape
        dfv = dfv.contiguous()
        a = dfv if dfv.dtype == torch.float32 else None
        bb = dfv if dfv.dtype == torch.bfloat16 else None
        assert a is not None or bb is not None
        dout3 = torch.empty(ctx.shape, device=dfv.device, dtype=torch.bfloat16)
        _lib.check(ops.lib().dpf_anm_gather_bwd(ops._p(a), ops._p(bb), ops._p(idx), ops._p(dout3), b, d, idx.shape[1], h4, w4, c,
                                                dfv.shape[-1], ops._stream()), "dpf_anm_gather_bwd")
        return dout3, None, None, None


class OffsetConvFn(Function):
    """offset = conv3x3x3(x; W[81,Cin,3,3,3]) + bias -> fp32 [B,K,H4,W4,96] (81 real channels, zero padded so that the voxel
    pitch is 384 B and the fp32 epilogue uses 128-bit stores); x is the 64-channel (zero padded) volume."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        wpad = F.pad(weight.detach(), (0, 0, 0, 0, 0, 0, 0, 0, 0, 96 - weight.shape[0]))
        bpad = F.pad(bias.detach().float(), (0, 96 - weight.shape[0])).contiguous()
        return TCConv3d(wpad, KIND_3x3x3, cin_pad=64)(x, shift=bpad, out_f32=True)

    @staticmethod
    def backward(ctx, doff):
        x, weight = ctx.saved_tensors
        cout, cin = weight.shape[:2]
        # 96 gradient channels (81 real) = a 64-wide and a 32-wide input window: two launches per output chunk chained
        # through the bf16 residual input of the engine (no fp32 partial-sum pass)
        dz_a = doff[..., :64].to(torch.bfloat16).contiguous()
        dz_b = doff[..., 64:].to(torch.bfloat16).contiguous()
        wt = weight.detach().float().transpose(0, 1).flip(2, 3, 4)                               # [cin, 81, 3,3,3]
        wa = torch.zeros(64, 64, 3, 3, 3, device=weight.device, dtype=torch.float32)
        wb = torch.zeros(64, 32, 3, 3, 3, device=weight.device, dtype=torch.float32)
        wa[:cin] = wt[:, :64]
        wb[:cin, :cout - 64] = wt[:, 64:]
        dx = TCConv3d(wb, KIND_3x3x3)(dz_b, residual=TCConv3d(wa, KIND_3x3x3)(dz_a))            # [.,64] bf16, pad channels zero
        dw = torch.cat([conv3d_wgrad(x, dz_a, KIND_3x3x3), conv3d_wgrad(x, dz_b, KIND_3x3x3)[:cout - 64]], 0)[:, :cin].to(weight.dtype)
        db = doff[..., :cout].reshape(-1, cout).sum(0)
        return dx, dw, db


class BNActFn(Function):
    """y = ReLU(BatchNorm3d_train(z + conv_bias)) on channels-last bf16.  With batch statistics the conv bias cancels in y (its
    gradient is exactly zero); it only enters the running mean."""

    @staticmethod
    def forward(ctx, z, gamma, beta, conv_bias, bn):
        mean, var, n = batch_stats(z)
        inv_std = torch.rsqrt(var + bn.eps)
        a = (gamma.float() * inv_std).contiguous()
        b = (beta.float() - mean * a).contiguous()
        y = _teacher(bn, _affine_act(z, a, b, None, 0.0))
        if bn.track_running_stats:
            m = bn.momentum
            unb = var * (n / max(n - 1, 1))
            bn.running_mean.mul_(1 - m).add_(mean + conv_bias.detach().float(), alpha=m)
            bn.running_var.mul_(1 - m).add_(unb, alpha=m)
            bn.num_batches_tracked += 1
            bn.__dict__["_dpf_last_stats"] = (mean + conv_bias.detach().float(), unb)
        ctx.save_for_backward(z, y, a, mean, inv_std)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, y, a, mean, inv_std = ctx.saved_tensors
        dz, _, dgamma, dbeta = _bn_bwd(dy.to(torch.bfloat16).contiguous(), y, z, a, mean, inv_std, True, False)
        return dz, dgamma, dbeta, torch.zeros_like(dbeta), None


class TailFn(Function):
    """x [B*K,H4,W4,3] bf16 -> normal [B,3,H,W] fp32."""

    @staticmethod
    def forward(ctx, x, b, k):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.bk = (b, k)
        return anm_tail(x, b, k)

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        b, k = ctx.bk
        dx = torch.zeros(x.shape, device=x.device, dtype=torch.float32)
        _lib.check(ops.lib().dpf_anm_tail_bwd(ops._p(x), ops._p(dout.float().contiguous()), ops._p(dx), b, k, x.shape[1], x.shape[2],
                                              ops._stream()), "dpf_anm_tail_bwd")
        return dx.to(torch.bfloat16), None, None


def anm_train(anm, out3: torch.Tensor, disp: torch.Tensor, batch: dict):
    """One (cost, disparity) pair through the normal branch in training mode -> (normal [B,3,H,W], offset1, offset2)."""
    b = out3.shape[0]
    kq = batch["K"].float().clone()
    kq[:, :2, :] = kq[:, :2, :] / 4.0
    kinv = torch.inverse(kq).contiguous()
    idx, coord, minmax = ops.anm_select(disp.detach().contiguous(), kinv, batch["abvalue"].float().contiguous(), anm.levels, anm.k)
    fv = GatherFn.apply(out3, idx, coord, minmax)
    x, offs = fv, []
    if not anm.use_deform:                                     # original_conv: two convbn_3d + ReLU on the conv engine
        from .train_ops import ConvBNAct, LayerCfg
        for seq in (anm.original_conv[0], anm.original_conv[2]):
            conv, bn = seq[0], seq[1]
            w = conv.weight
            if x.shape[-1] > w.shape[1]:                           # the gathered volume is zero-padded 35 -> 64 channels
                w = F.pad(w, (0, 0, 0, 0, 0, 0, 0, x.shape[-1] - w.shape[1]))
            x = ConvBNAct.apply(x, w, bn.weight, bn.bias, None, LayerCfg(KIND_3x3x3, True, bn))
        offs = [None, None]
    for i, (dc, act) in enumerate(((anm.deform_conv1, anm.act1), (anm.deform_conv2, anm.act2)) if anm.use_deform else ()):
        off = OffsetConvFn.apply(x, dc.conv_offset.weight, dc.conv_offset.bias)
        # layer 1 reads the gathered volume: only its 32 cost channels need a gradient (the coordinates are constants)
        z = DCNFn.apply(x, off, dc.weight, 32 if (i == 0 and out3.shape[-1] == 32) else 64)
        x = BNActFn.apply(z, act[0].weight, act[0].bias, dc.bias, act[0])
        offs.append(off[..., :81])
    f = x.view(b * anm.k, x.shape[2], x.shape[3], x.shape[4]).permute(0, 3, 1, 2)                # NCHW view, channels-last memory
    for m in anm.n_convs:
        conv = m[0]
        f = F.leaky_relu(F.conv2d(f, conv.weight.to(torch.bfloat16), None, 1, conv.dilation, conv.dilation), 0.1)
    normal = TailFn.apply(f.permute(0, 2, 3, 1), b, anm.k)
    return normal, offs[0], offs[1]
