"""Training path of the hot path: autograd Functions over the sm_100a kernels (forward AND backward).

Forward of one reference layer  ``Conv3d/ConvTranspose3d -> BatchNorm3d (batch statistics) [-> + residual] [-> ReLU]``
(convbn_3d, src/module/asm/basics.py:32-36; used by src/model/stereodpnet/modules.py:204-337):
    z = conv(x)                      tcgen05 engine (dpf_conv3d_fwd), raw bf16 output
    mean, var = stats(z)             dpf_channel_stats
    y = act(z*a + b + res)           dpf_affine_act
Backward:
    S1, S2 = sum g, sum g*z          dpf_bn_bwd_reduce      (g = dy masked by the saved ReLU output)
    dz, dres                         dpf_bn_bwd_apply
    dx = conv^T(dz)                  the SAME tcgen05 engine with transformed weights: a stride-1 conv's data gradient is a
                                     stride-1 conv with flipped/transposed taps, a stride-2 conv's is the transposed-conv kind,
                                     a transposed conv's is the stride-2 kind
    dW                               dpf_conv3d_wgrad: tcgen05, voxel positions as the GEMM K dimension, MN-major operands
                                     straight from the forward staging layout (all five kinds)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch.autograd import Function

from . import _lib, ops
from .layers import KIND_1x1x1, KIND_1x3x3, KIND_3x3x3, KIND_S2, KIND_T2, TCConv3d


def _npix(t):
    return t.numel() // t.shape[-1]


# SyncBN (the reference's `accelerator: "ddp"` sets Trainer(sync_batchnorm=True), config_/config_manager.py:57, main.py:55):
# when enabled, the per-channel partial sums of every train-mode BatchNorm on the 3-D path (forward: sum z, sum z^2; backward:
# sum g, sum g*z) are all-reduced over the data-parallel group, so that the statistics are those of the GLOBAL batch.  The
# tensors are [2C] fp32 (256-512 B); they are issued on the compute stream's NCCL communicator, in program order on every rank.
SYNC_BN = {"enabled": False, "group": None}


def set_sync_bn(enabled: bool, group=None):
    SYNC_BN["enabled"], SYNC_BN["group"] = bool(enabled), group


def _sync_world() -> int:
    import torch.distributed as dist
    if SYNC_BN["enabled"] and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(SYNC_BN["group"])
    return 1


def _sync_sums(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce of BatchNorm partial sums when SyncBN is on; identity otherwise."""
    if _sync_world() > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=SYNC_BN["group"])
    return t


def batch_stats(z: torch.Tensor):
    """(mean, biased var, global element count per channel) of a channels-last bf16 activation over all positions of the
    (global, when SyncBN is on) batch."""
    c, n = z.shape[-1], _npix(z)
    st = _sync_sums(ops.channel_stats(z.view(1, n, c))[0])
    n = n * _sync_world()
    mean = st[:, 0] / n
    var = (st[:, 1] / n - mean * mean).clamp_min(0.0)
    return mean, var, n


def bn_train_coefs(z: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bn) -> tuple:
    """Train-mode BatchNorm coefficients of a channels-last bf16 activation in three launches (two-pass statistics +
    dpf_bn_fwd_coefs): returns (fwd [4,C] fp32 = a | b | mean | inv_std, n).  The running statistics of `bn` are updated in place
    by the kernel (momentum, unbiased variance), `num_batches_tracked` by one tiny add."""
    c, n = z.shape[-1], _npix(z)
    st = _sync_sums(ops.channel_stats(z.view(1, n, c))[0])                      # [C,2] = sum z, sum z^2
    n = n * _sync_world()
    fwd = torch.empty(4, c, device=z.device, dtype=torch.float32)
    track = bn is not None and bn.track_running_stats and bn.running_mean is not None
    mom = 0.0
    if track:
        mom = bn.momentum if bn.momentum is not None else 1.0 / float(int(bn.num_batches_tracked) + 1)
    _lib.check(ops.lib().dpf_bn_fwd_coefs(ops._p(st), ops._p(gamma), ops._p(beta), float(n), float(bn.eps if bn is not None else 1e-5),
                                          float(mom), ops._p(bn.running_mean) if track else None, ops._p(bn.running_var) if track else None,
                                          ops._p(fwd), c, ops._stream()), "dpf_bn_fwd_coefs")
    if track:
        bn.num_batches_tracked += 1
    return fwd, n


def bn_train_bwd(dy, y, z, fwd, relu, want_dres, slope=0.0):
    """Backward of act(BN_train(z) [+ res]): (dz, dres, dgamma, dbeta) with dpf_bn_bwd_reduce -> dpf_bn_bwd_coefs -> dpf_bn_bwd_apply
    (SyncBN keeps the element-wise path: it needs the local AND the all-reduced sums)."""
    if _sync_world() > 1:
        return _bn_bwd(dy, y, z, fwd[0], fwd[2], fwd[3], relu, want_dres, slope)
    c, n = z.shape[-1], _npix(z)
    sums = torch.empty(2 * c, device=z.device, dtype=torch.float32)
    _lib.check(ops.lib().dpf_bn_bwd_reduce(ops._p(dy), ops._p(y), ops._p(z), ops._p(sums), n, c, int(relu), float(slope), ops._stream()),
               "dpf_bn_bwd_reduce")
    out = torch.empty(6, c, device=z.device, dtype=torch.float32)                 # dgamma | dbeta | coef [4,C]
    _lib.check(ops.lib().dpf_bn_bwd_coefs(ops._p(sums), ops._p(fwd), float(n), ops._p(out[0]), ops._p(out[1]), ops._p(out[2:]), c,
                                          ops._stream()), "dpf_bn_bwd_coefs")
    dz = torch.empty_like(z)
    dres = torch.empty_like(z) if want_dres else None
    _lib.check(ops.lib().dpf_bn_bwd_apply(ops._p(dy), ops._p(y), ops._p(z), ops._p(out[2:]), ops._p(dz), ops._p(dres), n, c, int(relu),
                                          float(slope), ops._stream()), "dpf_bn_bwd_apply")
    return dz, dres, out[0], out[1]


# Test hook (tests/test_gpu_teacher_forced.py): {id(BatchNorm module): the ORACLE's output of that layer, channels-last bf16}.
# A layer found here still runs its full forward on the kernels, but hands the oracle's activation downstream and to its own
# backward (ReLU mask), so that a gradient comparison isolates the backward kernels from bf16-induced ReLU-mask flips and from
# accumulated forward error.  TEACHER_ERR records the layer's own forward error against the teacher.  None in production.
TEACHER = None
TEACHER_ERR = {}


def _teacher(bn, y):
    if TEACHER is None or bn is None or id(bn) not in TEACHER:
        return y
    t = TEACHER[id(bn)]
    assert t.shape == y.shape and t.dtype == y.dtype, (t.shape, y.shape)
    TEACHER_ERR[id(bn)] = float((y.float() - t.float()).abs().max() / t.float().abs().max().clamp_min(1e-6))
    return t.clone()


def _affine_act(z, scale, bias, res, slope, out=None):
    y = torch.empty_like(z) if out is None else out
    _lib.check(ops.lib().dpf_affine_act(ops._p(z), ops._p(scale), ops._p(bias), ops._p(res), ops._p(y), _npix(z), z.shape[-1],
                                        float(slope), ops._stream()), "dpf_affine_act")
    return y


def _bn_bwd(dy, y, z, a, mean, inv_std, relu, want_dres, slope=0.0):
    c, n = z.shape[-1], _npix(z)
    sums = torch.empty(2 * c, device=z.device, dtype=torch.float32)
    _lib.check(ops.lib().dpf_bn_bwd_reduce(ops._p(dy), ops._p(y), ops._p(z), ops._p(sums), n, c, int(relu), float(slope), ops._stream()),
               "dpf_bn_bwd_reduce")
    # SyncBN: the mean-gradient terms of dz are sums over the GLOBAL batch; dgamma / dbeta stay LOCAL sums (the gradient
    # all-reduce averages them like every other parameter gradient -- the convention of torch.nn.SyncBatchNorm)
    dgamma, dbeta = inv_std * (sums[c:] - mean * sums[:c]), sums[:c].clone()
    _sync_sums(sums)
    ng = n * _sync_world()
    s1, s2 = sums[:c], sums[c:]
    centred = s2 - mean * s1
    coef = torch.cat([a, s1 / ng, inv_std * inv_std * centred / ng, mean]).contiguous()
    dz = torch.empty_like(z)
    dres = torch.empty_like(z) if want_dres else None
    _lib.check(ops.lib().dpf_bn_bwd_apply(ops._p(dy), ops._p(y), ops._p(z), ops._p(coef), ops._p(dz), ops._p(dres), n, c, int(relu),
                                          float(slope), ops._stream()), "dpf_bn_bwd_apply")
    return dz, dres, dgamma, dbeta


def _wgrad(x, dz, weight, kind):
    """dW on the tcgen05 wgrad kernel (dpf_conv3d_wgrad) for every kind; the transposed kind swaps the roles of x and dz."""
    from .ops_wgrad import conv3d_wgrad
    return conv3d_wgrad(x, dz, kind).to(weight.dtype)


def _dgrad(dz, weight, kind):
    if kind in (KIND_3x3x3, KIND_1x3x3, KIND_1x1x1):      # stride 1: same kind, taps flipped, channels transposed
        return TCConv3d(weight.transpose(0, 1).flip(2, 3, 4), kind)(dz)
    if kind == KIND_3x3x3:
        return TCConv3d(weight.transpose(0, 1).flip(2, 3, 4), KIND_3x3x3)(dz)
    if kind == KIND_S2:                     # adjoint of a stride-2 conv = transposed conv with the same weight tensor
        return TCConv3d(weight, KIND_T2, transposed=True)(dz)
    return TCConv3d(weight, KIND_S2)(dz)    # adjoint of a transposed conv = stride-2 conv with the same weight tensor


@dataclass
class LayerCfg:
    kind: int
    relu: bool
    bn: Optional[torch.nn.BatchNorm3d]      # running statistics are updated in place (momentum, unbiased variance)
    slope: float = 0.0                      # negative-side slope of the activation when relu (0 = ReLU; StereoNet's filter: 0.2)


class ConvBNAct(Function):
    """y = act(BN_train(conv(x)) + residual); all activations [B,D,H,W,C] bf16."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, residual, cfg: LayerCfg):
        z = TCConv3d(weight, cfg.kind, transposed=cfg.kind == KIND_T2)(x)
        fwd, n = bn_train_coefs(z, gamma, beta, cfg.bn)                     # a | b | mean | inv_std; running statistics updated
        y = _teacher(cfg.bn, _affine_act(z, fwd[0], fwd[1], residual, cfg.slope if cfg.relu else 1.0))
        if cfg.bn is not None and cfg.bn.track_running_stats:
            cfg.bn.__dict__["_dpf_last_fwd"] = (fwd, n)                       # train_asm replays the momentum update from these
        ctx.save_for_backward(x, weight, z, y, fwd)
        ctx.cfg, ctx.has_res = cfg, residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, z, y, fwd = ctx.saved_tensors
        cfg = ctx.cfg
        dy = dy.to(torch.bfloat16).contiguous()
        dz, dres, dgamma, dbeta = bn_train_bwd(dy, y, z, fwd, cfg.relu, ctx.has_res, cfg.slope)
        dx = _dgrad(dz, weight, cfg.kind) if ctx.needs_input_grad[0] else None
        dw = _wgrad(x, dz, weight, cfg.kind)
        return dx, dw, dgamma, dbeta, dres, None


class BN2dTrainFn(Function):
    """Train-mode BatchNorm2d on an NCHW-shaped, channels-last bf16 activation (modules.EncoderBatchNorm2d): batch statistics,
    normalisation and backward on the kernels of the 3-D path; running statistics updated as nn.BatchNorm2d does (momentum,
    unbiased variance).  With SyncBN switched on (set_sync_bn) the partial sums are all-reduced like those of the 3-D layers."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn):
        xh = x.permute(0, 2, 3, 1)                                   # [N,H,W,C] view of the channels-last memory
        assert xh.is_contiguous()
        fwd, _n = bn_train_coefs(xh, gamma, beta, bn)
        y = torch.empty_like(x)                                      # NCHW-shaped, channels-last: a fresh tensor, not a view (the
        _affine_act(xh, fwd[0], fwd[1], None, 1.0, out=y.permute(0, 2, 3, 1))   # in-place ReLU / PReLU that follows needs a non-view)
        ctx.save_for_backward(x, fwd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, fwd = ctx.saved_tensors
        xh = x.permute(0, 2, 3, 1)
        dyh = dy.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        dz, _, dgamma, dbeta = bn_train_bwd(dyh, None, xh, fwd, False, False)
        return dz.permute(0, 3, 1, 2), dgamma, dbeta, None


class HeadConv(Function):
    """cost = conv3x3x3(x, w[1,C,3,3,3]) + prev  (classif{k}.2 + the cumulative adds of modules.py:323-325); fp32 [B,D,H,W,1]."""

    @staticmethod
    def forward(ctx, x, weight, prev):
        y = TCConv3d(weight, KIND_3x3x3)(x, residual=prev, out_f32=True)
        ctx.save_for_backward(x, weight)
        ctx.has_prev = prev is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.float().contiguous()
        c = x.shape[-1]
        dzp = torch.zeros(*dy.shape[:-1], c, device=dy.device, dtype=torch.bfloat16)
        dzp[..., 0] = dy[..., 0]
        wt = torch.zeros(c, c, 3, 3, 3, device=weight.device, dtype=torch.float32)
        wt[:, 0] = weight[0].flip(1, 2, 3)                     # W'[ci, 0, k] = W[0, ci, flip(k)]
        dx = TCConv3d(wt, KIND_3x3x3)(dzp)
        dw = _wgrad(x, dy.to(torch.bfloat16), weight, KIND_3x3x3)
        return dx, dw, (dy if ctx.has_prev else None)


class CostVolumeFn(Function):
    """Integer-shift volume (dpf_costvol_fwd / dpf_costvol_bwd)."""

    @staticmethod
    def forward(ctx, ref, tgt, shifts, mode, groups):
        ctx.save_for_backward(ref, tgt)
        ctx.args = (list(shifts), mode, groups)
        return ops.costvol_fwd(ref, tgt, shifts, mode, groups)

    @staticmethod
    def backward(ctx, dvol):
        ref, tgt = ctx.saved_tensors
        shifts, mode, groups = ctx.args
        dref, dtgt = ops.costvol_bwd(ref, tgt, dvol.to(torch.bfloat16).contiguous(), shifts, mode, groups)
        return dref, dtgt, None, None, None


class RegressFn(Function):
    """Fused upsample + soft-argmin (dpf_regress_fwd / dpf_regress_bwd); cost [B,D,H4,W4] fp32 -> disparity [B,H,W]."""

    @staticmethod
    def forward(ctx, cost, mindisp, step):
        cost = cost.contiguous()
        ctx.save_for_backward(cost)
        ctx.args = (mindisp, step)
        return ops.regress_fwd(cost, mindisp, step, False)[0]

    @staticmethod
    def backward(ctx, ddisp):
        (cost,) = ctx.saved_tensors
        return ops.regress_bwd(cost, ddisp.float().contiguous(), *ctx.args), None, None
