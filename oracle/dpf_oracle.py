"""CPU restatement (fp32, plain PyTorch ops) of the DualPixelFace stereo hot path.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function cites the
reference file:line it follows (paths relative to the reference checkout).
The style is deliberately functional: each stage takes a flat ``state`` dict
(name -> tensor, the reference's ``state_dict`` key layout) and a key prefix,
so the same seeded synthetic weights can be fed to the reference (golden
generation), to this oracle and to the CUDA product.

Nothing here touches CUDA; everything runs on whatever device its inputs are on
(CPU in the tests and in the bench's cpu_baseline leg; tests/test_gpu_fullsize.py also
executes it on GPU tensors as the full-size checker).

Parity pins: bit-exact against the unmodified reference code run through
tests/golden/ref_shim.py (tests/test_oracle_golden.py, committed fixtures under
tests/golden/); the 3-D deformable convolution, which the reference only implements
in CUDA, against the reference's own kernels built into oracle/_ref/DCN.so by
oracle/build_ref_dcn.py (tests/test_gpu_dcn_reference.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------
# constants of the path
# --------------------------------------------------------------------------------------


def cost_range(mindisp: float, maxdisp: float, level: int) -> np.ndarray:
    """Quarter-resolution disparity hypotheses.

    src/model/stereodpnet/modules.py:144-145 and src/model/psmnet/modules.py:184-185:
    ``arange(level) * ((maxdisp/4 - mindisp/4)/level) + mindisp/4`` in float64.
    """
    return np.arange(int(level), dtype=np.float64) * ((maxdisp / 4.0 - mindisp / 4.0) / float(level)) + mindisp / 4.0


def disparity_bins(mindisp: float, maxdisp: float, level: int) -> np.ndarray:
    """Full-resolution regression bins, src/model/stereodpnet/modules.py:345."""
    n = int(4 * level)
    return np.arange(n, dtype=np.float64) * ((maxdisp - mindisp) / float(n)) + mindisp


# --------------------------------------------------------------------------------------
# integer-shift volumes (PSMNet concat / GwC, StereoNet difference)
# --------------------------------------------------------------------------------------


def _row_windows(h: int, d: int) -> Tuple[slice, slice, slice]:
    """(dst rows, ref rows, tgt rows) for an integer row shift d.

    src/model/psmnet/modules.py:229-238: d>0 -> rows [:-d] <- ref[:-d], tgt[d:];
    d<0 -> rows [-d:] <- ref[-d:], tgt[:d]; d==0 -> everything.
    """
    if d == 0:
        return slice(0, h), slice(0, h), slice(0, h)
    if d > 0:
        return slice(0, h - d), slice(0, h - d), slice(d, h)
    return slice(-d, h), slice(-d, h), slice(0, h + d)


def psm_concat_volume(ref: torch.Tensor, tgt: torch.Tensor, crange: Sequence[float]) -> torch.Tensor:
    """src/model/psmnet/modules.py:223-241 -> [B, 2C, D, H, W]; shifts are int(disp) (truncation)."""
    b, c, h, w = ref.shape
    vol = ref.new_zeros(b, 2 * c, len(crange), h, w)
    for i, disp in enumerate(crange):
        dst, rr, tr = _row_windows(h, int(disp))
        vol[:, :c, i, dst] = ref[:, :, rr]
        vol[:, c:, i, dst] = tgt[:, :, tr]
    return vol


def psm_gwc_volume(ref: torch.Tensor, tgt: torch.Tensor, crange: Sequence[float], groups: int) -> torch.Tensor:
    """src/model/psmnet/modules.py:215-221, 243-262 -> [B, G, D, H, W]; note the minus sign (:221)."""
    b, c, h, w = ref.shape
    assert c % groups == 0
    vol = ref.new_zeros(b, groups, len(crange), h, w)
    for i, disp in enumerate(crange):
        dst, rr, tr = _row_windows(h, int(disp))
        prod = ref[:, :, rr] * tgt[:, :, tr]
        vol[:, :, i, dst] = -prod.reshape(b, groups, c // groups, prod.shape[2], w).mean(dim=2)
    return vol


def diff_volume(ref: torch.Tensor, tgt: torch.Tensor, crange: Sequence[float]) -> torch.Tensor:
    """src/model/stereonet/mainmodel.py:100-114 -> [B, C, D, H, W] (ref - shifted tgt)."""
    b, c, h, w = ref.shape
    vol = ref.new_zeros(b, c, len(crange), h, w)
    for i, disp in enumerate(crange):
        dst, rr, tr = _row_windows(h, int(disp))
        vol[:, :, i, dst] = ref[:, :, rr] - tgt[:, :, tr]
    return vol


# --------------------------------------------------------------------------------------
# sub-pixel shift (ASM sampling), src/module/asm/asm.py:9-127
# --------------------------------------------------------------------------------------


def shift_grid(h: int, w: int, disp: float, direction: str, device=None) -> torch.Tensor:
    """Normalised sampling grid [H, W, 2] (x, y), src/module/asm/asm.py:32-49.

    Rows are displaced by ``+disp`` ('forward') or ``-disp`` ('backward'); columns are not.
    The op order (add, divide by n-1, times 2.0, minus 1.0, all fp32) is the reference's.
    """
    sign = 1.0 if direction == "forward" else -1.0
    deltar = torch.tensor(float(sign * disp), device=device)
    ys = torch.arange(0.0, h, device=device) + deltar
    xs = torch.arange(0.0, w, device=device) + torch.tensor(0.0, device=device)
    yv, xv = torch.meshgrid([ys, xs], indexing="ij")
    xv = xv / (w - 1) * 2.0 - 1.0
    yv = yv / (h - 1) * 2.0 - 1.0
    return torch.stack([xv, yv], dim=-1)


def phase_terms(h: int, w: int, disp: float, direction: str, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin multipliers of the Fourier shift, src/module/asm/asm.py:59-75 (even H, W)."""
    sign = 1.0 if direction == "forward" else -1.0
    deltar = torch.tensor(float(sign * disp), device=device) / h
    nr = torch.cat([torch.arange(0.0, math.ceil(h // 2)), torch.arange(-float(np.fix(h // 2)), 0.0)]).to(device)
    nc = torch.cat([torch.arange(0.0, math.ceil(w // 2)), torch.arange(-float(np.fix(w // 2)), 0.0)]).to(device)
    nr, nc = torch.meshgrid([nr, nc], indexing="ij")
    arg = torch.tensor(2.0 * np.pi, device=device) * (deltar * nr + 0.0 * nc)
    return torch.cos(arg), torch.sin(arg)


def subpixel_samples(x: torch.Tensor, disp: float, direction: str,
                     nearest: bool = True, bilinear: bool = True, phase: bool = True) -> List[torch.Tensor]:
    """The (up to) three resampled copies of ``x`` [B,C,H,W], src/module/asm/asm.py:87-127.

    nearest : grid_sample(mode='nearest') with the DEFAULT align_corners=False (:96) on a grid that
              was normalised for align_corners=True -- reproduced as is.
    bilinear: grid_sample(bilinear, align_corners=True) (:101-102).
    phase   : legacy ``torch.rfft(x, 2, onesided=False)`` -> rotate -> ``torch.irfft(.., 2, onesided=False)``
              (:112-125).  The legacy irfft is a C2R transform that consumes only the first W/2+1
              columns of the full spectrum; restated with torch.fft accordingly.
    """
    b, c, h, w = x.shape
    out: List[torch.Tensor] = []
    if nearest or bilinear:
        grid = shift_grid(h, w, disp, direction, x.device).expand(b, -1, -1, -1).to(x.dtype)
    if nearest:
        out.append(F.grid_sample(x, grid, mode="nearest", align_corners=False))
    if bilinear:
        out.append(F.grid_sample(x, grid, mode="bilinear", align_corners=True))
    if phase:
        cos_t, sin_t = phase_terms(h, w, disp, direction, x.device)
        spec = torch.fft.fft2(x.float())
        fr, fi = spec.real, spec.imag
        fr2 = fr * cos_t - fi * sin_t
        fi2 = fi * cos_t + fr * sin_t
        rot = torch.complex(fr2, fi2)[..., : w // 2 + 1]
        out.append(torch.fft.irfft2(rot, s=(h, w)).to(x.dtype))
    return out


# --------------------------------------------------------------------------------------
# small functional helpers over a flat state dict
# --------------------------------------------------------------------------------------


def _bn(x: torch.Tensor, st: State, prefix: str, training: bool, momentum: float = 0.1, eps: float = 1e-5,
        stats_out: Optional[dict] = None) -> torch.Tensor:
    """nn.BatchNorm{2,3}d semantics (src/module/asm/basics.py:17-36 build them with defaults)."""
    rm = st.get(prefix + ".running_mean")
    rv = st.get(prefix + ".running_var")
    if training and stats_out is not None:
        dims = [0] + list(range(2, x.dim()))
        stats_out[prefix] = (x.mean(dim=dims).detach(), x.var(dim=dims, unbiased=False).detach())
    # never mutate the caller's running statistics: the oracle is a pure function
    rm_c = rm.clone() if (rm is not None and training) else rm
    rv_c = rv.clone() if (rv is not None and training) else rv
    return F.batch_norm(x, rm_c, rv_c, st[prefix + ".weight"], st[prefix + ".bias"], training, momentum, eps)


def _conv3(x, st, key, stride=1, pad=1, bias=False):
    return F.conv3d(x, st[key + ".weight"], st.get(key + ".bias") if bias else None, stride=stride, padding=pad)


def _convT3(x, st, key):
    """nn.ConvTranspose3d(k3, s2, p1, output_padding=1), src/model/stereodpnet/modules.py:219-227."""
    return F.conv_transpose3d(x, st[key + ".weight"], None, stride=2, padding=1, output_padding=1)


def _conv2(x, st, key, stride=1, pad=1, dil=1, bias=False, groups=1):
    return F.conv2d(x, st[key + ".weight"], st.get(key + ".bias") if bias else None, stride=stride,
                    padding=pad, dilation=dil, groups=groups)


def _convbn2(x, st, key, stride, pad, dil, training, stats=None):
    """convbn(), src/module/asm/basics.py:17-22: padding = dilation if dilation > 1 else pad."""
    p = dil if dil > 1 else pad
    y = _conv2(x, st, key + ".0", stride=stride, pad=p, dil=dil)
    return _bn(y, st, key + ".1", training, stats_out=stats)


def _convbn3(x, st, key, stride, training, stats=None):
    """convbn_3d(), src/module/asm/basics.py:32-36."""
    y = _conv3(x, st, key + ".0", stride=stride, pad=1)
    return _bn(y, st, key + ".1", training, stats_out=stats)


# --------------------------------------------------------------------------------------
# ASM masking attention + StereoDPNet volume
# --------------------------------------------------------------------------------------


def masking_attention(x: torch.Tensor, st: State, prefix: str, training: bool, activation: str = "sigmoid",
                      feature_fetch: bool = False, stats=None) -> torch.Tensor:
    """src/module/asm/asm.py:131-173.  x: [B, C, S, H, W] (S = number of sampling modes)."""
    m = F.conv3d(x, st[prefix + ".mask_convs.0.weight"], None, padding=(0, 1, 1))
    m = _bn(m, st, prefix + ".mask_convs.1", training, stats_out=stats)
    m = F.relu(m)
    m = F.conv3d(m, st[prefix + ".mask_convs.3.0.weight"], None)
    # nn.InstanceNorm3d(affine=True): per-(b,c) biased statistics, eps 1e-5, no running stats (:138)
    m = F.instance_norm(m, None, None, st[prefix + ".mask_convs.3.1.weight"], st[prefix + ".mask_convs.3.1.bias"],
                        True, 0.1, 1e-5)
    if activation == "sigmoid":
        a = torch.sigmoid(m)
    elif activation == "relu":
        a = F.prelu(m, st[prefix + ".activation.weight"])
    else:
        raise NotImplementedError(activation)
    y = x * F.softmax(a, dim=2)
    if feature_fetch:
        return torch.mean(y ** 2, 2) - torch.mean(y, 2) ** 2
    return torch.mean(y, 2)


def sdp_cost_volume(ref: torch.Tensor, tgt: torch.Tensor, st: State, prefix: str, crange: Sequence[float],
                    training: bool, modes=(True, True, True), activation: str = "sigmoid",
                    feature_fetch: bool = False, cached_first_level: bool = True, stats=None) -> torch.Tensor:
    """src/model/stereodpnet/modules.py:181-197 -> [B, 2C, D, H, W].

    ``cached_first_level=True`` reproduces the reference as shipped: ``subpixel_shift.make_grid`` keeps
    the grids / phase terms of the FIRST call and never rebuilds them (src/module/asm/asm.py:29-30,56-57),
    so every level samples with ``crange[0]``.  ``False`` gives the evidently intended per-level shifts.
    In train mode the attention's BatchNorm is evaluated once per call (16 calls), each on its own batch
    statistics -- identical inputs give identical outputs, so level slices stay equal in cached mode.
    """
    b, c, h, w = ref.shape
    vol = ref.new_zeros(b, 2 * c, len(crange), h, w)
    for i, disp in enumerate(crange):
        d = crange[0] if cached_first_level else disp
        fwd = torch.stack(subpixel_samples(ref, d, "forward", *modes), dim=2)
        bwd = torch.stack(subpixel_samples(tgt, d, "backward", *modes), dim=2)
        vol[:, :c, i] = masking_attention(fwd, st, prefix + ".attention_layer", training, activation,
                                          feature_fetch, stats)
        vol[:, c:, i] = masking_attention(bwd, st, prefix + ".attention_layer", training, activation,
                                          feature_fetch, stats)
    return vol


# --------------------------------------------------------------------------------------
# 3-D hourglass aggregation, src/model/stereodpnet/modules.py:204-337 (PSMNet twin: psmnet/modules.py:279-416)
# --------------------------------------------------------------------------------------


def hourglass(x, presqu, postsqu, st: State, p: str, training: bool, stats=None):
    """PSMNetHourglass.forward, src/model/stereodpnet/modules.py:241-260."""
    out = F.relu(_convbn3(x, st, p + ".conv1.0", 2, training, stats))
    pre = _convbn3(out, st, p + ".conv2", 1, training, stats)
    pre = F.relu(pre + postsqu) if postsqu is not None else F.relu(pre)
    out = F.relu(_convbn3(pre, st, p + ".conv3.0", 2, training, stats))
    out = F.relu(_convbn3(out, st, p + ".conv4.0", 1, training, stats))
    up5 = _bn(_convT3(out, st, p + ".conv5.0"), st, p + ".conv5.1", training, stats_out=stats)
    post = F.relu(up5 + (presqu if presqu is not None else pre))
    out = _bn(_convT3(post, st, p + ".conv6.0"), st, p + ".conv6.1", training, stats_out=stats)
    return out, pre, post


def aggregation_lowres(cost: torch.Tensor, st: State, p: str, training: bool, stats=None):
    """PSMNetHGAggregation.forward up to (not including) the trilinear upsample, modules.py:310-325.

    Returns ([cost3, cost2, cost1] each [B,1,D,H4,W4], [out3, out2, out1] each [B,C,D,H4,W4]).
    """
    c0 = F.relu(_convbn3(cost, st, p + ".dres0.0", 1, training, stats))
    c0 = F.relu(_convbn3(c0, st, p + ".dres0.2", 1, training, stats))
    r = F.relu(_convbn3(c0, st, p + ".dres1.0", 1, training, stats))
    cost0 = _convbn3(r, st, p + ".dres1.2", 1, training, stats) + c0

    out1, pre1, post1 = hourglass(cost0, None, None, st, p + ".dres2", training, stats)
    out1 = out1 + cost0
    out2, _pre2, post2 = hourglass(out1, pre1, post1, st, p + ".dres3", training, stats)
    out2 = out2 + cost0
    out3, _pre3, _post3 = hourglass(out2, pre1, post2, st, p + ".dres4", training, stats)
    out3 = out3 + cost0

    def head(x, k):
        y = F.relu(_convbn3(x, st, f"{p}.classif{k}.0", 1, training, stats))
        return _conv3(y, st, f"{p}.classif{k}.2")

    cost1 = head(out1, 1)
    cost2 = head(out2, 2) + cost1
    cost3 = head(out3, 3) + cost2
    return [cost3, cost2, cost1], [out3, out2, out1]


def upsample_cost(cost: torch.Tensor, multiplier: int = 4) -> torch.Tensor:
    """F.interpolate(scale_factor=4, trilinear, align_corners=True) + squeeze(1), modules.py:327-334."""
    return F.interpolate(cost, scale_factor=multiplier, mode="trilinear", align_corners=True).squeeze(1)


def aggregation(cost: torch.Tensor, st: State, p: str, training: bool, stats=None):
    """Full PSMNetHGAggregation.forward (modules.py:310-337): eval keeps only head 3."""
    costs, outs = aggregation_lowres(cost, st, p, training, stats)
    if training:
        return [upsample_cost(c) for c in costs], outs
    return [upsample_cost(costs[0])], [outs[0]]


def regression(cost_full: torch.Tensor, bins: np.ndarray) -> Tuple[torch.Tensor, torch.Tensor]:
    """disp_regression.forward for one head, src/model/stereodpnet/modules.py:352-362.

    cost_full [B, 4*level, H, W] -> (disparity [B,H,W], probability [B,4*level,H,W]).
    """
    prob = F.softmax(cost_full, dim=1)
    d = torch.from_numpy(np.reshape(bins, [1, -1, 1, 1])).to(prob.dtype).to(prob.device)
    return torch.sum(prob * d, 1), prob


# --------------------------------------------------------------------------------------
# 3-D deformable convolution (D3D), restating src/module/dcn3d/src/cuda/deform_im2col_cuda.cuh
# --------------------------------------------------------------------------------------


def _trilinear_gather(x: torch.Tensor, d: torch.Tensor, h: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """dmcn_im2col_bilinear (:26-72): x [B,C,D,H,W]; d,h,w [B,Do,Ho,Wo] float coords -> [B,C,Do,Ho,Wo].

    A corner contributes only if it lies inside the volume; the whole sample is zero unless
    -1 < coord < size on every axis (:248).
    """
    bsz, c, dd, hh, ww = x.shape
    inside = (d > -1) & (h > -1) & (w > -1) & (d < dd) & (h < hh) & (w < ww)
    d0, h0, w0 = torch.floor(d), torch.floor(h), torch.floor(w)
    ld, lh, lw = d - d0, h - h0, w - w0
    d0, h0, w0 = d0.long(), h0.long(), w0.long()
    xf = x.reshape(bsz, c, -1)
    out = x.new_zeros(bsz, c, *d.shape[1:])
    for cd in (0, 1):
        for ch in (0, 1):
            for cw in (0, 1):
                di, hi, wi = d0 + cd, h0 + ch, w0 + cw
                ok = inside & (di >= 0) & (di <= dd - 1) & (hi >= 0) & (hi <= hh - 1) & (wi >= 0) & (wi <= ww - 1)
                wt = (ld if cd else 1 - ld) * (lh if ch else 1 - lh) * (lw if cw else 1 - lw)
                idx = (di.clamp(0, dd - 1) * hh + hi.clamp(0, hh - 1)) * ww + wi.clamp(0, ww - 1)
                g = torch.gather(xf, 2, idx.reshape(bsz, 1, -1).expand(-1, c, -1)).reshape(out.shape)
                out = out + g * (wt * ok.to(x.dtype)).unsqueeze(1)
    return out


def deform_conv3d(x: torch.Tensor, offset: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                  stride: int = 1, pad: int = 1, dil: int = 1) -> torch.Tensor:
    """deform_conv_cuda_forward (src/module/dcn3d/src/cuda/deform_conv_cuda.cu:18-126), groups = 1.

    y[b,o] = bias[o] + sum_{tap,c} W[o,c,tap] * trilinear(x[b,c], p*stride - pad + tap*dil + offset[b, 3*tap + (0,1,2)])
    with offsets ordered (d, h, w) per tap (deform_im2col_cuda.cuh:238-243).  Differentiable through
    autograd (used by the tests to check the hand-written backward).
    """
    bsz, c, dd, hh, ww = x.shape
    co, ci, kd, kh, kw = weight.shape
    assert ci == c
    do = (dd + 2 * pad - (dil * (kd - 1) + 1)) // stride + 1
    ho = (hh + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    wo = (ww + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    assert offset.shape == (bsz, 3 * kd * kh * kw, do, ho, wo), offset.shape
    gd, gh, gw = torch.meshgrid(torch.arange(do, device=x.device, dtype=x.dtype),
                                torch.arange(ho, device=x.device, dtype=x.dtype),
                                torch.arange(wo, device=x.device, dtype=x.dtype), indexing="ij")
    y = x.new_zeros(bsz, co, do, ho, wo)
    tap = 0
    for i in range(kd):
        for j in range(kh):
            for k in range(kw):
                pd = gd * stride - pad + i * dil + offset[:, 3 * tap + 0]
                ph = gh * stride - pad + j * dil + offset[:, 3 * tap + 1]
                pw = gw * stride - pad + k * dil + offset[:, 3 * tap + 2]
                col = _trilinear_gather(x, pd, ph, pw)
                y = y + torch.einsum("oc,bcdhw->bodhw", weight[:, :, i, j, k], col)
                tap += 1
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    return y


def deform_conv_pack(x: torch.Tensor, st: State, p: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """DeformConvPack_dv2.forward with dimension='THW' (src/module/dcn3d/modules/deform_conv.py:323-389)."""
    offset = F.conv3d(x.float(), st[p + ".conv_offset.weight"], st[p + ".conv_offset.bias"], stride=1, padding=1)
    y = deform_conv3d(x.float(), offset.float(), st[p + ".weight"].float(), st[p + ".bias"].float())
    return y, offset


# --------------------------------------------------------------------------------------
# ANM normal branch, src/model/stereodpnet/normal_module.py
# --------------------------------------------------------------------------------------


def disp2depth(disp: torch.Tensor, abvalue: torch.Tensor) -> torch.Tensor:
    """src/utils/geometry.py:21-45: depth = a / (disp - b) with a = abvalue[:,1], b = abvalue[:,0]; NaN/Inf -> 0."""
    a = abvalue[:, 1].view(-1, 1, 1, 1).to(disp.dtype)
    b = abvalue[:, 0].view(-1, 1, 1, 1).to(disp.dtype)
    depth = torch.div(a, disp - b)
    return torch.where(torch.isnan(depth) | torch.isinf(depth), torch.zeros_like(depth), depth)


def anm_select_levels(disp_q: torch.Tensor, crange_t: torch.Tensor, k: int) -> torch.Tensor:
    """sample_with_sort index selection, normal_module.py:130-134: topk of 1/(|c-d|+1e-6), then sort.

    disp_q [B,1,H4,W4]; crange_t [1,D,1,1] -> sorted indices [B,k,H4,W4] (int64).
    Ties (d exactly on a level) are resolved by torch.topk on the running platform.
    """
    diff = torch.abs(crange_t - disp_q)
    _, idx = torch.topk(1.0 / (diff + 1e-6), k=k, dim=1)
    return torch.sort(idx, dim=1)[0]


def anm_coord_volume(disp_sel: torch.Tensor, k_mat: torch.Tensor, abvalue: torch.Tensor) -> torch.Tensor:
    """grid_maker_3d, normal_module.py:80-118 -> [B, Dk, 3, H4, W4].

    K[:, :2] / 4 (:103); K^-1 [u,v,1] (:104-105); scaled by depth(disp) (:109-110); per-sample
    min/max normalisation with +1e-6 (:113-115).
    """
    b, dk, h, w = disp_sel.shape
    xs = torch.arange(0, w).to(k_mat)
    ys = torch.arange(0, h).to(k_mat)
    yg, xg = torch.meshgrid([ys, xs], indexing="ij")
    grid = torch.stack([xg, yg, torch.ones_like(xg)], 0).reshape(1, 3, -1).expand(b, -1, -1)
    kq = k_mat.clone()
    kq[:, :2, :] = kq[:, :2, :] / 4.0
    rays = torch.bmm(torch.inverse(kq), grid).view(b, 3, h, w).to(disp_sel.dtype)
    depth = disp2depth(disp_sel, abvalue)
    vol = rays.unsqueeze(2) * depth.unsqueeze(1)                      # [B,3,Dk,H,W]
    vmin = vol.reshape(b, -1).min(-1)[0].view(b, 1, 1, 1, 1)
    vmax = vol.reshape(b, -1).max(-1)[0].view(b, 1, 1, 1, 1)
    vol = (vol - vmin) / (vmax - vmin + 1e-6)
    return vol.permute(0, 2, 1, 3, 4).contiguous()


def anm_forward(out3: torch.Tensor, disp_full: torch.Tensor, k_mat: torch.Tensor, abvalue: torch.Tensor,
                st: State, p: str, crange: Sequence[float], training: bool, dsample_num: int = 4,
                stats=None, return_aux: bool = False, use_deform: bool = True, use_sampling: bool = True):
    """ANM.forward for one (cost, disparity) pair, normal_module.py:140-194.

    out3 [B,C,D,H4,W4]; disp_full [B,H,W] -> normal [B,3,H,W] in [-1,1].
    ``use_sampling=False`` (:159-163) keeps all D cost slices with the level disparities; ``use_deform=False`` (:52-56,181-183)
    replaces the two deformable layers by ``original_conv`` = two convbn_3d + ReLU.
    """
    b, c, d, h, w = out3.shape
    cost = out3.permute(0, 2, 1, 3, 4)                                # b d c h w
    disp_q = F.interpolate(disp_full.unsqueeze(1), scale_factor=0.25, mode="nearest") * 0.25
    crange_t = torch.as_tensor(np.asarray(crange), dtype=torch.float32, device=out3.device).view(1, -1, 1, 1)
    if use_sampling:
        idx = anm_select_levels(disp_q, crange_t, dsample_num)
        sel_cost = torch.gather(cost, 1, idx.unsqueeze(2).expand(-1, -1, c, -1, -1))
        sel_disp = torch.gather(crange_t.expand(b, d, h, w), 1, idx)
    else:
        dsample_num = d
        idx = torch.arange(d, device=out3.device).view(1, d, 1, 1).expand(b, d, h, w)
        sel_cost, sel_disp = cost, crange_t.expand(b, d, h, w).to(cost.dtype)
    coord = anm_coord_volume(sel_disp, k_mat, abvalue)
    fv = torch.cat([sel_cost, coord.to(sel_cost.dtype)], dim=2).permute(0, 2, 1, 3, 4).contiguous()  # b c+3 k h w

    if use_deform:
        f1, off1 = deform_conv_pack(fv, st, p + ".deform_conv1")
        f1 = F.relu(_bn(f1, st, p + ".act1.0", training, stats_out=stats))
        f2, off2 = deform_conv_pack(f1, st, p + ".deform_conv2")
        f2 = F.relu(_bn(f2, st, p + ".act2.0", training, stats_out=stats))
    else:
        off1 = off2 = None
        f1 = F.relu(_convbn3(fv, st, p + ".original_conv.0", 1, training, stats))
        f2 = F.relu(_convbn3(f1, st, p + ".original_conv.2", 1, training, stats))

    feat = f2.permute(0, 2, 1, 3, 4).reshape(b * dsample_num, f2.shape[1], h, w)
    for i, dil in enumerate((1, 2, 4, 8, 1, 1)):                      # convtext stack, normal_module.py:59-66
        feat = F.leaky_relu(F.conv2d(feat, st[f"{p}.n_convs.{i}.0.weight"], None, padding=dil, dilation=dil), 0.1)
    feat = torch.sigmoid(F.interpolate(feat, scale_factor=4, mode="bilinear", align_corners=True))
    normal = feat.reshape(b, dsample_num, 3, feat.shape[-2], feat.shape[-1]).mean(dim=1) * 2.0 - 1.0
    if return_aux:
        return normal, {"idx": idx, "fv": fv, "off1": off1, "off2": off2, "f1": f1, "f2": f2}
    return normal


# --------------------------------------------------------------------------------------
# 2-D encoders (adjacent to the hot path; needed for whole-model parity and the CPU baseline)
# --------------------------------------------------------------------------------------


def _dpblock(x, st: State, p: str, ratio_s: int, training: bool, stats=None):
    """DPBlock.forward, src/model/stereodpnet/modules.py:21-54."""
    def cb(inp, key, stride=1, pad=1, dil=1):
        return _convbn2(inp, st, key, stride, pad, dil, training, stats)

    o1 = F.prelu(cb(x, p + ".conv1.0"), st[p + ".conv1.1.weight"])
    o2 = F.prelu(cb(o1, p + ".conv2.0"), st[p + ".conv2.1.weight"])
    o2 = torch.cat([cb(o2, f"{p}.conv_dilate.{i}", 1, 2 * i + 1, 2 * i + 1) for i in range(3)], dim=1)
    o2 = cb(o2, p + ".conv3")
    out = F.prelu(o2 + o1, st[p + ".prelu.weight"])
    out = F.prelu(cb(out, p + ".conv4.0", ratio_s, ratio_s, 2), st[p + ".conv4.1.weight"])
    # depthwise_separable_conv (src/module/asm/basics.py:39-60): depthwise 3x3 pad 1, pointwise, BN, PReLU
    dw = F.conv2d(out, st[p + ".conv5.depthwise.weight"], None, padding=1, groups=out.shape[1])
    pw = F.conv2d(dw, st[p + ".conv5.pointwise.weight"], None)
    pw = F.prelu(_bn(pw, st, p + ".conv5.bn", training, stats_out=stats), st[p + ".conv5.prelu.weight"])
    skip = F.conv2d(x, st[p + ".conv_skip.weight"], st[p + ".conv_skip.bias"], stride=ratio_s)
    return pw + skip


def sdp_encoder(img: torch.Tensor, st: State, p: str, training: bool, block_stack: int = 1, stats=None):
    """feature_extraction.forward (StereoDPNet), src/model/stereodpnet/modules.py:93-134 -> [B,32,H/4,W/4]."""
    x = img
    for i, s in zip((0, 2, 4), (2, 1, 1)):
        x = F.relu(_convbn2(x, st, f"{p}.firstconv.{i}", s, 1, 1, training, stats))
    o1 = _dpblock(x, st, p + ".block1", 2, training, stats)
    o2 = o1
    for i in range(block_stack):
        o2 = _dpblock(o2, st, f"{p}.interblock1.{i}", 1, training, stats)
    o2 = _dpblock(o2, st, p + ".block2", 2, training, stats)
    o3 = o2
    for i in range(block_stack):
        o3 = _dpblock(o3, st, f"{p}.interblock2.{i}", 1, training, stats)
    o3 = _dpblock(o3, st, p + ".block3", 2, training, stats)

    # torchvision.ops.FeaturePyramidNetwork (0.26 key layout: inner_blocks.N.0 / layer_blocks.N.0), top-down
    feats = [o1, o2, o3]

    def inner(i):
        return F.conv2d(feats[i], st[f"{p}.fpn.inner_blocks.{i}.0.weight"], st[f"{p}.fpn.inner_blocks.{i}.0.bias"])

    def layer(t, i):
        return F.conv2d(t, st[f"{p}.fpn.layer_blocks.{i}.0.weight"], st[f"{p}.fpn.layer_blocks.{i}.0.bias"], padding=1)

    last = inner(2)
    res = [None, None, layer(last, 2)]
    for i in (1, 0):
        lat = inner(i)
        last = lat + F.interpolate(last, size=lat.shape[-2:], mode="nearest")
        res[i] = layer(last, i)
    s1 = F.interpolate(res[1], scale_factor=2, mode="bilinear", align_corners=True)
    s2 = F.interpolate(res[2], scale_factor=4, mode="bilinear", align_corners=True)
    f = torch.cat([res[0], s1, s2], dim=1)
    f = F.relu(_convbn2(f, st, p + ".lastconv.0", 1, 1, 1, training, stats))
    f = F.relu(_convbn2(f, st, p + ".lastconv.2", 1, 1, 1, training, stats))
    return f


def _psm_layer(x, st: State, p: str, blocks: int, stride: int, dil: int, has_down: bool, training: bool, stats=None):
    """feature_extraction._make_layer + BasicBlock.forward, src/model/psmnet/modules.py:14-34,129-143."""
    for i in range(blocks):
        s = stride if i == 0 else 1
        q = f"{p}.{i}"
        o = F.relu(_convbn2(x, st, q + ".conv1.0", s, 1, dil, training, stats))
        o = _convbn2(o, st, q + ".conv2", 1, 1, dil, training, stats)
        if i == 0 and has_down:
            x = _bn(F.conv2d(x, st[q + ".downsample.0.weight"], None, stride=s), st, q + ".downsample.1", training,
                    stats_out=stats)
        x = o + x
    return x


def psm_encoder(img: torch.Tensor, st: State, p: str, training: bool, inplanes: int = 32, stats=None,
                branch_align_corners: bool = True):
    """feature_extraction.forward (PSMNet SPP), src/model/psmnet/modules.py:145-171 -> [B,32,H/4,W/4].
    NNet's copy of the encoder (src/model/nnet/modules.py:46-139) differs only in up-sampling its four pooled branches with
    align_corners=False (:115-124): ``branch_align_corners=False``."""
    x = img
    for i, s in zip((0, 2, 4), (2, 1, 1)):
        x = F.relu(_convbn2(x, st, f"{p}.firstconv.{i}", s, 1, 1, training, stats))
    x = _psm_layer(x, st, p + ".layer1", 3, 1, 1, False, training, stats)
    raw = _psm_layer(x, st, p + ".layer2", inplanes // 2, 2, 1, True, training, stats)
    x = _psm_layer(raw, st, p + ".layer3", 3, 1, 1, True, training, stats)
    skip = _psm_layer(x, st, p + ".layer4", 3, 1, 2, False, training, stats)
    branches = []
    for name, k in (("branch1", inplanes * 2), ("branch2", inplanes), ("branch3", inplanes // 2),
                    ("branch4", inplanes // 4)):
        y = F.avg_pool2d(skip, (k, k), stride=(k, k))
        y = F.relu(_convbn2(y, st, f"{p}.{name}.1", 1, 0, 1, training, stats))
        branches.append(F.interpolate(y, size=skip.shape[-2:], mode="bilinear", align_corners=branch_align_corners))
    f = torch.cat([raw, skip, branches[3], branches[2], branches[1], branches[0]], dim=1)
    f = F.relu(_convbn2(f, st, p + ".lastconv.0", 1, 1, 1, training, stats))
    return F.conv2d(f, st[p + ".lastconv.2.weight"], None)


# --------------------------------------------------------------------------------------
# losses consuming the path
# --------------------------------------------------------------------------------------


def smooth_l1_multi(pred: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor],
                    weights: Sequence[float]) -> torch.Tensor:
    """SMOOTHL1Loss.forward ('given' conversion, disparity target), src/loss/depth/smoothL1.py:15-49."""
    n = pred.shape[1]
    ws = [1.0] if n == 1 else list(weights)
    assert len(ws) == n
    if mask is not None:
        m = mask > 0
        return sum(ws[i] * F.smooth_l1_loss(pred[:, i][m], gt[m]) for i in range(n))
    return sum(ws[i] * F.smooth_l1_loss(pred[:, i], gt) for i in range(n))


def cosine_normal_loss(pred: torch.Tensor, gt: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """COSINELoss.forward, masked branch, single prediction, src/loss/normal/cosine.py:35-55.

    pred [B,1,3,H,W], gt [B,3,H,W].  Quirk kept: the similarity is element-wise over the 3 components
    (no sum), so the loss is mean(1 - n_pred*n_gt / (|n_pred||n_gt|)) (:18-26).
    """
    m = mask > 0
    p = pred.permute(0, 3, 4, 1, 2)[m]                                 # [N, 1, 3]
    g = gt.permute(0, 2, 3, 1)[m]                                      # [N, 3]
    p = p / torch.norm(p, p=2, dim=-1, keepdim=True).clamp_min(1e-6)
    g = g / torch.norm(g, p=2, dim=-1, keepdim=True).clamp_min(1e-6)
    a, bb = p[:, 0], g
    den = (torch.norm(a, p=2, dim=-1, keepdim=True) * torch.norm(bb, p=2, dim=-1, keepdim=True)).clamp_min(1e-6)
    sim = ((a * bb) / den).clamp(min=-1.0, max=1.0)
    return torch.mean(1.0 - sim)


# --------------------------------------------------------------------------------------
# whole models
# --------------------------------------------------------------------------------------

SDP_CFG = dict(mindisp=-4, maxdisp=12, level=8, inplanes=32, block_stack=1, dsample_num=4, use_deform=True, use_sampling=True,
               loss_weight=(1.0, 0.7, 0.5), lambdas=(1.0, 1.0))
PSM_CFG = dict(mindisp=-4, maxdisp=12, level=8, inplanes=32, cost_volume="psmnet", group_num=40,
               loss_weight=(1.0, 0.7, 0.5), lambdas=(1.0,))


def _pick_ref_target(batch: dict, flip_lr: bool, training: bool):
    """src/model/stereodpnet/mainmodel.py:70-83."""
    if "groupname" in batch and not training:
        swap = batch["groupname"][0] == "2020-2-9_group20"
    else:
        swap = flip_lr
    return (batch["right"], batch["left"]) if swap else (batch["left"], batch["right"])


def stereodpnet_forward(batch: dict, st: State, training: bool, cfg: dict = SDP_CFG, flip_lr: bool = True,
                        cached_first_level: bool = True, predict_normal: bool = True, stages: Optional[dict] = None,
                        stats: Optional[dict] = None):
    """STEREODPNET.forward, src/model/stereodpnet/mainmodel.py:67-111.

    ``stats`` (train mode only) collects each BatchNorm's batch mean / biased variance, keyed by module prefix.
    """
    crange = cost_range(cfg["mindisp"], cfg["maxdisp"], cfg["level"])
    bins = disparity_bins(cfg["mindisp"], cfg["maxdisp"], cfg["level"])
    ref_img, tgt_img = _pick_ref_target(batch, flip_lr, training)
    ref = sdp_encoder(ref_img, st, "feature_extraction", training, cfg["block_stack"], stats)
    tgt = sdp_encoder(tgt_img, st, "feature_extraction", training, cfg["block_stack"], stats)
    vol = sdp_cost_volume(ref, tgt, st, "cost_volume", crange, training, cached_first_level=cached_first_level,
                          stats=stats)
    costs, outs = aggregation(vol, st, "aggregation", training, stats)
    disps, probs = zip(*[regression(c, bins) for c in costs])
    normal = None
    if predict_normal:
        normal = anm_forward(outs[0], disps[0], batch["K"], batch["abvalue"], st, "normal_estimator", crange,
                             training, cfg["dsample_num"], stats=stats, use_deform=cfg.get("use_deform", True),
                             use_sampling=cfg.get("use_sampling", True)).unsqueeze(1)
    res = {"pred_depth": torch.stack(disps, 1), "prob_depth": torch.stack(probs, 1), "pred_normal": normal,
           "ref_feature": ref.max(1)[0]}
    if stages is not None:
        stages.update(ref=ref, tgt=tgt, volume=vol, costs=costs, outs=outs)
    if training and "disp" in batch:
        l1 = smooth_l1_multi(res["pred_depth"], batch["disp"], batch.get("mask"), cfg["loss_weight"])
        res["smoothL1_loss"] = l1
        total = cfg["lambdas"][0] * l1
        if predict_normal:
            lc = cosine_normal_loss(res["pred_normal"], batch["normal"], batch["mask"])
            res["cosine_loss"] = lc
            total = total + cfg["lambdas"][1] * lc
        res["final_loss"] = total
    return res


def psmnet_forward(batch: dict, st: State, training: bool, cfg: dict = PSM_CFG, flip_lr: bool = True,
                   stages: Optional[dict] = None, stats: Optional[dict] = None):
    """PSMNET.forward, src/model/psmnet/mainmodel.py:74-113."""
    crange = cost_range(cfg["mindisp"], cfg["maxdisp"], cfg["level"])
    bins = disparity_bins(cfg["mindisp"], cfg["maxdisp"], cfg["level"])
    ref_img, tgt_img = _pick_ref_target(batch, flip_lr, training)
    ref = psm_encoder(ref_img, st, "feature_extraction", training, cfg["inplanes"], stats)
    tgt = psm_encoder(tgt_img, st, "feature_extraction", training, cfg["inplanes"], stats)
    if cfg["cost_volume"] == "psmnet":
        vol = psm_concat_volume(ref, tgt, crange)
    elif cfg["cost_volume"] == "gwcnet":
        vol = torch.cat([psm_concat_volume(ref, tgt, crange), psm_gwc_volume(ref, tgt, crange, cfg["group_num"])], 1)
    else:
        raise NotImplementedError(cfg["cost_volume"])
    costs, outs = aggregation(vol, st, "aggregation", training, stats)
    disps, probs = zip(*[regression(c, bins) for c in costs])
    res = {"pred_depth": torch.stack(disps, 1), "prob_depth": torch.stack(probs, 1), "ref_feature": ref.max(1)[0]}
    if stages is not None:
        stages.update(ref=ref, tgt=tgt, volume=vol, costs=costs, outs=outs)
    if training and "disp" in batch:
        l1 = smooth_l1_multi(res["pred_depth"], batch["disp"], batch.get("mask"), cfg["loss_weight"])
        res["smoothL1_loss"] = l1
        res["final_loss"] = cfg["lambdas"][0] * l1
    return res


STN_CFG = dict(mindisp=-4, maxdisp=12, k=3, loss_weight=(1.0, 1.0), lambdas=(1.0,))


def _stn_block(x, st: State, p: str, dil: int, training: bool, stats=None):
    """BasicBlock.forward of StereoNet, src/model/stereonet/modules.py:10-29: ``x + LeakyReLU_0.2(convbn(x))`` -- the block's
    conv2 is constructed (its weights are in the state dict) but never applied (:23, the reference's own "bug?" note :27)."""
    return x + F.leaky_relu(_convbn2(x, st, p + ".conv1.0", 1, 1, dil, training, stats), 0.2)


def stn_encoder(img: torch.Tensor, st: State, p: str, training: bool, k: int = 3, stats=None):
    """FeatureExtraction.forward, src/model/stereonet/modules.py:32-61: k 5x5 stride-2 convs (bias, no activation), six
    BasicBlocks, one 3x3 conv with bias -> [B,32,H/2^k,W/2^k]."""
    x = img
    for i in range(k):
        x = _conv2(x, st, f"{p}.downsample.{i}", stride=2, pad=2, bias=True)
    for i in range(6):
        x = _stn_block(x, st, f"{p}.residual_blocks.{i}", 1, training, stats)
    return _conv2(x, st, p + ".conv_alone", bias=True)


def stn_refine(low_disp: torch.Tensor, rgb: torch.Tensor, st: State, p: str, training: bool, stats=None):
    """EdgeAwareRefinement.forward, src/model/stereonet/modules.py:64-96: bilinear (align_corners=False) upsample of the
    low-resolution disparity to the image size, x8 when the size ratio is >= 1.5, 4 -> 32 convbn + LeakyReLU(0.2), six dilated
    BasicBlocks (1, 2, 4, 8, 1, 1), 32 -> 1 conv with bias, ReLU(upsampled + residual)."""
    up = F.interpolate(low_disp.unsqueeze(1), size=rgb.shape[-2:], mode="bilinear", align_corners=False)
    if rgb.shape[-1] / low_disp.shape[-1] >= 1.5:
        up = up * 8
    x = F.leaky_relu(_convbn2(torch.cat([up, rgb], 1), st, p + ".conv2d_feature.0", 1, 1, 1, training, stats), 0.2)
    for i, dil in enumerate((1, 2, 4, 8, 1, 1)):
        x = _stn_block(x, st, f"{p}.residual_astrous_blocks.{i}", dil, training, stats)
    return F.relu((up + _conv2(x, st, p + ".conv2d_out", bias=True)).squeeze(1))


def stereonet_forward(batch: dict, st: State, training: bool, cfg: dict = STN_CFG, flip_lr: bool = True,
                      stages: Optional[dict] = None, stats: Optional[dict] = None):
    """STEREONET.forward, src/model/stereonet/mainmodel.py:80-152: features at 1/2^k, DIFFERENCE volume over int(costrange)
    row shifts (:100-114; costrange keeps the /4 of the other models, :37-40), four convbn_3d + LeakyReLU(0.2) and a 32 -> 1
    conv with bias (:43-51,117-120), softmax regression over the 2^k levels WITHOUT up-sampling (modules.py:99-120; bins
    arange(level) * (maxdisp-mindisp)/level + mindisp), edge-aware refinement on batch['right'] (:125-126; always the right
    image, whatever flip_lr says), level 0 scaled by W / w and bilinearly up-sampled (:128-137)."""
    level = int(math.pow(2, cfg["k"]))
    crange = cost_range(cfg["mindisp"], cfg["maxdisp"], level)
    bins = np.arange(level, dtype=np.float64) * ((cfg["maxdisp"] - cfg["mindisp"]) / float(level)) + cfg["mindisp"]
    ref_img, tgt_img = _pick_ref_target(batch, flip_lr, training)
    ref = stn_encoder(ref_img, st, "feature_extraction", training, cfg["k"], stats)
    tgt = stn_encoder(tgt_img, st, "feature_extraction", training, cfg["k"], stats)
    vol = diff_volume(ref, tgt, crange)
    x = vol
    for i in range(4):
        x = F.leaky_relu(_convbn3(x, st, f"filter.{i}.0", 1, training, stats), 0.2)
    cost = _conv3(x, st, "conv3d_alone", bias=True).squeeze(1)
    disp, prob = regression(cost, bins)
    right = batch["right"]
    refined = stn_refine(disp, right, st, "edge_aware_refinements.0", training, stats)
    coarse = F.interpolate((disp * (right.shape[-1] / disp.shape[-1])).unsqueeze(1), size=right.shape[-2:], mode="bilinear",
                           align_corners=False).squeeze(1)
    res = {"pred_depth": torch.stack([coarse, refined], 1), "prob_depth": prob.unsqueeze(1), "ref_feature": ref.max(1)[0]}
    if stages is not None:
        stages.update(ref=ref, tgt=tgt, volume=vol, cost=cost, disp_low=disp)
    if training and "disp" in batch:
        l1 = smooth_l1_multi(res["pred_depth"], batch["disp"], batch.get("mask"), cfg["loss_weight"])
        res["smoothL1_loss"] = l1
        res["final_loss"] = cfg["lambdas"][0] * l1
    return res


NNET_CFG = dict(mindisp=-4, maxdisp=12, level=8, inplanes=32, loss_weight=(1.0, 1.0), lambdas=(1.0, 1.0))


def _convtext(x, st: State, key: str, dil: int):
    """convtext(), src/model/nnet/modules.py:37-43: bias-free 3x3 conv with padding = dilation, LeakyReLU(0.1)."""
    return F.leaky_relu(_conv2(x, st, key + ".0", pad=dil, dil=dil), 0.1)


NNET_CONTEXT_DILATIONS = (1, 2, 4, 8, 16, 1, 1)       # both `convs` (mainmodel.py:48-56) and `n_convs` (normal_module_.py:34-42)


def nnet_context(x, st: State, p: str):
    for i, dil in enumerate(NNET_CONTEXT_DILATIONS):
        x = _convtext(x, st, f"{p}.{i}", dil)
    return x


def nnet_normal_module(cost_in: torch.Tensor, k_mat: torch.Tensor, abvalue: torch.Tensor, st: State, p: str,
                       crange: Sequence[float], training: bool, stats=None) -> torch.Tensor:
    """NormalModule.forward, src/model/nnet/normal_module_.py:84-118: world-coordinate volume of the 8 cost levels themselves
    (grid_maker_3d :46-82, the arithmetic of ANM's -- K[:2]/4, K^-1 [u,v,1], depth = a/(d-b), per-sample min/max normalisation),
    cat with cost_in -> 67 channels, two convbn_3d + ReLU, three (2,3,3)/stride (2,1,1)/pad (0,1,1) convbn_3d + ReLU that fold the
    8 levels to 1, the 7 dilated `n_convs` per remaining slice (summed), bilinear x4 (align_corners=True), F.normalize."""
    b, ch, d, h, w = cost_in.shape
    disp_range = torch.tensor(np.asarray(crange), dtype=torch.float32).view(1, -1, 1, 1).expand(b, -1, h, w).to(cost_in)
    wc = anm_coord_volume(disp_range, k_mat, abvalue).permute(0, 2, 1, 3, 4)            # [B,3,D,h,w]
    x = torch.cat([wc, cost_in], 1)
    x = F.relu(_convbn3(x, st, p + ".wc0.0", 1, training, stats))
    x = F.relu(_convbn3(x, st, p + ".wc0.2", 1, training, stats))
    for name in ("pool1", "pool2", "pool3"):
        y = F.conv3d(x, st[f"{p}.{name}.0.0.weight"], None, stride=(2, 1, 1), padding=(0, 1, 1))
        x = F.relu(_bn(y, st, f"{p}.{name}.0.1", training, stats_out=stats))
    nmap = sum(nnet_context(x[:, :, i], st, p + ".n_convs") for i in range(x.shape[2]))
    nmap = F.interpolate(nmap, scale_factor=4, mode="bilinear", align_corners=True)
    return F.normalize(nmap, dim=1)


def nnet_forward(batch: dict, st: State, training: bool, cfg: dict = NNET_CFG, flip_lr: bool = True,
                 stages: Optional[dict] = None, stats: Optional[dict] = None):
    """NNET.forward, src/model/nnet/mainmodel.py:113-177: PSMNet-style SPP encoder, concat volume over int(costrange) row shifts
    (modules.py:169-188), dres0 (two convbn_3d + ReLU), four residual pairs dres1-4 (no ReLU after the add), classify (convbn_3d,
    ReLU, 32 -> 1 conv), a per-level 2-D context refinement `convs` on cat(ref features, cost slice) added back to the slice
    (:142-146), BOTH volumes up-sampled x4 trilinear with align_corners=False (:149-151) and regressed -> pred_depth [B,2,H,W]
    (raw, refined); the normal module takes cat(dres0 output, dres4 output) (:141,155)."""
    crange = cost_range(cfg["mindisp"], cfg["maxdisp"], cfg["level"])
    bins = disparity_bins(cfg["mindisp"], cfg["maxdisp"], cfg["level"])
    ref_img, tgt_img = _pick_ref_target(batch, flip_lr, training)
    ref = psm_encoder(ref_img, st, "feature_extraction", training, cfg["inplanes"], stats, branch_align_corners=False)
    tgt = psm_encoder(tgt_img, st, "feature_extraction", training, cfg["inplanes"], stats, branch_align_corners=False)
    vol = psm_concat_volume(ref, tgt, crange)
    c0 = F.relu(_convbn3(vol, st, "dres0.0", 1, training, stats))
    c0 = F.relu(_convbn3(c0, st, "dres0.2", 1, training, stats))
    cost_in0 = c0
    for k in (1, 2, 3, 4):
        c0 = _convbn3(F.relu(_convbn3(c0, st, f"dres{k}.0", 1, training, stats)), st, f"dres{k}.2", 1, training, stats) + c0
    cost_in = torch.cat([cost_in0, c0], 1)
    costs = _conv3(F.relu(_convbn3(c0, st, "classify.0", 1, training, stats)), st, "classify.2")       # [B,1,D,h,w]
    refined = torch.stack([nnet_context(torch.cat([ref, costs[:, :, i]], 1), st, "convs") + costs[:, :, i]
                           for i in range(costs.shape[2])], 2)
    up = lambda c: F.interpolate(c, scale_factor=4, mode="trilinear", align_corners=False).squeeze(1)
    disps, probs = zip(*[regression(up(c), bins) for c in (costs, refined)])
    normal = nnet_normal_module(cost_in, batch["K"], batch["abvalue"], st, "normal_module", crange, training, stats).unsqueeze(1)
    res = {"pred_depth": torch.stack(disps, 1), "prob_depth": torch.stack(probs, 1), "pred_normal": normal,
           "ref_feature": ref.max(1)[0]}
    if stages is not None:
        stages.update(ref=ref, tgt=tgt, volume=vol, cost_in=cost_in, costs=costs, refined=refined)
    if training and "disp" in batch:
        l1 = smooth_l1_multi(res["pred_depth"], batch["disp"], batch.get("mask"), cfg["loss_weight"])
        lc = cosine_normal_loss(res["pred_normal"], batch["normal"], batch["mask"])
        res.update(smoothL1_loss=l1, cosine_loss=lc, final_loss=cfg["lambdas"][0] * l1 + cfg["lambdas"][1] * lc)
    return res


def calibrate_running_stats(st: State, stats: dict) -> State:
    """Copy of ``st`` whose BatchNorm running statistics are the batch statistics collected in ``stats``.

    Test helper: with random weights and default running stats (0, 1) an eval-mode forward saturates the
    soft-argmin (SURVEY.md section 8c caveats); one train-mode oracle pass + this gives a well-conditioned
    eval-mode state.  Not part of the reference.
    """
    out = dict(st)
    for prefix, (mean, var) in stats.items():
        out[prefix + ".running_mean"] = mean.clone()
        out[prefix + ".running_var"] = var.clone()
    return out
