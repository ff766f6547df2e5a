"""Tensor-level wrappers over the C ABI (include/dpf_sm100.h).

PyTorch is used only for device memory and streams: every function takes CUDA tensors, passes raw pointers and the
current stream to libdpf_sm100.so, and returns tensors it allocated.  No function here has a CPU or eager fallback.

Layouts: features [B,H4,W4,C] bf16; volumes / 3-D activations [B,D,H,W,C] bf16; costs [B,D,H4,W4] fp32;
disparity [B,H,W] fp32.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ConvArgs, check

KIND_3x3x3, KIND_S2, KIND_T2, KIND_1x3x3, KIND_1x1x1 = 0, 1, 2, 3, 4
_MODES = {"concat": 0, "diff": 1, "gwc": 2}


_checked = False


def lib():
    """The loaded library; the sm_100a device check (a slow cudaGetDeviceProperties) runs once per process."""
    global _checked
    if not _checked:
        _lib.load(check_device=True)
        _checked = True
    return _lib.load()


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.DpfError(f"{name}: expected a CUDA tensor (the hot path has no CPU implementation)")
    if t.dtype != dtype:
        raise _lib.DpfError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.DpfError(f"{name}: expected a contiguous tensor")


def launch_count() -> int:
    return int(_lib.load().dpf_launch_count())


# ----------------------------------------------------------------------------------------------------------
# cost volume
# ----------------------------------------------------------------------------------------------------------
def costvol_channels(mode: str, c: int, groups: int) -> int:
    return {"concat": 2 * c, "diff": c, "gwc": groups}[mode]


def costvol_fwd(ref: torch.Tensor, tgt: torch.Tensor, shifts: Sequence[int], mode: str = "concat", groups: int = 0) -> torch.Tensor:
    _req(ref, torch.bfloat16, "ref"); _req(tgt, torch.bfloat16, "tgt")
    b, h, w, c = ref.shape
    d = len(shifts)
    vol = torch.empty(b, d, h, w, costvol_channels(mode, c, groups), device=ref.device, dtype=torch.bfloat16)
    sh = (C.c_int * d)(*[int(s) for s in shifts])
    check(lib().dpf_costvol_fwd(_MODES[mode], _p(ref), _p(tgt), _p(vol), b, h, w, c, d, groups, sh, _stream()), "dpf_costvol_fwd")
    return vol


def costvol_bwd(ref, tgt, dvol: torch.Tensor, shifts: Sequence[int], mode: str = "concat", groups: int = 0):
    _req(dvol, torch.bfloat16, "dvol")
    b, d, h, w, _ = dvol.shape
    c = ref.shape[-1]
    dref = torch.empty(b, h, w, c, device=dvol.device, dtype=torch.bfloat16)
    dtgt = torch.empty_like(dref)
    sh = (C.c_int * d)(*[int(s) for s in shifts])
    check(lib().dpf_costvol_bwd(_MODES[mode], _p(ref), _p(tgt), _p(dvol), _p(dref), _p(dtgt), b, h, w, c, d, groups, sh,
                                _stream()), "dpf_costvol_bwd")
    return dref, dtgt


# ----------------------------------------------------------------------------------------------------------
# tensor-core convolution
# ----------------------------------------------------------------------------------------------------------
def npad_for(cout: int) -> int:
    return 16 if cout <= 16 else (32 if cout <= 32 else 64)


# Transposed kind (KIND_T2 = 2): weight-slot order of the fused groups, as dpf_conv3d_fwd builds its tap table (conv3d_tc.cu):
# out[2q + r] += in[q + o] * W[k] per dimension with r = 0 -> (k 1, o 0); r = 1 -> (k 0, o 1), (k 2, o 0).  For every (output-plane
# parity rd, depth tap) and every in-plane input shift (oh, ow), runs of adjacent parity classes cls = 2 rh + rw share one MMA:
_T2_DEPTH = ((0, 1), (1, 0), (1, 2))                          # (rd, kd) in issue order
_T2_RUNS = ((0, 0, (0, 1, 2, 3)), (0, 1, (1,)), (0, 1, (3,)), (1, 0, (2, 3)), (1, 1, (3,)))     # (oh, ow, classes of the run)


def _t2_k(r: int, o: int) -> int:
    return 1 if r == 0 else (0 if o == 1 else 2)


def fuse_t2_weight(wp: torch.Tensor) -> torch.Tensor:
    """Plain packed weights [27][Cin/8][Npad][8] -> the transposed kind's fused-group layout (same size, same shape): 15 groups of
    1, 2 or 4 tap matrices, each stored as ONE operand [Cin/8][ncls * Npad][8]."""
    groups = []
    for _rd, kd in _T2_DEPTH:
        for oh, ow, classes in _T2_RUNS:
            taps = [(kd * 3 + _t2_k(c >> 1, oh)) * 3 + _t2_k(c & 1, ow) for c in classes]
            groups.append(torch.cat([wp[t] for t in taps], 1).reshape(-1))          # [Cin/8][ncls*Npad][8]
    return torch.cat(groups).reshape(wp.shape).contiguous()


def pack_conv_weight(w: torch.Tensor, cin_pad: Optional[int] = None, transposed: bool = False, kind: Optional[int] = None) -> torch.Tensor:
    """nn.Conv3d weight [Cout,Cin,kd,kh,kw] (or ConvTranspose3d [Cin,Cout,...]) -> packed bf16 [taps][Cin/8][Npad][8].

    Tap order is (kd, kh, kw) row-major; output channels are zero-padded to Npad in {16,32,64}, input channels to cin_pad.
    kind = KIND_T2: the fused-group layout the transposed kind of dpf_conv3d_fwd expects (fuse_t2_weight).
    """
    if kind == KIND_T2:
        return fuse_t2_weight(pack_conv_weight(w, cin_pad, transposed))
    if transposed:
        w = w.transpose(0, 1)
    cout, cin = w.shape[:2]
    cin_pad = cin_pad or cin
    assert cin_pad % 8 == 0 and cin_pad >= cin
    taps = w.shape[2] * w.shape[3] * w.shape[4]
    npad = npad_for(cout)
    wt = w.detach().float().permute(2, 3, 4, 1, 0).reshape(taps, cin, cout)
    buf = torch.zeros(taps, cin_pad, npad, device=w.device, dtype=torch.float32)
    buf[:, :cin, :cout] = wt
    return buf.reshape(taps, cin_pad // 8, 8, npad).permute(0, 1, 3, 2).contiguous().to(torch.bfloat16)


# Opt-in per-launch timing (bench.py sets KERNEL_TIMING = {} around its stage passes): CUDA events on the launching stream
# around single launches, keyed by kernel + shape class; value = list of (start, end, algorithmic work, unit).
KERNEL_TIMING = None


def _timing_begin():
    if KERNEL_TIMING is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _timing_end(e0, key, work, unit):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    KERNEL_TIMING.setdefault(key, []).append((e0, e1, float(work), unit))


def conv3d(x: torch.Tensor, w_packed: torch.Tensor, kind: int, cout: int, scale: Optional[torch.Tensor] = None,
           shift: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, relu: bool = False,
           out: Optional[torch.Tensor] = None, out_f32: bool = False, y_coff: int = 0, cin: Optional[int] = None,
           x_coff: int = 0, res_pre: bool = False, slope: float = 0.0) -> torch.Tensor:
    """One launch: y = relu?(conv(x) * scale + shift + residual); x [B,D,H,W,Cx] bf16 -> y [B,Do,Ho,Wo,Cstride].
    slope: negative-side slope of the activation (LeakyReLU; kd-fused 3x3x3 stride-1 layers with Cout <= 32 only).

    `cin`/`x_coff` select a channel window of x; `y_coff` a channel offset of the output tensor `out`.
    """
    _req(x, torch.bfloat16, "x"); _req(w_packed, torch.bfloat16, "w_packed")
    b, d, h, w, cx = x.shape
    cin = cin or cx
    if kind in (KIND_3x3x3, KIND_1x3x3, KIND_1x1x1):
        do, ho, wo = d, h, w
    elif kind == KIND_S2:
        do, ho, wo = (d + 1) // 2, (h + 1) // 2, (w + 1) // 2
    else:
        do, ho, wo = 2 * d, 2 * h, 2 * w
    if out is None:
        out = torch.empty(b, do, ho, wo, cout, device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    assert out.shape[:4] == (b, do, ho, wo) and out.is_contiguous()
    out_f32 = out.dtype == torch.float32
    if residual is not None:
        assert residual.shape == out.shape and residual.dtype == out.dtype and residual.is_contiguous()
    for t, n in ((scale, "scale"), (shift, "shift")):
        if t is not None:
            _req(t, torch.float32, n)
            assert t.numel() == cout
    a = ConvArgs(kind=kind, B=b, D=d, H=h, W=w, Cin=cin, Cout=cout, x=x.data_ptr(), w=w_packed.data_ptr(), y=out.data_ptr(),
                 y_f32=int(out_f32), y_cstride=out.shape[-1], y_coff=y_coff,
                 scale=scale.data_ptr() if scale is not None else None,
                 shift=shift.data_ptr() if shift is not None else None,
                 residual=residual.data_ptr() if residual is not None else None, relu=int(relu), stats=None,
                 x_cstride=cx, x_coff=x_coff, res_pre=int(res_pre), slope=float(slope))
    tm = _timing_begin()
    check(lib().dpf_conv3d_fwd(C.byref(a), _stream()), "dpf_conv3d_fwd")
    vox = b * (d * h * w if kind == KIND_T2 else do * ho * wo)       # transposed: every input voxel meets all 27 taps
    _timing_end(tm, f"conv3d kind{kind} {cin}->{cout}", 2.0 * w_packed.shape[0] * cin * cout * vox, "flop")
    return out


# ----------------------------------------------------------------------------------------------------------
# fused upsample + soft-argmin
# ----------------------------------------------------------------------------------------------------------
def regress_fwd(cost: torch.Tensor, mindisp: float, step: float, want_prob: bool = False, align_corners: bool = True):
    """align_corners=False: half-pixel coordinates on all three axes (NNet, dpf_regress_fwd_halfpixel)."""
    _req(cost, torch.float32, "cost")
    b, d, h4, w4 = cost.shape
    disp = torch.empty(b, 4 * h4, 4 * w4, device=cost.device, dtype=torch.float32)
    prob = torch.empty(b, 4 * d, 4 * h4, 4 * w4, device=cost.device, dtype=torch.float32) if want_prob else None
    tm = _timing_begin()
    fn = lib().dpf_regress_fwd if align_corners else lib().dpf_regress_fwd_halfpixel
    check(fn(_p(cost), _p(disp), _p(prob), b, d, h4, w4, float(mindisp), float(step), _stream()), "dpf_regress_fwd")
    _timing_end(tm, "regress_fwd", cost.numel() * 4.0 + disp.numel() * 4.0 + (prob.numel() * 4.0 if prob is not None else 0.0), "byte")
    return disp, prob


def regress_bwd(cost: torch.Tensor, ddisp: torch.Tensor, mindisp: float, step: float) -> torch.Tensor:
    _req(cost, torch.float32, "cost"); _req(ddisp, torch.float32, "ddisp")
    b, d, h4, w4 = cost.shape
    dcost = torch.empty_like(cost)
    check(lib().dpf_regress_bwd(_p(cost), _p(ddisp), _p(dcost), b, d, h4, w4, float(mindisp), float(step), _stream()), "dpf_regress_bwd")
    return dcost


# ----------------------------------------------------------------------------------------------------------
# ASM sampling / blend
# ----------------------------------------------------------------------------------------------------------
def asm_sample(x: torch.Tensor, tables: dict) -> torch.Tensor:
    """x [B,H4,W4,C] bf16 + device tables (shift_tables.build_tables) -> samples [B,S,H4,W4,C] bf16."""
    _req(x, torch.bfloat16, "x")
    b, h, w, c = x.shape
    s = tables["ri"].shape[0]
    assert tables["ri"].shape == (s, h, 2) and tables["ci"].shape == (s, w, 2)
    out = torch.empty(b, s, h, w, c, device=x.device, dtype=torch.bfloat16)
    check(lib().dpf_asm_sample_fwd(_p(x), _p(out), b, h, w, c, s, _p(tables["ri"]), _p(tables["rw"]), _p(tables["ci"]),
                                   _p(tables["cw"]), _stream()), "dpf_asm_sample_fwd")
    return out


def channel_stats(x: torch.Tensor) -> torch.Tensor:
    """x [B,...,C] bf16 -> [B,C,2] fp32 (sum, sum of squares over all positions); deterministic two-pass reduction."""
    _req(x, torch.bfloat16, "x")
    b, c = x.shape[0], x.shape[-1]
    p = x.numel() // (b * c)
    stats = torch.empty(b, c, 2, device=x.device, dtype=torch.float32)
    ws = torch.empty(int(lib().dpf_channel_stats_ws_floats(b, p, c)), device=x.device, dtype=torch.float32)   # caller-owned scratch
    check(lib().dpf_channel_stats(_p(x), _p(stats), _p(ws), b, p, c, _stream()), "dpf_channel_stats")
    return stats


def asm_blend(samples: torch.Tensor, logits: torch.Tensor, in_a: torch.Tensor, in_d: torch.Tensor, vol: torch.Tensor,
              d0: int, d_rep: int, ch_off: int) -> None:
    _req(samples, torch.bfloat16, "samples"); _req(logits, torch.bfloat16, "logits")
    _req(in_a, torch.float32, "in_a"); _req(in_d, torch.float32, "in_d"); _req(vol, torch.bfloat16, "vol")
    b, s, h, w, c = samples.shape
    assert logits.shape == samples.shape and in_a.shape == (b, c) and in_d.shape == (b, c)
    assert vol.shape[0] == b and vol.shape[2:4] == (h, w)
    tm = _timing_begin()
    check(lib().dpf_asm_blend_fwd(_p(samples), _p(logits), _p(in_a), _p(in_d), _p(vol), b, h, w, c, s, vol.shape[1], d0, d_rep,
                                  ch_off, vol.shape[-1], _stream()), "dpf_asm_blend_fwd")
    # algorithmic bytes: samples + logits read once, d_rep volume slices of c channels written
    _timing_end(tm, "asm_blend", 2.0 * (samples.numel() + logits.numel()) + 2.0 * b * h * w * c * d_rep, "byte")


# ----------------------------------------------------------------------------------------------------------
# ANM front end
# ----------------------------------------------------------------------------------------------------------
def anm_select(disp: torch.Tensor, kinv: torch.Tensor, abvalue: torch.Tensor, levels: Sequence[float], k: int):
    """disp [B,H,W] fp32 -> idx [B,K,H4,W4] int32, coord [B,K,H4,W4,3] fp32 (un-normalised), minmax [B,2] fp32."""
    _req(disp, torch.float32, "disp"); _req(kinv, torch.float32, "kinv"); _req(abvalue, torch.float32, "abvalue")
    b, h, w = disp.shape
    h4, w4 = h // 4, w // 4
    d = len(levels)
    idx = torch.empty(b, k, h4, w4, device=disp.device, dtype=torch.int32)
    coord = torch.empty(b, k, h4, w4, 3, device=disp.device, dtype=torch.float32)
    minmax = torch.empty(b, 2, device=disp.device, dtype=torch.float32)     # device-side fills: capturable in a CUDA graph
    minmax[:, 0] = float("inf")
    minmax[:, 1] = float("-inf")
    lv = (C.c_float * d)(*[float(v) for v in levels])
    check(lib().dpf_anm_select(_p(disp), _p(kinv), _p(abvalue), lv, _p(idx), _p(coord), _p(minmax), b, d, k, h4, w4, _stream()),
          "dpf_anm_select")
    return idx, coord, minmax


def dcn3d(x: torch.Tensor, offset: torch.Tensor, w_packed: torch.Tensor, cin_pad: int, scale: Optional[torch.Tensor] = None,
          shift: Optional[torch.Tensor] = None, relu: bool = False, cin_real: Optional[int] = None) -> torch.Tensor:
    """3-D deformable conv 3x3x3 (stride 1, pad 1): x [B,D,H,W,Cs] bf16, offset [B,D,H,W,>=81] fp32 -> [B,D,H,W,64] bf16.
    `cin_real` (<= cin_pad) is the layer's true input width; it only enters the algorithmic-FLOP accounting of the timing hook."""
    _req(x, torch.bfloat16, "x"); _req(offset, torch.float32, "offset"); _req(w_packed, torch.bfloat16, "w_packed")
    b, d, h, w, cs = x.shape
    assert offset.shape[:4] == (b, d, h, w) and offset.shape[-1] >= 81 and offset.is_contiguous()
    y = torch.empty(b, d, h, w, 64, device=x.device, dtype=torch.bfloat16)
    tm = _timing_begin()
    check(lib().dpf_dcn3d_fwd(_p(x), _p(offset), _p(w_packed), _p(scale), _p(shift), _p(y), b, d, h, w, cin_pad, cs,
                              offset.shape[-1], 64, int(relu), _stream()), "dpf_dcn3d_fwd")
    _timing_end(tm, "dcn3d_kernel<64>", 2.0 * 27 * (cin_real or cin_pad) * 64 * b * d * h * w, "flop")
    return y


def anm_gather(out3: torch.Tensor, idx: torch.Tensor, coord: torch.Tensor, minmax: torch.Tensor, cpad: int = 64) -> torch.Tensor:
    """out3 [B,D,H4,W4,C] bf16 -> feature volume [B,K,H4,W4,Cpad] bf16 (C cost channels, 3 coords, zero pad)."""
    _req(out3, torch.bfloat16, "out3")
    b, d, h4, w4, c = out3.shape
    k = idx.shape[1]
    fv = torch.empty(b, k, h4, w4, cpad, device=out3.device, dtype=torch.bfloat16)
    check(lib().dpf_anm_gather(_p(out3), _p(idx), _p(coord), _p(minmax), _p(fv), b, d, k, h4, w4, c, cpad, _stream()), "dpf_anm_gather")
    return fv


# ----------------------------------------------------------------------------------------------------------
# fused bias + residual + activation (channels-last bf16)
# ----------------------------------------------------------------------------------------------------------
def bias_act(x: torch.Tensor, bias: Optional[torch.Tensor], slope: float, res: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None, y_coff: int = 0) -> torch.Tensor:
    """y[..., y_coff:y_coff+C] = act(x + bias + res); x is an NCHW-shaped tensor in channels_last memory format (or any
    [...,C]-contiguous bf16 tensor).  slope: 0 = ReLU, 1 = identity, else PReLU / LeakyReLU slope.  In place when out is None."""
    if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
        c = x.shape[1]
    elif x.is_contiguous():
        c = x.shape[-1] if x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last) else x.shape[1]
    else:
        raise _lib.DpfError("bias_act: expected a dense channels-last bf16 tensor")
    if x.dtype != torch.bfloat16 or not x.is_cuda:
        raise _lib.DpfError("bias_act: expected a CUDA bf16 tensor")
    npix = x.numel() // c
    if out is None:
        out, cstride = x, c
    else:
        cstride = out.numel() // npix
    if bias is not None:
        _req(bias, torch.float32, "bias")
    check(lib().dpf_bias_act(_p(x), _p(bias), _p(res), _p(out), npix, c, cstride, y_coff, float(slope), _stream()), "dpf_bias_act")
    return out


def channel_max(x: torch.Tensor) -> torch.Tensor:
    """x [..., C] bf16 contiguous -> [...] fp32, maximum over the channels (ref_feature of the model output)."""
    _req(x, torch.bfloat16, "x")
    c = x.shape[-1]
    y = torch.empty(x.shape[:-1], device=x.device, dtype=torch.float32)
    check(lib().dpf_channel_max(_p(x), _p(y), x.numel() // c, c, _stream()), "dpf_channel_max")
    return y


def pack_conv2d_weight(w: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d weight [Cout<=32, 32|64, 3,3] -> the packed 3x3x3 layout of the kd-fused kernel with the image's kh on the
    depth taps and only the centre in-plane row populated (dpf_conv2d_fwd)."""
    cout, cin, kh, kw = w.shape
    assert (kh, kw) == (3, 3)
    w3 = torch.zeros(cout, cin, 3, 3, 3, device=w.device, dtype=torch.float32)
    w3[:, :, :, 1, :] = w.detach().float()
    return pack_conv_weight(w3)


def conv2d_rows(x: torch.Tensor, w_packed: torch.Tensor, cout: int, scale: Optional[torch.Tensor] = None,
                shift: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, relu: bool = False,
                slope: float = 0.0, out: Optional[torch.Tensor] = None, y_coff: int = 0, x_coff: int = 0, dil: int = 1) -> torch.Tensor:
    """3x3 stride-1 conv (dilation dil, padding dil) on channels-last images: x [N,H,W,Cx] bf16 (32 input channels from x_coff) -> y [N,H,W,Cy]
    (cout channels at y_coff), y = act(conv * scale + shift + residual), act = LeakyReLU(slope) when relu (slope 0 = ReLU)."""
    _req(x, torch.bfloat16, "x"); _req(w_packed, torch.bfloat16, "w_packed")
    n, h, w, cx = x.shape
    if out is None:
        out = torch.empty(n, h, w, cout, device=x.device, dtype=torch.bfloat16)
    _req(out, torch.bfloat16, "out")
    assert out.shape[:3] == (n, h, w)
    if residual is not None:
        _req(residual, torch.bfloat16, "residual")
        assert residual.shape == out.shape
    for t, nm in ((scale, "scale"), (shift, "shift")):
        if t is not None:
            _req(t, torch.float32, nm)
            assert t.numel() == cout
    cin = w_packed.shape[1] * 8
    check(lib().dpf_conv2d_fwd(_p(x), _p(w_packed), _p(out), _p(scale), _p(shift), _p(residual), n, h, w, cin, cout, cx, x_coff,
                               out.shape[-1], y_coff, int(dil), int(relu), float(slope), _stream()), "dpf_conv2d_fwd")
    return out


def conv2d_rows_plan(weight: torch.Tensor):
    """Per-launch packed weights of a 3x3 conv with 32 | 64 input channels and any multiple-of-8 output width: output-channel
    chunks of <= 32 -> [(packed, y_coff, n)]."""
    cout = weight.shape[0]
    return [(pack_conv2d_weight(weight[co:co + 32]), co, min(32, cout - co)) for co in range(0, cout, 32)]


def conv2d_rows_multi(x, plan, bias=None, residual=None, relu=False, slope=0.0, out=None, y_coff=0, dil=1):
    """conv2d_rows over the output-channel chunks of `plan`; out [N,H,W,Cy] receives them at y_coff."""
    cout = sum(n for _, _, n in plan)
    if out is None:
        out = torch.empty(*x.shape[:3], cout, device=x.device, dtype=torch.bfloat16)
    for wp, co, n in plan:
        sh = bias[co:co + n].contiguous() if bias is not None else None
        if residual is not None and (residual.shape[-1] != out.shape[-1] or y_coff != 0):
            raise _lib.DpfError("conv2d_rows_multi: the residual must have the layout of the output tensor")
        conv2d_rows(x, wp, n, None, sh, residual, relu, slope, out, y_coff + co, dil=dil)
    return out


# ----------------------------------------------------------------------------------------------------------
# dedicated 2-D tcgen05 convolution (any dilation, Cin 32 | 64 | 96, Cout <= 96 in one launch): conv2d_tc.cu
# ----------------------------------------------------------------------------------------------------------
def pack_conv2d_tc_weight(w: torch.Tensor, cin_pad: Optional[int] = None) -> torch.Tensor:
    """nn.Conv2d weight [Cout,Cin,3,3] -> bf16 [9 taps (kh,kw)][Cin_pad/8][Npad][8], Npad = ceil16(Cout) (zero padded)."""
    cout, cin, kh, kw = w.shape
    assert (kh, kw) == (3, 3)
    cin_pad = cin_pad or cin
    assert cin_pad % 8 == 0 and cin_pad >= cin
    npad = (cout + 15) // 16 * 16
    buf = torch.zeros(9, cin_pad, npad, device=w.device, dtype=torch.float32)
    buf[:, :cin, :cout] = w.detach().float().permute(2, 3, 1, 0).reshape(9, cin, cout)
    return buf.reshape(9, cin_pad // 8, 8, npad).permute(0, 1, 3, 2).contiguous().to(torch.bfloat16)


def conv2d_tc(x: torch.Tensor, w_packed: torch.Tensor, cout: int, dil: int = 1, scale: Optional[torch.Tensor] = None,
              shift: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, relu: bool = False, slope: float = 0.0,
              out: Optional[torch.Tensor] = None, y_coff: int = 0, x_coff: int = 0, res_post: bool = False) -> torch.Tensor:
    """3x3 / stride 1 / dilation dil / padding dil conv on channels-last images in ONE launch (dpf_conv2d_tc_fwd):
    x [N,H,W,Cx] bf16 (Cin = 8 * w_packed.shape[1] channels from x_coff) -> y [N,H,W,Cy] (ceil8(cout) channels at y_coff; the ones
    beyond cout are exact zeros), y = act(conv * scale + shift + residual), act = LeakyReLU(slope) when relu (slope 0 = ReLU);
    res_post: y = act(conv * scale + shift) + residual."""
    _req(x, torch.bfloat16, "x"); _req(w_packed, torch.bfloat16, "w_packed")
    n, h, w, cx = x.shape
    cin = w_packed.shape[1] * 8
    cst = (cout + 7) // 8 * 8
    if out is None:
        out = torch.empty(n, h, w, cst, device=x.device, dtype=torch.bfloat16)
    _req(out, torch.bfloat16, "out")
    assert out.shape[:3] == (n, h, w) and w_packed.shape[0] == 9 and w_packed.shape[2] == (cout + 15) // 16 * 16
    if residual is not None:
        _req(residual, torch.bfloat16, "residual")
        assert residual.shape == out.shape
    for t, nm in ((scale, "scale"), (shift, "shift")):
        if t is not None:
            _req(t, torch.float32, nm)
            assert t.numel() == cout
    tm = _timing_begin()
    check(lib().dpf_conv2d_tc_fwd(_p(x), _p(w_packed), _p(out), _p(scale), _p(shift), _p(residual), n, h, w, cin, cout, cx, x_coff,
                                  out.shape[-1], y_coff, int(dil), int(relu) | (2 if res_post else 0), float(slope), _stream()), "dpf_conv2d_tc_fwd")
    _timing_end(tm, f"conv2d_tc {cin}->{cout} d{dil}", 2.0 * 9 * cin * cout * n * h * w, "flop")
    return out


def conv3d_s2(x: torch.Tensor, w_packed: torch.Tensor, cout: int, scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
              relu: bool = False, out: Optional[torch.Tensor] = None, y_coff: int = 0) -> torch.Tensor:
    """Stride-2 3x3x3 conv (+ affine + ReLU) in ONE launch of the plane-streamed kernel (dpf_conv3d_s2_fwd): x [B,D,H,W,Cin] bf16
    with Cin = 8 * w_packed.shape[1] in {32, 64} -> [B,ceil(D/2),ceil(H/2),ceil(W/2),Cy] (cout channels at y_coff)."""
    _req(x, torch.bfloat16, "x"); _req(w_packed, torch.bfloat16, "w_packed")
    b, d, h, w, cx = x.shape
    cin = w_packed.shape[1] * 8
    do, ho, wo = (d + 1) // 2, (h + 1) // 2, (w + 1) // 2
    if out is None:
        out = torch.empty(b, do, ho, wo, cout, device=x.device, dtype=torch.bfloat16)
    _req(out, torch.bfloat16, "out")
    assert out.shape[:4] == (b, do, ho, wo) and w_packed.shape[0] == 27
    tm = _timing_begin()
    check(lib().dpf_conv3d_s2_fwd(_p(x), _p(w_packed), _p(out), _p(scale), _p(shift), b, d, h, w, cin, cout, cx, 0, out.shape[-1], y_coff,
                                  int(relu), _stream()), "dpf_conv3d_s2_fwd")
    _timing_end(tm, f"conv3d_s2 {cin}->{cout}", 2.0 * 27 * cin * cout * b * do * ho * wo, "flop")
    return out


def softargmin(cost: torch.Tensor, mindisp: float, step: float, want_prob: bool = False):
    """Soft-argmin over dim 1 without up-sampling (dpf_softargmin_fwd): cost [B,D,*] fp32 -> (disp [B,*], prob [B,D,*] | None)."""
    _req(cost, torch.float32, "cost")
    b, d = cost.shape[:2]
    p = cost[0, 0].numel()
    disp = torch.empty(b, *cost.shape[2:], device=cost.device, dtype=torch.float32)
    prob = torch.empty_like(cost) if want_prob else None
    check(lib().dpf_softargmin_fwd(_p(cost), _p(disp), _p(prob), b, d, p, float(mindisp), float(step), _stream()), "dpf_softargmin_fwd")
    return disp, prob


def pack_head_weight(w: torch.Tensor) -> torch.Tensor:
    """nn.Conv3d(32, 1, 3) weight [1,32,3,3,3] -> bf16 [4 (c/8)][32 (tap; 27 real + 5 zero)][8 (c%8)] for dpf_conv3d_head_fwd."""
    assert tuple(w.shape) == (1, 32, 3, 3, 3), w.shape
    buf = torch.zeros(4, 32, 8, device=w.device, dtype=torch.float32)
    buf[:, :27] = w.detach().float().reshape(4, 8, 27).permute(0, 2, 1)
    return buf.to(torch.bfloat16).contiguous()


def conv3d_head(x: torch.Tensor, w_packed: torch.Tensor, residual: Optional[torch.Tensor] = None, shift: float = 0.0,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """32 -> 1 channel 3x3x3 conv on the bandwidth-bound head kernel: x [B,D,H,W,Cx>=32] bf16 -> [B,D,H,W,1] fp32 (+ shift + residual)."""
    _req(x, torch.bfloat16, "x"); _req(w_packed, torch.bfloat16, "w_packed")
    b, d, h, w, cx = x.shape
    if out is None:
        out = torch.empty(b, d, h, w, 1, device=x.device, dtype=torch.float32)
    _req(out, torch.float32, "out")
    assert out.numel() == b * d * h * w
    if residual is not None:
        _req(residual, torch.float32, "residual")
        assert residual.numel() == out.numel()
    tm = _timing_begin()
    check(lib().dpf_conv3d_head_fwd(_p(x), _p(w_packed), _p(out), _p(residual), float(shift), b, d, h, w, cx, _stream()), "dpf_conv3d_head_fwd")
    _timing_end(tm, "conv3d_head 32->1", float(b * d * h * w * (32 * 2 + 4 + (4 if residual is not None else 0))), "byte")
    return out


def pack_stem_weight(w: torch.Tensor) -> torch.Tensor:
    """[32, Cin <= 8, 3, 3] (BatchNorm scale already folded in) -> bf16 [80][32], k = (kh*3 + kw)*8 + ci, zero rows elsewhere."""
    cout, cin, kh, kw = w.shape
    assert cout == 32 and cin <= 8 and (kh, kw) == (3, 3), w.shape
    buf = torch.zeros(10, 8, 32, device=w.device, dtype=torch.float32)
    buf[:9, :cin] = w.detach().float().permute(2, 3, 1, 0).reshape(9, cin, 32)
    return buf.reshape(80, 32).to(torch.bfloat16).contiguous()


def stem_conv(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = True) -> torch.Tensor:
    """Encoder stem (dpf_stem_conv_fwd): x [N,H,W,8] bf16 channels-last -> [N,ceil(H/2),ceil(W/2),32] bf16, 3x3 stride 2 pad 1 + bias + ReLU."""
    _req(x, torch.bfloat16, "x"); _req(w_packed, torch.bfloat16, "w_packed")
    n, h, w, c = x.shape
    assert c == 8 and tuple(w_packed.shape) == (80, 32)
    if bias is not None:
        _req(bias, torch.float32, "bias")
    y = torch.empty(n, (h + 1) // 2, (w + 1) // 2, 32, device=x.device, dtype=torch.bfloat16)
    tm = _timing_begin()
    check(lib().dpf_stem_conv_fwd(_p(x), _p(w_packed), _p(bias), _p(y), n, h, w, int(relu), _stream()), "dpf_stem_conv_fwd")
    _timing_end(tm, "stem_conv 3->32 s2", float(x.numel() * 2 + y.numel() * 2), "byte")
    return y
