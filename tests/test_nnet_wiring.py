"""CPU: the EVAL orchestration of dualpixelface_b200.nnet (weight re-ordering, channel layouts, depth-pair folding, batching of the
context refinement over the levels, half-pixel regression call) against the oracle, with every kernel wrapper replaced by a plain
PyTorch fp32 emulation of its documented contract.  Checks the host-side wiring only -- the kernels themselves are checked on a
B200 by tests/test_gpu_nnet.py."""
import json

import torch
import torch.nn.functional as F

from conftest import GOLDEN, ROOT
from dualpixelface_b200 import nnet, ops
from dualpixelface_b200.runner import load_config, model_selector
from dualpixelface_b200.synthetic import synth_state, synthetic_batch
from oracle import dpf_oracle as O


class EmuConv3d:
    """layers.TCConv3d's call contract on [B,D,H,W,C] tensors."""

    def __init__(self, weight, kind, transposed=False, cin_pad=None):
        self.w = weight.detach().float()
        self.cin = cin_pad or self.w.shape[1]
        self.cout = self.w.shape[0]

    def __call__(self, x, scale=None, shift=None, residual=None, relu=False, out_f32=False, slope=0.0, **kw):
        assert x.shape[-1] == self.cin
        assert float(x[..., self.w.shape[1]:].abs().max() if x.shape[-1] > self.w.shape[1] else 0.0) == 0.0      # pad channels are zero
        y = F.conv3d(x.float().permute(0, 4, 1, 2, 3)[:, : self.w.shape[1]], self.w, None, 1, 1).permute(0, 2, 3, 4, 1)
        if scale is not None:
            y = y * scale
        if shift is not None:
            y = y + shift
        if residual is not None:
            y = y + residual.float()
        if relu:
            y = F.leaky_relu(y, slope)
        return y.contiguous()


def emu_conv2d_tc(x, w, cout, dil=1, scale=None, shift=None, residual=None, relu=False, slope=0.0, res_post=False, **kw):
    assert x.shape[-1] == w.shape[1] and residual is None, (x.shape, w.shape)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, None, 1, dil, dil).permute(0, 2, 3, 1)
    if scale is not None:
        y = y * scale
    if shift is not None:
        y = y + shift
    if relu:
        y = F.leaky_relu(y, slope)
    return F.pad(y, (0, (cout + 7) // 8 * 8 - cout)).contiguous()              # ceil8(cout) channels, the extra ones exact zeros


def emu_costvol(ref, tgt, shifts, mode="concat", groups=0):
    assert mode == "concat"
    b, h, w, c = ref.shape
    vol = ref.new_zeros(b, len(shifts), h, w, 2 * c)
    for i, d in enumerate(shifts):
        dst, rr, tr = O._row_windows(h, d)
        vol[:, i, dst, :, :c] = ref[:, rr]
        vol[:, i, dst, :, c:] = tgt[:, tr]
    return vol


def emu_regress(cost, mindisp, step, want_prob=False, align_corners=True):
    up = F.interpolate(cost.unsqueeze(1), scale_factor=4, mode="trilinear", align_corners=align_corners).squeeze(1)
    prob = F.softmax(up, 1)
    bins = torch.arange(up.shape[1], dtype=torch.float32) * step + mindisp
    return (prob * bins.view(1, -1, 1, 1)).sum(1), (prob if want_prob else None)


def emu_bias_act(x, bias, slope, res=None, out=None, y_coff=0):
    assert bias is None and res is None and out is None
    return F.leaky_relu(x, slope)


def test_nnet_eval_orchestration_matches_oracle(monkeypatch):
    torch.set_num_threads(8)
    monkeypatch.setattr(nnet, "TCConv3d", EmuConv3d)
    monkeypatch.setattr(ops, "pack_conv2d_tc_weight", lambda w, cin_pad=None: w.detach().float())
    monkeypatch.setattr(ops, "conv2d_tc", emu_conv2d_tc)
    monkeypatch.setattr(ops, "costvol_fwd", emu_costvol)
    monkeypatch.setattr(ops, "regress_fwd", emu_regress)
    monkeypatch.setattr(ops, "bias_act", emu_bias_act)
    monkeypatch.setattr(ops, "channel_max", lambda x: x.amax(-1).float())
    shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_nnet.json").read_text()).items()}
    batch = synthetic_batch(2, 256, 256, training=True, seed=0)
    st = synth_state(shapes, seed=1)
    stats = {}
    with torch.no_grad():
        O.nnet_forward(dict(batch), st, True, stats=stats)
        st = O.calibrate_running_stats(st, stats)
        want = O.nnet_forward(dict(batch), st, False)
    model = model_selector(load_config("eval_faceDP_nnet", "test", root=ROOT, make_dirs=False), root=ROOT)
    missing = model.load_state_dict(st, strict=False)
    assert list(missing.missing_keys) == ["normal_module.costrange"] and not missing.unexpected_keys      # a derived constant
    model.eval()
    model.want_prob = True

    def features(ref_img, tgt_img):                      # CPU stand-in for the CUDA-only encoder entry point of the base class
        f = model.feature_extraction(torch.cat([ref_img, tgt_img], 0).float()).permute(0, 2, 3, 1).contiguous()
        return f[: ref_img.shape[0]], f[ref_img.shape[0]:]

    monkeypatch.setattr(model, "_features", features)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    # fp32 emulation: the bf16 staging buffers of the orchestration become fp32, so that the comparison isolates the wiring
    zeros = torch.zeros
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: zeros(*a, **{**k, "dtype": torch.float32 if k.get("dtype") == torch.bfloat16 else k.get("dtype")}))
    plan = model._build()
    plan["ctx_cudnn"] = [(w.float(), d) for w, d in plan["ctx_cudnn"]]
    with torch.no_grad():
        got = model(dict(batch))
    err = {k: (got[k].float() - want[k]).abs().max().item() for k in ("pred_depth", "pred_normal", "ref_feature", "prob_depth")}
    raw = (got["pred_depth"][:, 0] - want["pred_depth"][:, 0]).abs().max().item()
    print(err, raw)
    assert raw < 2e-4 and err["pred_normal"] < 1e-3 and err["ref_feature"] < 1e-5 and err["prob_depth"] < 1e-3
    assert err["pred_depth"] < 2e-2          # the four cuDNN-layer weights of the plan are bf16-rounded (refined head only)
