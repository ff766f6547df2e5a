#!/usr/bin/env python
"""CPU dry run of dualpixelface_b200.nnet's EVAL orchestration (weight re-ordering, channel layouts, depth-pair folding, batching of the
context refinement) against the oracle, with every kernel wrapper replaced by a plain PyTorch fp32 emulation of its contract.
Development tool for a container without a GPU: it checks the host-side wiring only -- the kernels themselves are checked on a
B200 by tests/test_gpu_nnet.py."""
import json
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from dualpixelface_b200 import models, nnet, ops  # noqa: E402
from dualpixelface_b200.runner import load_config, model_selector  # noqa: E402
from dualpixelface_b200.synthetic import synth_state, synthetic_batch  # noqa: E402
from oracle import dpf_oracle as O  # noqa: E402


class EmuConv3d:
    def __init__(self, weight, kind, transposed=False, cin_pad=None):
        self.w = weight.detach().float()
        self.cin = cin_pad or self.w.shape[1]
        self.cout = self.w.shape[0]

    def __call__(self, x, scale=None, shift=None, residual=None, relu=False, out_f32=False, slope=0.0, **kw):
        assert x.shape[-1] == self.cin
        xi = x.float().permute(0, 4, 1, 2, 3)[:, : self.w.shape[1]]
        y = F.conv3d(xi, self.w, None, 1, 1).permute(0, 2, 3, 4, 1)
        if scale is not None:
            y = y * scale
        if shift is not None:
            y = y + shift
        if residual is not None:
            y = y + residual.float()
        if relu:
            y = F.leaky_relu(y, slope)
        return y.contiguous()


def emu_pack2d(w, cin_pad=None):
    return w.detach().float()


def emu_conv2d_tc(x, w, cout, dil=1, scale=None, shift=None, residual=None, relu=False, slope=0.0, res_post=False, **kw):
    assert x.shape[-1] == w.shape[1], (x.shape, w.shape)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, None, 1, dil, dil).permute(0, 2, 3, 1)
    if scale is not None:
        y = y * scale
    if shift is not None:
        y = y + shift
    if relu:
        y = F.leaky_relu(y, slope)
    cst = (cout + 7) // 8 * 8
    return F.pad(y, (0, cst - cout)).contiguous()


def emu_costvol(ref, tgt, shifts, mode="concat", groups=0):
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
    assert bias is None and res is None
    return F.leaky_relu(x, slope)


if __name__ == "__main__":
    torch.set_num_threads(8)
    nnet.TCConv3d = EmuConv3d
    ops.pack_conv2d_tc_weight = emu_pack2d
    ops.conv2d_tc = emu_conv2d_tc
    ops.costvol_fwd = emu_costvol
    ops.regress_fwd = emu_regress
    ops.bias_act = emu_bias_act
    ops.channel_max = lambda x: x.amax(-1).float()
    shapes = {k: tuple(v) for k, v in json.loads((ROOT / "tests/golden/state_keys_nnet.json").read_text()).items()}
    batch = synthetic_batch(2, 256, 256, training=True, seed=0)
    st = synth_state(shapes, seed=1)
    stats = {}
    with torch.no_grad():
        O.nnet_forward(dict(batch), st, True, stats=stats)
        st = O.calibrate_running_stats(st, stats)
        want = O.nnet_forward(dict(batch), st, False)
    model = model_selector(load_config("eval_faceDP_nnet", "test", root=ROOT, make_dirs=False), root=ROOT)
    sd = model.state_dict()
    assert set(sd) == set(shapes), (set(sd) ^ set(shapes))
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    missing = model.load_state_dict(st, strict=False)
    assert list(missing.missing_keys) == ["normal_module.costrange"] and not missing.unexpected_keys      # a derived constant
    model.eval()
    model.want_prob = True

    # CPU stand-ins for the two CUDA-only entry points of the base class
    def features(ref_img, tgt_img):
        f = model.feature_extraction(torch.cat([ref_img, tgt_img], 0).float()).permute(0, 2, 3, 1).contiguous()
        b = ref_img.shape[0]
        return f[:b], f[b:]

    model._features = features
    torch.Tensor.is_cuda = property(lambda self: True)
    # fp32 emulation: keep the bf16 staging buffers of the orchestration in fp32 so that the comparison isolates the wiring
    _zeros = torch.zeros
    torch.zeros = lambda *a, **k: _zeros(*a, **{**k, "dtype": torch.float32 if k.get("dtype") == torch.bfloat16 else k.get("dtype")})
    plan = model._build()
    plan["ctx_cudnn"] = [(w.float(), d) for w, d in plan["ctx_cudnn"]]
    with torch.no_grad():
        got = model(dict(batch))
    for k in ("pred_depth", "pred_normal", "ref_feature", "prob_depth"):
        err = (got[k].float() - want[k]).abs().max().item()
        print(f"{k}: max err {err:.3e} (range {want[k].abs().max().item():.3f})")
        assert err < 2e-2, k                    # the four cuDNN-layer weights of the plan are bf16-rounded (refined head only)
    e0 = (got["pred_depth"][:, 0] - want["pred_depth"][:, 0]).abs().max().item()
    print(f"raw head (no bf16 anywhere in this emulation): max err {e0:.3e}")
    assert e0 < 2e-4
    print("OK")
