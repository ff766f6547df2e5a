"""CPU: host-side logic of the product -- sampling tables, launch planning, weight packing, BN folding, config / model
contract -- checked against the reference-derived golden fixtures where one exists."""
import json

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, ROOT
from dualpixelface_b200 import layers
from dualpixelface_b200.shift_tables import build_tables
from dualpixelface_b200.synthetic import synth_state, synthetic_batch


def apply_tables(x, tb):
    """Pure-torch evaluation of what dpf_asm_sample_fwd computes (test helper)."""
    b, c, h, w = x.shape
    out = []
    for s in range(tb["ri"].shape[0]):
        acc = torch.zeros_like(x)
        for i in range(2):
            for j in range(2):
                ri, ci = tb["ri"][s, :, i].long(), tb["ci"][s, :, j].long()
                wgt = tb["rw"][s, :, i].view(h, 1) * tb["cw"][s, :, j].view(1, w)
                ok = (ri >= 0).view(h, 1) & (ci >= 0).view(1, w)
                g = x[:, :, ri.clamp_min(0)][:, :, :, ci.clamp_min(0)]
                acc = acc + g * (wgt * ok)
        out.append(acc)
    return torch.stack(out, -1)


@pytest.mark.parametrize("direction", ["forward", "backward"])
@pytest.mark.parametrize("disp", [-1.0, 1.0])
def test_shift_tables_reproduce_reference_samples(golden_stages, disp, direction):
    g = torch.Generator().manual_seed(11)
    x = torch.relu(torch.randn(2, 4, 16, 24, generator=g))
    got = apply_tables(x, build_tables(16, 24, disp, direction))
    want = torch.as_tensor(golden_stages[f"shift/{disp}/{direction}"])
    assert torch.allclose(got[..., 0], want[..., 0], atol=0)                 # nearest: exact gather
    assert torch.allclose(got[..., 1], want[..., 1], atol=1e-5)              # bilinear: same weights, other sum order
    assert torch.allclose(got[..., 2], want[..., 2], atol=1e-5)              # integer phase shift == circular roll


@pytest.mark.parametrize("hw", [(112, 112), (128, 192), (280, 420), (560, 840)])
@pytest.mark.parametrize("direction", ["forward", "backward"])
def test_nearest_tables_bitexact_at_baseline_resolutions(golden_stages, hw, direction):
    tb = build_tables(hw[0], hw[1], -1.0, direction, (True, False, False))
    assert np.array_equal(tb["ri"][0, :, 0].numpy(), golden_stages[f"nearest_rows/{hw[0]}x{hw[1]}/{direction}"])
    assert np.array_equal(tb["ci"][0, :, 0].numpy(), golden_stages[f"nearest_cols/{hw[0]}x{hw[1]}/{direction}"])
    assert tb["ci"][0, -1, 0].item() == -1                                   # last column rounds out of bounds


@pytest.mark.parametrize("disp", [-0.5, 0.5, 2.5])
@pytest.mark.parametrize("direction", ["forward", "backward"])
def test_fractional_phase_shift_matches_reference(golden_stages, disp, direction):
    """cached_first_level=False: a fractional Fourier (phase) shift is not a table sample; build_tables hands out the row-frequency
    rotation and shift_tables.fourier_row_shift evaluates asm.py:112-125 (legacy C2R irfft included) == the reference's samples."""
    from dualpixelface_b200.shift_tables import fourier_row_shift
    g = torch.Generator().manual_seed(11)
    x = torch.relu(torch.randn(2, 4, 16, 24, generator=g))
    tb = build_tables(16, 24, disp, direction)
    assert tb["ri"].shape[0] == 2 and tb["rot"].shape == (16,)              # nearest + bilinear tables, phase by FFT
    got = torch.cat([apply_tables(x, tb), fourier_row_shift(x.permute(0, 2, 3, 1).contiguous(), tb["rot"]).permute(0, 3, 1, 2).unsqueeze(-1)], -1)
    want = torch.as_tensor(golden_stages[f"shift/{disp}/{direction}"])
    assert torch.allclose(got[..., 0], want[..., 0], atol=0) and torch.allclose(got[..., 1:], want[..., 1:], atol=1e-5)


def test_plan_launches():
    P = layers.plan_launches
    assert [(l.cin, l.cout) for l in P(layers.KIND_3x3x3, 32, 32)] == [(32, 32)]
    assert [(l.y_coff, l.cout) for l in P(layers.KIND_3x3x3, 64, 64)] == [(0, 32), (32, 32)]
    assert [(l.y_coff, l.cout) for l in P(layers.KIND_3x3x3, 64, 81)] == [(0, 32), (32, 32), (64, 17)]
    s2 = P(layers.KIND_S2, 64, 64)
    assert [(l.x_coff, l.y_coff, l.first_k, l.last_k) for l in s2] == [(0, 0, True, False), (32, 0, False, True),
                                                                      (0, 32, True, False), (32, 32, False, True)]
    assert len(P(layers.KIND_T2, 64, 64)) == 2
    k96 = P(layers.KIND_3x3x3, 96, 64)                       # offset-conv data gradient: 81 -> 96 padded channels
    assert [(l.x_coff, l.cin, l.y_coff, l.first_k, l.last_k) for l in k96] == [
        (0, 32, 0, True, False), (32, 32, 0, False, False), (64, 32, 0, False, True),
        (0, 32, 32, True, False), (32, 32, 32, False, False), (64, 32, 32, False, True)]
    assert [(l.y_coff, l.cout) for l in P(layers.KIND_1x3x3, 32, 64)] == [(0, 64)]
    with pytest.raises(ValueError):
        P(layers.KIND_3x3x3, 48, 32)


def test_fold_bn_matches_batchnorm_eval():
    g = torch.Generator().manual_seed(3)
    bn = torch.nn.BatchNorm3d(8).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(8, generator=g) + 0.5); bn.bias.copy_(torch.randn(8, generator=g))
        bn.running_mean.copy_(torch.randn(8, generator=g)); bn.running_var.copy_(torch.rand(8, generator=g) + 0.5)
    x = torch.randn(2, 8, 3, 4, 5, generator=g)
    sc, sh = layers.fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
    assert torch.allclose(bn(x), x * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1), atol=1e-5)


def test_pack_conv_weight_layout():
    from dualpixelface_b200.ops import pack_conv_weight
    w = torch.arange(16 * 32 * 27, dtype=torch.float32).reshape(16, 32, 3, 3, 3) / 1000.0
    p = pack_conv_weight(w)                                                   # [27][4][16][8]
    assert p.shape == (27, 4, 16, 8) and p.dtype == torch.bfloat16
    t, c8, n, j = 14, 2, 5, 3
    assert p[t, c8, n, j] == w[n, c8 * 8 + j, t // 9, (t // 3) % 3, t % 3].to(torch.bfloat16)
    pt = pack_conv_weight(w.transpose(0, 1).contiguous(), transposed=True)    # ConvTranspose3d layout [Cin,Cout,...]
    assert torch.equal(pt, p)
    assert pack_conv_weight(w[:, :19], cin_pad=32)[:, 2, :, 3:].abs().sum() == 0   # zero padded input channels


@pytest.mark.parametrize("name,cfg", [("stereodpnet", "eval_faceDP"), ("psmnet", "eval_faceDP_psmnet")])
def test_model_contract_state_dict_keys(name, cfg):
    """The drop-in classes expose exactly the reference's state_dict keys / shapes (fixture from the real reference)."""
    from dualpixelface_b200.runner import load_config, model_selector
    model = model_selector(load_config(cfg, "pytest", root=ROOT, make_dirs=False), root=ROOT)
    ref = json.loads((GOLDEN / f"state_keys_{name}.json").read_text())
    ref.pop("normal_estimator.grid", None)            # registered lazily by the reference at its first forward
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine == ref
    model.load_state_dict(synth_state({k: tuple(v) for k, v in ref.items()}, seed=1), strict=False)
    assert type(model).__name__ == name.upper()
    for hook in ("forward", "train_dataloader", "test_dataloader", "training_step", "test_step", "configure_optimizers"):
        assert callable(getattr(model, hook))
    with pytest.raises(RuntimeError):                  # no CPU path, loudly
        model.eval()(synthetic_batch(1, 64, 96))


def test_psm_costvolume_shifts_and_config():
    from dualpixelface_b200.runner import load_config
    from dualpixelface_b200.modules import CostVolumePSM
    opt = load_config("eval_faceDP_psmnet", "pytest", root=ROOT, make_dirs=False)
    cv = CostVolumePSM(opt, opt.model.mindisp, opt.model.maxdisp)
    assert cv.shifts == [-1, 0, 0, 0, 1, 1, 2, 2]                          # int() truncation, psmnet/modules.py:229
    assert opt.model.level == 8 and opt.dataset.flip_lr is True


def test_synthetic_is_deterministic_and_aliased():
    a, b = synthetic_batch(1, 8, 8, True, seed=3), synthetic_batch(1, 8, 8, True, seed=3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    sh = {"cost_volume.attention_layer.normalize.weight": (32,), "cost_volume.attention_layer.mask_convs.3.1.weight": (32,)}
    st = synth_state(sh)
    assert torch.equal(*st.values())                                          # one InstanceNorm registered under two names


def test_losses_match_oracle_on_partial_mask():
    """Dense masked-mean form of the losses (no boolean-mask gathers) == the reference's `x[mask]` form (oracle restatement)."""
    from dualpixelface_b200 import losses
    from oracle import dpf_oracle as O
    g = torch.Generator().manual_seed(0)
    pred, disp = torch.randn(2, 3, 16, 24, generator=g), torch.randn(2, 16, 24, generator=g)
    mask = (torch.rand(2, 16, 24, generator=g) > 0.3).float()
    pn, nrm = torch.randn(2, 1, 3, 16, 24, generator=g), torch.randn(2, 3, 16, 24, generator=g)
    b = {"disp": disp, "mask": mask, "normal": nrm}
    assert abs(float(losses.smooth_l1(pred, b, (1.0, 0.7, 0.5))) - float(O.smooth_l1_multi(pred, disp, mask, (1.0, 0.7, 0.5)))) < 1e-5
    assert abs(float(losses.cosine(pn, b)) - float(O.cosine_normal_loss(pn, nrm, mask))) < 1e-5


def test_checkpoint_key_conversion():
    """Released checkpoints: lazily-registered ANM grid dropped, torchvision-0.6 FPN key names mapped to the current ones."""
    from dualpixelface_b200.models import _StereoBase
    old = {"feature_extraction.fpn.inner_blocks.1.weight": 1, "feature_extraction.fpn.layer_blocks.2.bias": 2,
           "feature_extraction.fpn.inner_blocks.0.0.weight": 3, "normal_estimator.grid": 4, "aggregation.dres0.0.0.weight": 5}
    new = _StereoBase.convert_checkpoint_keys(old)
    assert new == {"feature_extraction.fpn.inner_blocks.1.0.weight": 1, "feature_extraction.fpn.layer_blocks.2.0.bias": 2,
                   "feature_extraction.fpn.inner_blocks.0.0.weight": 3, "aggregation.dres0.0.0.weight": 5}


def test_input_size_contract():
    """H, W multiples of 16 (both models); PSMNet additionally needs a 64x64 quarter-resolution map for its pooling branch."""
    from dualpixelface_b200.models import PSMNET, STEREODPNET
    STEREODPNET.check_input_size(STEREODPNET, 1120, 1680)
    PSMNET.check_input_size(PSMNET, 448, 448)
    with pytest.raises(ValueError):
        STEREODPNET.check_input_size(STEREODPNET, 1120, 1684)
    with pytest.raises(ValueError):
        PSMNET.check_input_size(PSMNET, 128, 256)


def test_conv2d_weight_packing_puts_kh_on_the_depth_taps():
    """dpf_conv2d_fwd reuses the 3x3x3 weight layout: the image kernel's kh becomes the (fused) depth tap, kw stays, and only the
    centre in-plane row is populated."""
    from dualpixelface_b200 import ops
    w = torch.arange(32 * 32 * 9, dtype=torch.float32).reshape(32, 32, 3, 3) / 1000.0
    packed = ops.pack_conv2d_weight(w)                       # [27 taps][Cin/8][Npad][8]
    assert packed.shape == (27, 4, 32, 8)
    full = packed.float().permute(0, 1, 3, 2).reshape(3, 3, 3, 32, 32)          # [kd][kh'][kw][ci][co]
    assert float(full[:, 0].abs().max()) == 0.0 and float(full[:, 2].abs().max()) == 0.0
    want = w.to(torch.bfloat16).float().permute(2, 3, 1, 0)                      # [kh][kw][ci][co]
    assert torch.equal(full[:, 1], want)
    plan = ops.conv2d_rows_plan(torch.zeros(96, 64, 3, 3))
    assert [(co, n, tuple(p.shape)) for p, co, n in plan] == [(0, 32, (27, 8, 32, 8)), (32, 32, (27, 8, 32, 8)), (64, 32, (27, 8, 32, 8))]


def test_fused_encoder_routes_the_32_channel_convs():
    """Eval plan of the StereoDPNet encoder: the 3x3 / stride-1 32 -> 32 convs with dilation 1, 3 or 5 (firstconv 2-3, conv1 / conv2 /
    the three dilated branches of the 32-channel DPBlocks) carry a dpf_conv2d_fwd plan, everything else stays on cuDNN."""
    from dualpixelface_b200.encoder_fused import FusedSDPEncoder
    from dualpixelface_b200.runner import load_config, model_selector
    model = model_selector(load_config("eval_faceDP", "pytest", make_dirs=False))
    enc = FusedSDPEncoder(model.feature_extraction)
    routed = sum("wp" in f for f in enc.first)
    blocks = [enc.block1, *enc.inter1, enc.block2, *enc.inter2, enc.block3]
    for b in blocks:
        routed += sum("wp" in f for f in (b.c1, b.c2, *b.dil, b.c3, b.c4, b.pw, b.skip))
    assert routed == 2 + 3 * 5            # firstconv[2], firstconv[4]; block1, interblock1[0], block2 (c = 32): conv1, conv2, dil[0..2]
    assert "wp" not in enc.first[0] and all("wp" not in f for f in enc.last)
    assert not any(hasattr(b, "c3_windows") for b in blocks)       # the chained-window conv3 is built but switched off (slower)


def test_reference_weight_init_statistics():
    """Constructors reproduce the reference initialisation (stereodpnet/mainmodel.py:49-64 + the hourglass loops): N(0, sqrt(2 /
    (prod(kernel) * C_out))) for Conv2d / Conv3d / ConvTranspose3d -- including the re-randomised conv_offset of the deformable
    layers -- BatchNorm weight 1 / bias 0; conv biases and the D3D weight Parameter keep their defaults."""
    import math
    from dualpixelface_b200.runner import load_config, model_selector
    torch.manual_seed(1)
    model = model_selector(load_config("train_faceDP", "pytest", make_dirs=False))
    checked = 0
    for name, m in model.named_modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.Conv3d, torch.nn.ConvTranspose3d)) and m.weight.numel() >= 4096:
            want = math.sqrt(2.0 / (m.out_channels * math.prod(m.kernel_size)))
            assert abs(float(m.weight.std()) / want - 1.0) < 0.08, (name, float(m.weight.std()), want)
            assert abs(float(m.weight.mean())) < 0.1 * want, name
            checked += 1
        elif isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            assert float(m.weight.min()) == 1.0 == float(m.weight.max()) and float(m.bias.abs().max()) == 0.0
    assert checked > 60
    off = model.normal_estimator.deform_conv1.conv_offset
    assert float(off.weight.std()) > 0.02 and float(off.bias.abs().max()) == 0.0      # probe in SURVEY 8a-7: std 0.030, bias 0


def test_plan_caches_are_dropped_on_mode_change_and_move():
    """the eval-mode plans (packed weights + folded running statistics) must not survive train()/eval()/.to()."""
    from dualpixelface_b200.runner import load_config, model_selector
    model = model_selector(load_config("eval_faceDP", "pytest", make_dirs=False))
    def poison():
        model.aggregation._plan = {"stale": 1}
        model.normal_estimator._plan = {"stale": 1}
        model.cost_volume._packed = {"stale": 1}
        model.__dict__["_enc_fused"] = "stale"
    def clean():
        return (model.aggregation._plan is None and model.normal_estimator._plan is None and model.cost_volume._packed is None
                and "_enc_fused" not in model.__dict__)
    for change in (lambda: model.train(), lambda: model.eval(), lambda: model.to(torch.float32), lambda: model.float(),
                   lambda: model.load_state_dict(model.state_dict())):
        poison()
        change()
        assert clean()


def test_lightning_style_checkpoint_loads(tmp_path):
    """The reference's released checkpoints are pytorch_lightning files written after save_hyperparameters(): they pickle the
    config object under 'hyper_parameters'.  torch >= 2.6 defaults to weights_only=True, which rejects that; the loader falls
    back to an unpickler that stubs unknown classes and keeps only the tensors of 'state_dict'."""
    import sys, types
    from dualpixelface_b200.runner import load_checkpoint_state
    mod = types.ModuleType("config_fake_manager")
    class obj:                                             # stands in for config_.config_manager.obj
        def __init__(self, d):
            self.__dict__.update(d)
    obj.__module__ = "config_fake_manager"; obj.__qualname__ = "obj"
    mod.obj = obj
    sys.modules["config_fake_manager"] = mod
    sd = {"aggregation.dres0.0.0.weight": torch.randn(4, 3), "normal_estimator.grid": torch.zeros(2)}
    path = tmp_path / "checkpoint_epoch=03.ckpt"
    torch.save({"state_dict": sd, "epoch": 3, "hyper_parameters": {"option": obj({"mode": "train", "model": obj({"level": 8})})},
                "pytorch-lightning_version": "1.4.9"}, path)
    del sys.modules["config_fake_manager"]                  # the class is NOT importable at load time
    got = load_checkpoint_state(path)
    assert set(got) == set(sd) and torch.equal(got["aggregation.dres0.0.0.weight"], sd["aggregation.dres0.0.0.weight"])
    plain = tmp_path / "plain.ckpt"
    torch.save({"model": sd}, plain)
    assert set(load_checkpoint_state(plain)) == set(sd)
    torch.save({"weights": sd}, plain)
    with pytest.raises(NotImplementedError):
        load_checkpoint_state(plain)


def test_fuse_t2_weight_layout():
    """Transposed kind: the 27 tap matrices regrouped into 15 operands (runs of parity classes that read the same input shift)."""
    import torch
    from dualpixelface_b200.ops import _T2_DEPTH, _T2_RUNS, _t2_k, fuse_t2_weight, pack_conv_weight
    w = torch.arange(32 * 64 * 27, dtype=torch.float32).reshape(32, 64, 3, 3, 3) % 251            # [Cout,Cin,kd,kh,kw], bf16-exact values
    plain = pack_conv_weight(w)                                                                   # [27][8][32][8]
    fused = fuse_t2_weight(plain)
    assert fused.shape == plain.shape and torch.equal(fused.flatten().sort().values, plain.flatten().sort().values)
    flat, pos, ngroups, nslots = fused.flatten(), 0, 0, 0
    for _rd, kd in _T2_DEPTH:
        for oh, ow, classes in _T2_RUNS:
            n = len(classes) * 32
            grp = flat[pos: pos + 8 * n * 8].reshape(8, n, 8)                                     # [Cin/8][ncls*Npad][8]
            for i, c in enumerate(classes):
                tap = (kd * 3 + _t2_k(c >> 1, oh)) * 3 + _t2_k(c & 1, ow)
                assert torch.equal(grp[:, 32 * i: 32 * (i + 1)], plain[tap])
            pos += 8 * n * 8
            ngroups += 1
            nslots += len(classes)
    assert (ngroups, nslots, pos) == (15, 27, fused.numel())
    # out[2q + r] += in[q + o] * W[k]:  r = 0 -> (k 1, o 0);  r = 1 -> (k 0, o 1), (k 2, o 0)
    assert [_t2_k(0, 0), _t2_k(1, 1), _t2_k(1, 0)] == [1, 0, 2]


def test_pack_head_and_stem_weights():
    import torch
    from dualpixelface_b200.ops import pack_head_weight, pack_stem_weight
    w = torch.randn(1, 32, 3, 3, 3).to(torch.bfloat16).float()
    p = pack_head_weight(w)                                                                       # [4][32 taps][8]
    assert p.shape == (4, 32, 8) and p[:, 27:].abs().sum() == 0
    assert float(p[2, 13, 5]) == float(w[0, 2 * 8 + 5, 1, 1, 1])                                  # tap 13 = (1,1,1), channel 21
    ws = torch.randn(32, 3, 3, 3).to(torch.bfloat16).float()
    q = pack_stem_weight(ws)                                                                      # [80][32], k = tap*8 + ci
    assert q.shape == (80, 32) and q[72:].abs().sum() == 0 and q.reshape(10, 8, 32)[:, 3:].abs().sum() == 0
    assert float(q[(2 * 3 + 1) * 8 + 2, 17]) == float(ws[17, 2, 2, 1])                            # tap (kh 2, kw 1), ci 2, output 17


def test_psmnet_encoder_routing_finds_the_eligible_convs():
    """models.route_convs_to_tc on a BN-folded copy of PSMNet's encoder: firstconv 2-3, layer1 (6) and layer2 (31) = 39 3x3
    stride-1 convolutions with <= 96 channels go to the tcgen05 kernel, a directly following ReLU fused; the other biased ones
    stay on cuDNN behind CudnnConvBiasAct."""
    import copy
    import torch.nn as nn
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    from conftest import ROOT
    from dualpixelface_b200 import models as MM
    from dualpixelface_b200.runner import load_config, model_selector
    model = model_selector(load_config("eval_faceDP_psmnet", "t", root=ROOT, make_dirs=False), root=ROOT).eval()
    enc = copy.deepcopy(model.feature_extraction).eval()

    def fold(mod):
        for _, ch in list(mod.named_children()):
            if isinstance(ch, nn.Sequential) and len(ch) >= 2 and isinstance(ch[0], nn.Conv2d) and isinstance(ch[1], nn.BatchNorm2d):
                ch[0], ch[1] = fuse_conv_bn_eval(ch[0], ch[1]), nn.Identity()
            fold(ch)

    fold(enc)
    assert MM.route_convs_to_tc(enc) == 39
    tc = [m for m in enc.modules() if isinstance(m, MM.TCConv2dEval)]
    assert len(tc) == 39
    assert sum(m.relu for m in tc) == 2 + 3 + 15                       # firstconv 2-3, conv1 of layer1 (3) and of layer2 (15: its first block is stride 2)
    assert all(m.cout in (32, 64) and m.dil == 1 for m in tc)
    # every other biased convolution stays on cuDNN, bias-free, with ONE dpf_bias_act pass for bias + ReLU + skip connection:
    # stem, layer2.0 conv1 (stride 2), the two 1x1 downsamples, layer3 / layer4 (12), the four SPP branches, lastconv.0 = 21
    cd = [m for m in enc.modules() if isinstance(m, MM.CudnnConvBiasAct)]
    assert len(cd) == 21 and sum(m.relu for m in cd) == 1 + 1 + 6 + 4 + 1
    assert isinstance(enc.firstconv[0][0], MM.CudnnConvBiasAct) and enc.firstconv[0][0].stride == (2, 2)      # the stem
    assert isinstance(enc.layer3[0].conv1[0][0], MM.CudnnConvBiasAct) and enc.layer3[0].conv1[0][0].cout == 128
    left = [m for m in enc.modules() if isinstance(m, nn.Conv2d)]
    assert len(left) == 1 and left[0] is enc.lastconv[2] and left[0].bias is None                            # the bias-free 1x1 output conv


def test_encoder_batchnorm_is_plain_batchnorm_off_the_fused_path():
    """EncoderBatchNorm2d == nn.BatchNorm2d wherever the fused training path does not apply (CPU, eval, fp32)."""
    import torch
    from dualpixelface_b200.modules import EncoderBatchNorm2d
    torch.manual_seed(0)
    a, b = EncoderBatchNorm2d(16), torch.nn.BatchNorm2d(16)
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 16, 5, 7)
    for mode in (True, False):
        a.train(mode); b.train(mode)
        assert torch.equal(a(x), b(x))
    assert set(a.state_dict()) == set(b.state_dict())


def test_nnet_contract_and_eval_plan():
    """NNET: the reference's state_dict layout (473 keys), its input-size rule, the lazily registered pixel grid of
    released checkpoints, and the host-side weight re-layouts of the eval plan -- the 67-channel first layer of the normal module with
    its input channels reordered to [cost_in | coordinates | zero pad], the depth-halving (2,3,3) convolutions folded to 2-D
    64-channel ones, the 33 -> 40 channel pad of the first context layer."""
    from dualpixelface_b200.nnet import NNET, CONTEXT_DILATIONS
    from dualpixelface_b200.runner import load_config, model_selector
    shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_nnet.json").read_text()).items()}
    model = model_selector(load_config("eval_faceDP_nnet", "t", root=ROOT, make_dirs=False), root=ROOT).eval()
    assert isinstance(model, NNET)
    sd = model.state_dict()
    assert set(sd) == set(shapes) and len(sd) == 473
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    assert sd["normal_module.costrange"].flatten().tolist() == [-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0, 2.5]
    assert model.cost_volume.shifts == [-1, 0, 0, 0, 1, 1, 2, 2]
    assert model.convert_checkpoint_keys({"normal_module.grid": 1, "dres0.0.0.weight": 2}) == {"dres0.0.0.weight": 2}
    model.check_input_size(256, 256)
    with pytest.raises(ValueError):
        model.check_input_size(128, 256)                    # the encoder's 64 x 64 average pool
    assert not model.feature_extraction.branch_align_corners
    p = model._build()
    w = model.normal_module.wc0[0][0].weight
    conv0 = p["wc0.0"][0]
    assert conv0.cin == 96 and [(ln.x_coff, ln.cin) for ln in conv0.plan] == [(0, 32), (32, 32), (64, 32)]
    assert [tuple(wt.shape[:2]) + (d,) for wt, d in p["ctx_cudnn"]] == [(128, 40, 1), (128, 128, 2), (128, 128, 4), (96, 128, 8)]
    assert torch.equal(p["ctx_cudnn"][0][0][:, 33:].float(), torch.zeros(128, 7, 3, 3))
    assert [(c, d) for _, c, d in p["ctx_tc"]] == [(64, 16), (32, 1), (1, 1)]
    assert [(c, d) for _, c, d in p["n_convs"]] == list(zip((96, 96, 96, 64, 64, 32, 3), CONTEXT_DILATIONS))
    # depth-pair folding: conv3d over (kd, kh, kw) == conv2d over channels (kd * C + c)
    pool = model.normal_module.pool1[0][0]
    x = torch.randn(1, 32, 4, 6, 7)
    want = pool(x)                                           # [1,32,2,6,7]
    w2 = pool.weight.detach().permute(0, 2, 1, 3, 4).reshape(32, 64, 3, 3)
    pairs = x.permute(0, 2, 3, 4, 1).reshape(1, 2, 2, 6, 7, 32).permute(0, 1, 3, 4, 2, 5).reshape(2, 6, 7, 64)
    got = F.conv2d(pairs.permute(0, 3, 1, 2), w2, None, 1, 1).view(1, 2, 32, 6, 7).permute(0, 2, 1, 3, 4)
    assert torch.allclose(got, want, atol=1e-5)
    assert p["pools"][0][0].shape == (9, 8, 32, 8)          # 64 input channels in 8-channel pieces, 32 outputs
    # plans are dropped with the weights they were packed from
    model.train()
    assert model._plan is None


def test_nnet_coord_volume_matches_oracle():
    from dualpixelface_b200.nnet import NormalModule
    from dualpixelface_b200.runner import load_config
    from oracle import dpf_oracle as O
    opt = load_config("eval_faceDP_nnet", "t", root=ROOT, make_dirs=False)
    nm = NormalModule(opt, -4, 12)
    batch = synthetic_batch(2, 64, 96, training=False, seed=0)
    got = nm.coord_volume(batch["K"], batch["abvalue"], 16, 24)
    cr = torch.tensor(np.asarray(O.cost_range(-4, 12, 8)), dtype=torch.float32).view(1, -1, 1, 1).expand(2, -1, 16, 24)
    want = O.anm_coord_volume(cr, batch["K"], batch["abvalue"]).permute(0, 2, 1, 3, 4)
    assert got.shape == want.shape == (2, 3, 8, 16, 24)
    assert (got - want).abs().max().item() < 1e-5


def test_routed_encoder_fuses_the_skip_connections():
    """The BN-folded, routed copy of the SPP encoder (PSMNet / NNet) == the original encoder in eval mode: the 19 residual blocks
    whose conv2 runs on dpf_conv2d_tc_fwd hand their skip connection to that launch (TCConv2dEval.forward(x, residual)), the 6
    128-channel ones to the dpf_bias_act pass behind their cuDNN convolution (CudnnConvBiasAct); on CPU the stand-ins take their
    plain PyTorch fallback, which checks the same wiring."""
    import copy
    import torch.nn as nn
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    from dualpixelface_b200 import models as MM
    from dualpixelface_b200.runner import load_config, model_selector
    torch.manual_seed(3)
    model = model_selector(load_config("eval_faceDP_nnet", "t", root=ROOT, make_dirs=False), root=ROOT).eval()
    for m in model.feature_extraction.modules():                         # non-trivial running statistics
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    enc = copy.deepcopy(model.feature_extraction).eval()

    def fold(mod):
        for _, ch in list(mod.named_children()):
            if isinstance(ch, nn.Sequential) and len(ch) >= 2 and isinstance(ch[0], nn.Conv2d) and isinstance(ch[1], nn.BatchNorm2d):
                ch[0], ch[1] = fuse_conv_bn_eval(ch[0], ch[1]), nn.Identity()
            fold(ch)

    fold(enc)
    assert MM.route_convs_to_tc(enc) == 39
    blocks = [b for layer in (enc.layer1, enc.layer2, enc.layer3, enc.layer4) for b in layer]
    assert len(blocks) == 25 and all(getattr(b.conv2[0], "fuses_residual", False) for b in blocks)
    calls = []
    orig = {c: c.forward for c in (MM.TCConv2dEval, MM.CudnnConvBiasAct)}
    for c, f in orig.items():
        c.forward = (lambda f: lambda self, x, residual=None: (calls.append(residual is not None), f(self, x, residual))[1])(f)
    try:
        x = torch.randn(1, 3, 256, 256)
        with torch.no_grad():
            got, want = enc(x), model.feature_extraction(x)
    finally:
        for c, f in orig.items():
            c.forward = f
    assert sum(calls) == 25 and len(calls) == 39 + 21
    assert (got - want).abs().max().item() < 1e-4 * want.abs().max().item()
