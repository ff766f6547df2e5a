"""GPU parity at the BASELINE.json configurations themselves ("What's weak" 1 and 3):

  config 1  PSMNet eval forward, 1 x 448x448                        vs the oracle (CPU, fp32)
  config 3  StereoDPNet training step (depth + normal losses), 1120x1680, one pair: forward, losses and probed gradients
            vs the oracle's torch code executed on the GPU in fp32 (TF32 off)
  config 4  PSMNet training step on 512x768 crops, batch 2           vs the oracle's torch code on the GPU in fp32
  run-to-run determinism of the inference path (bit-equal, or within the stated bound)

Thresholds are <= 2x the values measured on a B200 (printed by each test; recorded in DESIGN.md section 2).
"""
import json

import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from dualpixelface_b200.synthetic import synth_state, synthetic_batch
from oracle import dpf_oracle as O

pytestmark = pytest.mark.gpu


def _shapes(name):
    return {k: tuple(v) for k, v in json.loads((GOLDEN / f"state_keys_{name}.json").read_text()).items()}


def rel2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def cos(a, b):
    return F.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


@pytest.fixture()
def fp32_oracle_on_gpu():
    """Strict fp32 for the oracle's cuDNN / cuBLAS calls, and a memory-bounded trilinear gather for the D3D restatement: the
    oracle's own function, re-evaluated in the backward instead of keeping its 8 corner tensors per tap alive (at 1120x1680
    they are 27 x 2 x ~1 GB).  The arithmetic is untouched."""
    from torch.utils.checkpoint import checkpoint
    flags = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    orig = O._trilinear_gather
    O._trilinear_gather = lambda x, d, h, w: checkpoint(orig, x, d, h, w, use_reentrant=False)
    yield
    O._trilinear_gather = orig
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = flags
    torch.cuda.empty_cache()


def test_config1_psmnet_eval_448():
    from test_gpu_models import build, calibrated_state
    batch = synthetic_batch(1, 448, 448, training=True, seed=0)
    st, fwd = calibrated_state("psmnet", synthetic_batch(2, 256, 256, training=True, seed=0))
    with torch.no_grad():
        want = fwd(dict(batch), st, False)
    model = build("psmnet")
    model.load_state_dict(st, strict=False)
    model.cuda().eval()
    model.encoder_autocast = False
    with torch.no_grad():
        got = model({k: v.cuda() for k, v in batch.items()})
    err = (got["pred_depth"].float().cpu() - want["pred_depth"]).abs()
    print(f"config 1 (PSMNet 448x448 eval): disparity max err {err.max():.4f} px, mean {err.mean():.5f} px")
    assert got["pred_depth"].shape == (1, 1, 448, 448)
    assert err.max().item() < 0.30 and err.mean().item() < 0.04       # measured on a B200: 0.147 / 0.0196 px


def test_config3_stereodpnet_training_step_1120x1680(fp32_oracle_on_gpu):
    from test_gpu_models import build
    torch.cuda.empty_cache()
    st = synth_state(_shapes("stereodpnet"), seed=1)
    batch = {k: v.cuda() for k, v in synthetic_batch(1, 1120, 1680, training=True, seed=3).items()}
    probe = ["aggregation.classif3.2.weight", "aggregation.dres4.conv6.0.weight", "aggregation.dres3.conv3.0.0.weight",
             "aggregation.dres2.conv1.0.0.weight", "aggregation.dres1.0.0.weight", "aggregation.dres0.0.0.weight",
             "cost_volume.attention_layer.mask_convs.0.weight", "normal_estimator.deform_conv2.weight",
             "normal_estimator.deform_conv1.conv_offset.weight", "feature_extraction.lastconv.2.0.weight"]
    so = {k: v.cuda() for k, v in st.items()}
    for k in probe:
        so[k] = so[k].clone().requires_grad_(True)
    lin = torch.randn(1, 3, 1120, 1680, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    want = O.stereodpnet_forward(dict(batch), so, True, predict_normal=True)
    # gradient objective: smooth-L1 + a linear probe of the normals (the cosine loss is ill-conditioned at random init, see
    # test_gpu_training_sdp.py); the shipped losses themselves are compared as values
    (want["smoothL1_loss"] + 10.0 * (want["pred_normal"] * lin).mean()).backward()
    ref = {k: so[k].grad.clone() for k in probe}
    want = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in want.items() if k != "prob_depth"}
    del so
    torch.cuda.empty_cache()
    model = build("stereodpnet")
    model.load_state_dict(st, strict=False)
    model.cuda().train()
    model.encoder_autocast = False
    res = model(batch)
    (res["smoothL1_loss"] + 10.0 * (res["pred_normal"] * lin).mean()).backward()
    torch.cuda.synchronize()
    d_err = (res["pred_depth"].detach().float() - want["pred_depth"]).abs()
    n_err = (res["pred_normal"].detach().float() - want["pred_normal"]).abs()
    print(f"config 3 (1x1120x1680 train): disparity max err {d_err.max():.4f} px mean {d_err.mean():.5f}; normal max err "
          f"{n_err.max():.4f} mean {n_err.mean():.5f}; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    assert res["pred_depth"].shape == (1, 3, 1120, 1680) and res["pred_normal"].shape == (1, 1, 3, 1120, 1680)
    assert d_err.max().item() < 0.20 and d_err.mean().item() < 0.016 and n_err.mean().item() < 9.4e-3   # measured 0.098 / 0.0081 / 0.0047
    for name in ("smoothL1_loss", "cosine_loss", "final_loss"):
        g_, r_ = float(res[name].detach()), float(want[name])
        print(f"   {name}: {g_:.5f} vs {r_:.5f}")
        assert abs(g_ - r_) < 1.5e-3 * abs(r_)                       # measured 5.2e-4 relative
    params = dict(model.named_parameters())
    # Whole-network gradients of two forward passes that differ by bf16 rounding: ReLU-mask flips dominate (the backward kernels
    # themselves are held to 0.999 / 2e-2 with teacher-forced activations in test_gpu_teacher_forced.py).  Floors = measured on
    # a B200 minus a margin: aggregation 0.998-1.000, mask conv 0.990, encoder lastconv 0.986, normal branch 0.84-0.89 (its
    # gradients are ill-conditioned in the branch input at random init, see tools/anm_grad_sensitivity.py).
    for k in probe:
        c, r = cos(params[k].grad, ref[k]), rel2(params[k].grad, ref[k])
        print(f"   grad {k}: cosine {c:.4f}, relative L2 error {r:.4f}")
        floor = 0.75 if k.startswith("normal_estimator.") else (0.996 if k.startswith("aggregation.") else 0.975)
        assert c > floor, (k, c)


def test_config4_psmnet_training_step_512x768(fp32_oracle_on_gpu):
    from test_gpu_models import build
    st = synth_state(_shapes("psmnet"), seed=1)
    batch = {k: v.cuda() for k, v in synthetic_batch(2, 512, 768, training=True, seed=4).items()}
    probe = ["aggregation.classif3.2.weight", "aggregation.dres4.conv6.0.weight", "aggregation.dres3.conv5.0.weight",
             "aggregation.dres2.conv1.0.0.weight", "aggregation.dres0.0.0.weight", "feature_extraction.lastconv.2.weight",
             "feature_extraction.firstconv.0.0.weight"]
    so = {k: v.cuda() for k, v in st.items()}
    for k in probe:
        so[k] = so[k].clone().requires_grad_(True)
    want = O.psmnet_forward(dict(batch), so, True)
    want["final_loss"].backward()
    model = build("psmnet")
    model.load_state_dict(st, strict=False)
    model.cuda().train()
    model.encoder_autocast = False
    res = model(batch)
    res["final_loss"].backward()
    torch.cuda.synchronize()
    d_err = (res["pred_depth"].detach().float() - want["pred_depth"].detach()).abs()
    print(f"config 4 (PSMNet 2x512x768 train): disparity max err {d_err.max():.4f} px mean {d_err.mean():.5f}; "
          f"loss {float(res['final_loss'].detach()):.5f} vs {float(want['final_loss'].detach()):.5f}")
    assert res["pred_depth"].shape == (2, 3, 512, 768)
    assert d_err.max().item() < 0.18 and d_err.mean().item() < 0.015        # measured 0.087 / 0.0073 px
    assert abs(float(res["final_loss"].detach()) - float(want["final_loss"].detach())) < 1.2e-3 * float(want["final_loss"].detach())
    params = dict(model.named_parameters())
    for k in probe:
        c, r = cos(params[k].grad, so[k].grad), rel2(params[k].grad, so[k].grad)
        print(f"   grad {k}: cosine {c:.4f}, relative L2 error {r:.4f}")
        assert c > (0.94 if "firstconv" in k else 0.993), (k, c)          # measured: >= 0.9969 (firstconv 0.961: deepest layer)


def test_inference_is_run_to_run_deterministic():
    """Two forward passes of the same model on the same input.  Several MMA-issuing warps accumulate into one TMEM tile
    (conv3d_tc.cu) in round 1 and the InstanceNorm statistics used fp32 atomics: run-to-run differences of up to 0.26 px were
    measured.  Round 2: one MMA-issuing thread per CTA (fixed accumulation order, DPF_CONV_ISSUERS=1 default) and a two-pass
    statistics reduction without atomics -> the inference path must be BIT-IDENTICAL from run to run (north_star: bit-exact
    disparity-index selection)."""
    from test_gpu_models import build, calibrated_state
    small = synthetic_batch(2, 128, 160, training=True, seed=0)
    st, _ = calibrated_state("stereodpnet", small)
    model = build("stereodpnet")
    model.load_state_dict(st, strict=False)
    model.cuda().eval()
    batch = {k: v.cuda() for k, v in synthetic_batch(2, 448, 672, training=True, seed=1).items()}
    outs = []
    with torch.no_grad():
        for _ in range(3):
            o = model(batch)
            outs.append((o["pred_depth"].clone(), o["pred_normal"].clone()))
    torch.cuda.synchronize()
    dd = max((outs[0][0] - o[0]).abs().max().item() for o in outs[1:])
    dn = max((outs[0][1] - o[1]).abs().max().item() for o in outs[1:])
    print(f"determinism over 3 runs: max |d disparity| {dd:.3e} px, max |d normal| {dn:.3e}; bit-equal: {dd == 0.0 and dn == 0.0}")
    assert dd == 0.0 and dn == 0.0


@pytest.mark.parametrize("cin,cout,kind", [(32, 32, 0), (64, 32, 0), (32, 64, 1), (64, 32, 2)])
def test_conv_kernel_run_to_run(cin, cout, kind):
    """The same convolution launched 4 times must give bit-identical outputs (single MMA issuer, fixed accumulation order)."""
    from dualpixelface_b200.layers import TCConv3d
    g = torch.Generator(device="cuda").manual_seed(3)
    shape = (2, 4, 70, 105) if kind == 2 else (2, 8, 140, 210)
    x = torch.randn(*shape, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(*((cin, cout) if kind == 2 else (cout, cin)), 3, 3, 3, device="cuda", generator=g) * 0.05
    conv = TCConv3d(w, kind, transposed=kind == 2)
    ys = [conv(x).float() for _ in range(4)]
    torch.cuda.synchronize()
    diff = max((ys[0] - y).abs().max().item() for y in ys[1:])
    nd = max(int((ys[0] != y).sum()) for y in ys[1:])
    print(f"conv kind {kind} {cin}->{cout}: max run-to-run difference {diff:.3e} ({nd} of {ys[0].numel()} elements differ)")
    assert diff == 0.0 and nd == 0
