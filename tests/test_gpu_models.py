"""GPU parity of the drop-in model classes (sm_100a hot path) against the CPU oracle, same seeded inputs and weights."""
import json

import pytest
import torch

from conftest import GOLDEN, ROOT
from dualpixelface_b200.synthetic import synth_state, synthetic_batch
from oracle import dpf_oracle as O

pytestmark = pytest.mark.gpu


def build(name, **model_overrides):
    from dualpixelface_b200.runner import load_config, model_selector
    opt = load_config("eval_faceDP" if name == "stereodpnet" else "eval_faceDP_psmnet", "pytest", root=ROOT, make_dirs=False)
    for k, v in model_overrides.items():
        setattr(opt.model, k, v)
    return model_selector(opt, root=ROOT)


def calibrated_state(name, batch):
    shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / f"state_keys_{name}.json").read_text()).items()}
    st = synth_state(shapes, seed=1)
    fwd = O.psmnet_forward if name == "psmnet" else O.stereodpnet_forward
    stats = {}
    with torch.no_grad():
        fwd(dict(batch), st, True, stats=stats)
    return O.calibrate_running_stats(st, stats), fwd


def to_cuda(batch):
    return {k: v.cuda() for k, v in batch.items()}


@pytest.mark.parametrize("name,hw", [("psmnet", (256, 256)), ("stereodpnet", (64, 96)), ("stereodpnet", (128, 160))])
def test_model_eval_parity(name, hw):
    batch = synthetic_batch(2, hw[0], hw[1], training=True, seed=0)
    st, fwd = calibrated_state(name, batch)
    stages = {}
    with torch.no_grad():
        want = fwd(dict(batch), st, False, stages=stages)
    model = build(name)
    model.load_state_dict(st, strict=False)
    model.cuda().eval()
    model.encoder_autocast = False            # fp32 cuDNN encoder: the comparison then isolates the bf16 hot path
    model.regression_layer.want_prob = True
    with torch.no_grad():
        got = model(to_cuda(batch))
    d_got, d_want = got["pred_depth"].float().cpu(), want["pred_depth"]
    err = (d_got - d_want).abs()
    rng = float(d_want.max() - d_want.min())
    print(f"{name} {hw}: disparity max err {err.max():.4f} mean {err.mean():.5f} (range {rng:.2f})")
    # north_star tolerance for the bf16 path: 2e-2 relative (= 0.32 px of the 16 px disparity span); held to <= 2x what a B200
    # measures instead: max 0.073-0.090 px, mean 0.0095-0.0102 px
    assert err.max().item() < 0.18 and err.mean().item() < 0.02
    assert (got["prob_depth"].float().cpu() - want["prob_depth"]).abs().max().item() < 2e-2
    fe = (got["ref_feature"].cpu() - want["ref_feature"]).abs().max().item()
    assert fe < 2e-2 * want["ref_feature"].abs().max().item()
    if name == "stereodpnet":
        n_err = (got["pred_normal"].float().cpu() - want["pred_normal"]).abs()
        print(f"   normal max err {n_err.max():.4f} mean {n_err.mean():.5f}")
        assert n_err.mean().item() < 7e-3 and n_err.max().item() < 0.125          # measured mean 0.0027-0.0034, max 0.051-0.062


def test_cpu_input_fails_loudly():
    model = build("psmnet").eval()
    with pytest.raises(RuntimeError):
        model(synthetic_batch(1, 256, 256))


def test_psmnet_default_path_bf16_encoder_on_tc_kernels():
    """PSMNet as benchmarked: BN-folded bf16 encoder whose 39 eligible 3x3 convolutions run on dpf_conv2d_tc_fwd (models.route_convs_to_tc)."""
    from dualpixelface_b200.models import TCConv2dEval
    batch = synthetic_batch(2, 256, 256, training=True, seed=0)
    st, fwd = calibrated_state("psmnet", batch)
    with torch.no_grad():
        want = fwd(dict(batch), st, False)
    model = build("psmnet")
    model.load_state_dict(st, strict=False)
    model.cuda().eval()
    TCConv2dEval.min_pixels = 0                     # the size threshold is a performance heuristic; exercise the kernels at this small size
    try:
        with torch.no_grad():
            got = model(to_cuda(batch))
    finally:
        TCConv2dEval.min_pixels = 100_000
    import os
    if os.environ.get("DPF_ENC_GENERIC_TC", "1") != "0":
        assert sum(isinstance(m, TCConv2dEval) for m in model._fused_encoder().modules()) == 39
    err = (got["pred_depth"].float().cpu() - want["pred_depth"]).abs()
    print(f"psmnet bf16 encoder: disparity max err {err.max():.4f} mean {err.mean():.5f}")
    # the 50-layer bf16 encoder is the error source here (cuDNN bf16 on the same layers: mean 0.173 / max 1.36 px; these kernels:
    # 0.164 / 1.22); thresholds <= 2x measured.  The fp32-encoder tests above isolate the hot path (mean 0.010 px).
    assert err.mean().item() < 0.33 and err.max().item() < 2.5


def test_model_default_path_bf16_encoder():
    """The bench configuration: BN-folded bf16 cuDNN encoder in front of the sm_100a hot path."""
    batch = synthetic_batch(2, 128, 160, training=True, seed=0)
    st, fwd = calibrated_state("stereodpnet", batch)
    with torch.no_grad():
        want = fwd(dict(batch), st, False)
    model = build("stereodpnet")
    model.load_state_dict(st, strict=False)
    model.cuda().eval()
    with torch.no_grad():
        got = model(to_cuda(batch))
    err = (got["pred_depth"].float().cpu() - want["pred_depth"]).abs()
    n_err = (got["pred_normal"].float().cpu() - want["pred_normal"]).abs()
    print(f"bf16 encoder: disparity max err {err.max():.4f} mean {err.mean():.5f}; normal mean {n_err.mean():.5f}")
    assert err.mean().item() < 0.11 and n_err.mean().item() < 0.02            # measured 0.055 px / 0.0103 (bf16 encoder included)


def test_psmnet_gwcnet_style():
    """cost_volume = 'gwcnet' (psmnet/modules.py:243-271): cat(concat volume, group-wise correlation volume) -> a 2C+G = 72-channel
    first aggregation layer (zero-padded to 96 on the conv engine).  group_num = 8: the shipped 40 does not divide the 32 feature
    channels and fails the reference's own assert (psmnet/modules.py:217).  Eval parity vs the oracle + a finite training step."""
    cfg = dict(O.PSM_CFG, cost_volume="gwcnet", group_num=8)
    shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_psmnet.json").read_text()).items()}
    shapes["aggregation.dres0.0.0.weight"] = (32, 72, 3, 3, 3)
    st = synth_state(shapes, seed=1)
    batch = synthetic_batch(2, 256, 256, training=True, seed=0)
    stats = {}
    with torch.no_grad():
        O.psmnet_forward(dict(batch), st, True, cfg=cfg, stats=stats)
        st = O.calibrate_running_stats(st, stats)
        want = O.psmnet_forward(dict(batch), st, False, cfg=cfg)
    model = build("psmnet", cost_volume="gwcnet", group_num=8)
    assert model.aggregation.dres0[0][0].weight.shape == (32, 72, 3, 3, 3)
    model.load_state_dict(st, strict=False)
    model.cuda().eval()
    model.encoder_autocast = False
    with torch.no_grad():
        got = model(to_cuda(batch))
    err = (got["pred_depth"].float().cpu() - want["pred_depth"]).abs()
    print(f"gwcnet (256, 256): disparity max err {err.max():.4f} mean {err.mean():.5f}")
    assert err.max().item() < 0.3 and err.mean().item() < 0.03
    model.train()
    res = model(to_cuda(batch))
    res["final_loss"].backward()
    g = model.aggregation.dres0[0][0].weight.grad
    assert g is not None and g.shape == (32, 72, 3, 3, 3) and torch.isfinite(g).all() and float(g[:, 64:].abs().max()) > 0
    with pytest.raises(ValueError):
        build("psmnet", cost_volume="gwcnet", group_num=40).cuda().eval()(to_cuda(batch))


@pytest.mark.parametrize("use_deform,use_sampling", [(False, True), (False, False), (True, False)])
def test_anm_variants(use_deform, use_sampling):
    """Non-default ANM configurations (normal_module.py:45-56,159-163,181-183): plain convbn_3d layers instead of the deformable
    ones, and all 8 levels instead of the 4 sampled ones.  The oracle's variants are pinned against the unmodified reference by
    tests/test_oracle_golden.py (fixtures of tests/golden/make_golden_variants.py)."""
    import numpy as np
    if use_deform:
        shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_stereodpnet.json").read_text()).items()}
    else:
        gold = np.load(GOLDEN / "variants.npz")
        shapes = {k: tuple(v) for k, v in json.loads(bytes(gold["sdp_nodeform_sampling/state_keys"]).decode()).items()}
    cfg = dict(O.SDP_CFG, use_deform=use_deform, use_sampling=use_sampling)
    st = synth_state(shapes, seed=1)
    batch = synthetic_batch(2, 64, 96, training=True, seed=0)
    stats = {}
    with torch.no_grad():
        O.stereodpnet_forward(dict(batch), st, True, cfg=cfg, stats=stats)
        st = O.calibrate_running_stats(st, stats)
        want = O.stereodpnet_forward(dict(batch), st, False, cfg=cfg)
    model = build("stereodpnet", use_deform=use_deform, use_sampling=use_sampling)
    assert set(model.state_dict()) - {"normal_estimator.costrange"} <= set(shapes) | {"normal_estimator.costrange"}
    model.load_state_dict(st, strict=False)
    model.cuda().eval()
    model.encoder_autocast = False
    with torch.no_grad():
        got = model(to_cuda(batch))
    n_err = (got["pred_normal"].float().cpu() - want["pred_normal"]).abs()
    print(f"ANM use_deform={use_deform} use_sampling={use_sampling}: normal max err {n_err.max():.4f} mean {n_err.mean():.5f}")
    assert n_err.mean().item() < 8e-3 and n_err.max().item() < 0.15
    model.train()                                             # and the training path of the variant runs end to end
    res = model(to_cuda(batch))
    res["final_loss"].backward()
    assert all(torch.isfinite(p.grad).all() for p in model.normal_estimator.parameters() if p.grad is not None)
    assert sum(p.grad is not None for p in model.normal_estimator.parameters()) >= 10


@pytest.mark.parametrize("training", [False, True])
def test_sdp_volume_per_level_shifts_with_fractional_phase(training):
    """cached_first_level=False -- the evidently intended behaviour of the ASM volume: one shift per level (-1, -0.5, ..., 2.5
    rows) incl. the fractional Fourier shifts (asm.py:63-75,112-125) -- vs the oracle, whose per-shift samples are pinned
    against the reference (tests/test_oracle_golden.py::test_subpixel_shift)."""
    shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_stereodpnet.json").read_text()).items()}
    st = synth_state(shapes, seed=1)
    g = torch.Generator().manual_seed(31)
    ref = torch.relu(torch.randn(2, 32, 16, 24, generator=g)).to(torch.bfloat16)
    tgt = torch.relu(torch.randn(2, 32, 16, 24, generator=g)).to(torch.bfloat16)
    with torch.no_grad():
        want = O.sdp_cost_volume(ref.float(), tgt.float(), st, "cost_volume", O.cost_range(-4, 12, 8), training, cached_first_level=False)
    assert not torch.equal(want[:, :, 0], want[:, :, 3])
    model = build("stereodpnet", predict_normal=False)
    model.load_state_dict(st, strict=False)
    cv = model.cost_volume.cuda().train(training)
    cv.cached_first_level = False
    with torch.no_grad():
        got = cv(ref.permute(0, 2, 3, 1).contiguous().cuda(), tgt.permute(0, 2, 3, 1).contiguous().cuda())
    err = (got.permute(0, 4, 1, 2, 3).float().cpu() - want).abs().max().item() / want.abs().max().item()
    print(f"per-level ASM volume (training={training}): max err {err:.4f} of the volume max")
    assert err < 2e-2
