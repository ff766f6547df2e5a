"""CPU: pin the oracle against fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from dualpixelface_b200.synthetic import synth_state, synthetic_batch
from oracle import dpf_oracle as O

from conftest import GOLDEN

CR = O.cost_range(-4, 12, 8)


def feat(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(*shape, generator=g))


def close(a, b, tol=1e-5):
    a = torch.as_tensor(np.asarray(a)).float()
    b = b.detach().float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, a.abs().max().item()), err


def test_cost_range_and_bins():
    assert CR.tolist() == [-1, -0.5, 0, 0.5, 1, 1.5, 2, 2.5]
    assert [int(c) for c in CR] == [-1, 0, 0, 0, 1, 1, 2, 2]          # int() truncation, psmnet/modules.py:229
    bins = O.disparity_bins(-4, 12, 8)
    assert bins[0] == -4 and bins[-1] == 11.5 and len(bins) == 32


@pytest.mark.parametrize("disp", [-1.0, -0.5, 0.5, 1.0, 2.5])
@pytest.mark.parametrize("direction", ["forward", "backward"])
def test_subpixel_shift(golden_stages, disp, direction):
    x = feat((2, 4, 16, 24), 11)
    got = torch.stack(O.subpixel_samples(x, disp, direction), -1)
    close(golden_stages[f"shift/{disp}/{direction}"], got, 1e-6)


def test_psm_volumes(golden_stages):
    ref, tgt = feat((2, 32, 12, 10), 21), feat((2, 32, 12, 10), 22)
    assert np.array_equal(golden_stages["psm/concat"], O.psm_concat_volume(ref, tgt, CR).numpy())
    assert np.array_equal(golden_stages["psm/gwc8"], O.psm_gwc_volume(ref, tgt, CR, 8).numpy())
    # difference volume (StereoNet) = concat halves subtracted, row windows identical
    v = O.psm_concat_volume(ref, tgt, CR)
    assert torch.equal(O.diff_volume(ref, tgt, CR), v[:, :32] - v[:, 32:])


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_sdp_volume_bugcompat(golden_stages, state_shapes, mode):
    st = synth_state(state_shapes["stereodpnet"], seed=1)
    ref, tgt = feat((2, 32, 16, 24), 31), feat((2, 32, 16, 24), 32)
    v = O.sdp_cost_volume(ref, tgt, st, "cost_volume", CR, mode == "train")
    close(golden_stages[f"sdp/volume/{mode}"], v[:, :, :2], 1e-5)
    assert golden_stages[f"sdp/volume_all_equal/{mode}"].all()
    assert all(torch.equal(v[:, :, 0], v[:, :, i]) for i in range(8))
    v2 = O.sdp_cost_volume(ref, tgt, st, "cost_volume", CR, mode == "train", cached_first_level=False)
    assert not torch.equal(v2[:, :, 0], v2[:, :, 3])


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_aggregation(golden_stages, state_shapes, mode):
    st = synth_state(state_shapes["stereodpnet"], seed=1)
    g = torch.Generator().manual_seed(41)
    vol = torch.relu(torch.randn(1, 64, 8, 16, 16, generator=g))
    costs, outs = O.aggregation(vol, st, "aggregation", mode == "train")
    for i, c in enumerate(costs):
        close(golden_stages[f"agg/{mode}/cost{3 - i}"], c, 1e-5)
    close(golden_stages[f"agg/{mode}/out3"], outs[0], 1e-5)


def test_regression(golden_stages):
    g = torch.Generator().manual_seed(51)
    cost_full = torch.randn(2, 32, 20, 28, generator=g) * 2.0
    d, p = O.regression(cost_full, O.disparity_bins(-4, 12, 8))
    close(golden_stages["regress/disp"], d, 1e-6)
    close(golden_stages["regress/prob"], p, 1e-6)


def test_anm_pieces(golden_stages):
    g = torch.Generator().manual_seed(61)
    cost = torch.randn(2, 8, 6, 10, 12, generator=g)
    dq = torch.rand(2, 1, 10, 12, generator=g) * 4.2 - 1.3
    crt = torch.as_tensor(CR, dtype=torch.float32).view(1, -1, 1, 1)
    idx = O.anm_select_levels(dq, crt, 4)
    sel_cost = torch.gather(cost, 1, idx.unsqueeze(2).expand(-1, -1, 6, -1, -1))
    sel_disp = torch.gather(crt.expand(2, 8, 10, 12), 1, idx)
    assert np.array_equal(golden_stages["anm/sel_cost"], sel_cost.numpy())
    assert np.array_equal(golden_stages["anm/sel_disp"], sel_disp.numpy())
    batch = synthetic_batch(2, 40, 48, seed=3)
    close(golden_stages["anm/coord"], O.anm_coord_volume(sel_disp, batch["K"], batch["abvalue"]), 1e-6)


def test_losses(golden_stages):
    b = synthetic_batch(2, 16, 24, training=True, seed=5)
    g = torch.Generator().manual_seed(71)
    pd = torch.randn(2, 3, 16, 24, generator=g) * 3
    pn = torch.randn(2, 1, 3, 16, 24, generator=g)
    mask = torch.as_tensor(golden_stages["loss/mask"])
    l1 = O.smooth_l1_multi(pd, b["disp"], mask, (1.0, 0.7, 0.5))
    lc = O.cosine_normal_loss(pn, b["normal"], mask)
    close(golden_stages["loss/smoothL1"], l1, 1e-6)
    close(golden_stages["loss/cosine"], lc, 1e-6)
    close(golden_stages["loss/final"], l1 + lc, 1e-6)


def test_deform_conv_zero_offset_is_dense_conv():
    """D3D restatement pin (the compiled reference op cannot run here): zero offsets == F.conv3d."""
    g = torch.Generator().manual_seed(81)
    x = torch.randn(1, 5, 4, 6, 7, generator=g)
    w = torch.randn(6, 5, 3, 3, 3, generator=g) * 0.1
    b = torch.randn(6, generator=g)
    off = torch.zeros(1, 81, 4, 6, 7)
    y = O.deform_conv3d(x, off, w, b)
    assert torch.allclose(y, torch.nn.functional.conv3d(x, w, b, padding=1), atol=1e-5)
    # integer offsets == shifted dense taps; fractional offsets stay finite and differentiable
    off2 = (torch.rand(1, 81, 4, 6, 7, generator=g) - 0.5).requires_grad_(True)
    y2 = O.deform_conv3d(x, off2, w, b)
    y2.sum().backward()
    assert torch.isfinite(off2.grad).all() and off2.grad.abs().sum() > 0


@pytest.mark.parametrize("name", ["psmnet", "stereodpnet"])
def test_whole_model(state_shapes, name):
    gold = np.load(GOLDEN / f"model_{name}.npz")
    st = synth_state(state_shapes[name], seed=1)
    h, w = (256, 256) if name == "psmnet" else (64, 96)
    batch = synthetic_batch(2, h, w, training=True, seed=0)
    fwd = O.psmnet_forward if name == "psmnet" else O.stereodpnet_forward
    stats = {}
    with torch.no_grad():
        res_t = fwd(dict(batch), st, True, stats=stats)
        res_e = fwd(dict(batch), O.calibrate_running_stats(st, stats), False)
    for tag, res in (("train", res_t), ("eval", res_e)):
        for key in ("pred_depth", "pred_normal", "ref_feature", "smoothL1_loss", "cosine_loss", "final_loss"):
            gk = f"{tag}/{key}"
            if gk in gold.files:
                close(gold[gk], res[key], 2e-5)
        close(gold[f"{tag}/prob_depth_sub"], res["prob_depth"][..., ::8, ::8], 2e-5)
    assert res_e["pred_depth"].shape[1] == 1 and res_t["pred_depth"].shape[1] == 3


def test_whole_model_grads_psmnet(state_shapes):
    """Backward of the oracle == backward of the reference (autograd through the same ops)."""
    gold = np.load(GOLDEN / "model_psmnet.npz")
    st = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v)
          for k, v in synth_state(state_shapes["psmnet"], seed=1).items()}
    batch = synthetic_batch(2, 256, 256, training=True, seed=0)
    res = O.psmnet_forward(dict(batch), st, True)
    res["final_loss"].backward()
    for key in ("aggregation.dres0.0.0.weight", "aggregation.classif3.2.weight",
                "aggregation.dres4.conv6.0.weight", "feature_extraction.firstconv.0.0.weight"):
        close(gold[f"train/grad/{key}"], st[key].grad, 1e-4)


# ---- non-default configurations, fixtures from tests/golden/make_golden_variants.py (the unmodified reference with overrides) ----
def _variant_case(tag, name, cfg):
    import json
    gold = np.load(GOLDEN / "variants.npz")
    shapes = {k: tuple(v) for k, v in json.loads(bytes(gold[f"{tag}/state_keys"]).decode()).items()}
    st = synth_state(shapes, seed=1)
    hw = (256, 256) if name == "psmnet" else (64, 96)
    batch = synthetic_batch(2, hw[0], hw[1], training=True, seed=0)
    fwd = O.psmnet_forward if name == "psmnet" else O.stereodpnet_forward
    stats = {}
    with torch.no_grad():
        fwd(dict(batch), st, True, cfg=cfg, stats=stats)
        st = O.calibrate_running_stats(st, stats)
        res = fwd(dict(batch), st, False, cfg=cfg)
    return gold, res, shapes


@pytest.mark.parametrize("sampling", [True, False])
def test_oracle_anm_without_deformable_convs(sampling):
    """use_deform=false (original_conv) x use_sampling=true/false == the reference's STEREODPNET with those overrides."""
    tag = f"sdp_nodeform_{'sampling' if sampling else 'nosampling'}"
    gold, res, shapes = _variant_case(tag, "stereodpnet", dict(O.SDP_CFG, use_deform=False, use_sampling=sampling))
    assert "normal_estimator.original_conv.0.0.weight" in shapes and not any("deform_conv" in k for k in shapes)
    close(gold[f"{tag}/pred_depth"], res["pred_depth"], 1e-5)
    close(gold[f"{tag}/pred_normal"], res["pred_normal"], 1e-5)


def test_oracle_gwcnet_style():
    gold, res, shapes = _variant_case("psm_gwcnet8", "psmnet", dict(O.PSM_CFG, cost_volume="gwcnet", group_num=8))
    assert shapes["aggregation.dres0.0.0.weight"] == (32, 72, 3, 3, 3)
    close(gold["psm_gwcnet8/pred_depth"], res["pred_depth"], 1e-5)


# ---- StereoNet, fixtures from tests/golden/make_golden_stereonet.py (the unmodified reference) ----
def _stereonet_shapes():
    import json
    return {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_stereonet.json").read_text()).items()}


def test_oracle_stereonet_forward():
    """Oracle == the reference's STEREONET in train mode (batch statistics, loss) and in eval mode with calibrated statistics."""
    gold = np.load(GOLDEN / "model_stereonet.npz")
    st = synth_state(_stereonet_shapes(), seed=1)
    batch = synthetic_batch(2, 64, 96, training=True, seed=0)
    stats = {}
    with torch.no_grad():
        res_t = O.stereonet_forward(dict(batch), st, True, stats=stats)
        res_e = O.stereonet_forward(dict(batch), O.calibrate_running_stats(st, stats), False)
    close(gold["train/pred_depth"], res_t["pred_depth"], 2e-5)
    close(gold["train/final_loss"], res_t["final_loss"], 2e-5)
    for key in ("pred_depth", "prob_depth", "ref_feature"):
        close(gold[f"eval/{key}"], res_e[key], 2e-5)
    assert res_e["pred_depth"].shape == (2, 2, 64, 96) and res_e["prob_depth"].shape == (2, 1, 8, 8, 12)


def test_oracle_stereonet_grads():
    gold = np.load(GOLDEN / "model_stereonet.npz")
    st = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v)
          for k, v in synth_state(_stereonet_shapes(), seed=1).items()}
    res = O.stereonet_forward(dict(synthetic_batch(2, 64, 96, training=True, seed=0)), st, True)
    res["final_loss"].backward()
    for key in [k[len("train/grad/"):] for k in gold.files if k.startswith("train/grad/")]:
        close(gold[f"train/grad/{key}"], st[key].grad, 1e-4)
    assert st["feature_extraction.residual_blocks.0.conv2.0.weight"].grad is None      # constructed but never applied (modules.py:23)


# ---- NNet, fixtures from tests/golden/make_golden_nnet.py (the unmodified reference) ----
def _nnet_shapes():
    import json
    return {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_nnet.json").read_text()).items()}


def test_oracle_nnet_forward_and_grads():
    """Oracle == the reference's NNET: train mode (batch statistics, both losses, ten parameter gradients spread over the encoder,
    the 3-D trunk, the context refinement and the normal module) and eval mode with calibrated statistics."""
    gold = np.load(GOLDEN / "model_nnet.npz")
    st0 = synth_state(_nnet_shapes(), seed=1)
    batch = synthetic_batch(2, 256, 256, training=True, seed=0)
    st = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v) for k, v in st0.items()}
    stats = {}
    res_t = O.nnet_forward(dict(batch), st, True, stats=stats)
    res_t["final_loss"].backward()
    close(gold["train/pred_depth_s2"], res_t["pred_depth"][..., ::2, ::2].detach(), 2e-5)
    close(gold["train/pred_normal_s2"], res_t["pred_normal"][..., ::2, ::2].detach(), 2e-5)
    for key in ("smoothL1_loss", "cosine_loss", "final_loss"):
        close(gold[f"train/{key}"], res_t[key].detach(), 2e-5)
    for key in [k[len("train/grad/"):] for k in gold.files if k.startswith("train/grad/")]:
        close(gold[f"train/grad/{key}"], st[key].grad, 1e-4)
    with torch.no_grad():
        res_e = O.nnet_forward(dict(batch), O.calibrate_running_stats(st0, stats), False)
    for key in ("pred_depth", "pred_normal", "ref_feature"):
        close(gold[f"eval/{key}"], res_e[key], 2e-5)
    close(gold["eval/prob_depth_s8"], res_e["prob_depth"][..., ::8, ::8], 2e-5)
    assert res_e["pred_depth"].shape == (2, 2, 256, 256) and res_e["pred_normal"].shape == (2, 1, 3, 256, 256)
