"""GPU: NNET on the sm_100a kernels == the oracle's restatement of src/model/nnet (which is pinned bit-exact
against the unmodified reference by tests/test_oracle_golden.py::test_oracle_nnet_forward_and_grads)."""
import json

import numpy as np
import pytest
import torch

from dualpixelface_b200.runner import load_config, model_selector
from dualpixelface_b200.synthetic import synth_state, synthetic_batch
from oracle import dpf_oracle as O

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _shapes():
    return {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_nnet.json").read_text()).items()}


def _model():
    return model_selector(load_config("eval_faceDP_nnet", "test", root=ROOT, make_dirs=False), root=ROOT)


def _load(model, st):
    res = model.load_state_dict(st, strict=False)
    assert list(res.missing_keys) == ["normal_module.costrange"] and not res.unexpected_keys       # a derived constant, not in synth_state


@pytest.mark.parametrize("hw", [(256, 256), (256, 384)])
def test_nnet_eval_parity(hw):
    batch = synthetic_batch(2, hw[0], hw[1], training=True, seed=0)
    st = synth_state(_shapes(), seed=1)
    stats = {}
    with torch.no_grad():
        O.nnet_forward(dict(batch), st, True, stats=stats)
        st = O.calibrate_running_stats(st, stats)
        want = O.nnet_forward(dict(batch), st, False)
    model = _model()
    _load(model, st)
    model.cuda().eval()
    model.encoder_autocast = False            # fp32 cuDNN encoder: the comparison then isolates the bf16 hot path (as test_gpu_models does)
    model.want_prob = True
    with torch.no_grad():
        got = model({k: v.cuda() for k, v in batch.items()})
    assert got["pred_depth"].shape == want["pred_depth"].shape and got["prob_depth"].shape == want["prob_depth"].shape
    assert got["pred_normal"].shape == want["pred_normal"].shape
    d = (got["pred_depth"].float().cpu() - want["pred_depth"]).abs()
    p_err = (got["prob_depth"].float().cpu() - want["prob_depth"]).abs().max().item()
    cos = (got["pred_normal"].float().cpu() * want["pred_normal"]).sum(2)                     # both are unit vectors
    n_err = (got["pred_normal"].float().cpu() - want["pred_normal"]).abs()
    span = float(want["pred_depth"].max() - want["pred_depth"].min())
    print(f"nnet {hw}: raw head max err {d[:, 0].max():.4f} mean {d[:, 0].mean():.5f}; refined head max err {d[:, 1].max():.4f} mean "
          f"{d[:, 1].mean():.5f} (output span {span:.2f}); prob max err {p_err:.5f}; normal max err {n_err.max():.4f} mean {n_err.mean():.5f}, "
          f"min cosine {cos.min():.5f}")
    # north_star tolerance for the bf16 path is 2e-2 relative to the 16 px disparity span (0.32 px); held to <= 2.5x what a B200
    # measures instead: disparity max 0.031 / 0.036 px, mean 0.0039 / 0.0054 px, probabilities 0.0009 / 0.0011, normal mean 0.0047 / 0.0044
    # (single pixels where the un-normalised normal is ~0 flip under F.normalize: max 0.44 / 1.0, hence the mean-based bounds)
    assert d.max().item() < 0.09 and d.mean().item() < 0.013 and p_err < 3e-3
    assert n_err.mean().item() < 0.012 and cos.mean().item() > 0.999
    fe = (got["ref_feature"].cpu() - want["ref_feature"]).abs().max().item()
    assert fe < 2e-2 * want["ref_feature"].abs().max().item()


def test_nnet_default_path_bf16_encoder():
    """The configuration a user runs: BN-folded bf16 encoder (39 of its convs on dpf_conv2d_tc_fwd) in front of the hot path.  The
    50-layer bf16 encoder is the error source here, as it is for PSMNet (tests/test_gpu_models.py: mean 0.16 px there)."""
    batch = synthetic_batch(2, 256, 256, training=True, seed=0)
    st = synth_state(_shapes(), seed=1)
    stats = {}
    with torch.no_grad():
        O.nnet_forward(dict(batch), st, True, stats=stats)
        st = O.calibrate_running_stats(st, stats)
        want = O.nnet_forward(dict(batch), st, False)
    model = _model()
    _load(model, st)
    model.cuda().eval()
    with torch.no_grad():
        got = model({k: v.cuda() for k, v in batch.items()})
    d = (got["pred_depth"].float().cpu() - want["pred_depth"]).abs()
    cos = (got["pred_normal"].float().cpu() * want["pred_normal"]).sum(2)
    print(f"nnet bf16 encoder: disparity max err {d.max():.4f} mean {d.mean():.5f}; normal mean cosine {cos.mean():.5f}")
    assert d.mean().item() < 0.2 and d.max().item() < 1.5 and cos.mean().item() > 0.96          # measured 0.090 / 0.64 / 0.985


def test_nnet_is_deterministic_and_prob_is_lazy():
    batch = {k: v.cuda() for k, v in synthetic_batch(1, 256, 320, training=True, seed=3).items()}
    model = _model().cuda().eval()
    with torch.no_grad():
        a = model(batch)
        b = model(batch)
    assert a["prob_depth"] is None                                   # materialised only when want_prob is set (240 MB per head at 1120x1680)
    assert torch.equal(a["pred_depth"], b["pred_depth"]) and torch.equal(a["pred_normal"], b["pred_normal"])
    assert a["pred_depth"].shape == (1, 2, 256, 320) and a["pred_normal"].shape == (1, 1, 3, 256, 320)


def test_nnet_training_step_vs_reference_gradients():
    """One training step (train-mode BatchNorm, smooth-L1 over the raw and the refined disparity + cosine loss): forward, losses and
    ten parameter gradients of the UNMODIFIED reference (tests/golden/model_nnet.npz, make_golden_nnet.py).  fp32 2-D parts
    (autocast off) so that the comparison isolates the bf16 3-D trunk and its backward kernels."""
    gold = np.load(GOLDEN / "model_nnet.npz")
    batch = synthetic_batch(2, 256, 256, training=True, seed=0)
    model = _model()
    _load(model, synth_state(_shapes(), seed=1))
    model.cuda().train()
    model.encoder_autocast = False
    res = model({k: v.cuda() for k, v in batch.items()})
    res["final_loss"].backward()
    d = (res["pred_depth"][..., ::2, ::2].detach().float().cpu() - torch.from_numpy(gold["train/pred_depth_s2"])).abs()
    n = (res["pred_normal"][..., ::2, ::2].detach().float().cpu() - torch.from_numpy(gold["train/pred_normal_s2"])).abs()
    msg = f"nnet train: pred_depth max err {d.max():.4f} mean {d.mean():.5f}; pred_normal max err {n.max():.4f} mean {n.mean():.5f}"
    for key in ("smoothL1_loss", "cosine_loss", "final_loss"):
        msg += f"; {key} {float(res[key].detach()):.5f} vs {float(gold['train/' + key]):.5f}"
    print(msg)
    assert d.mean().item() < 0.012 and d.max().item() < 0.08 and n.mean().item() < 0.015          # measured 0.0043 / 0.026 / 0.0050
    for key in ("smoothL1_loss", "cosine_loss", "final_loss"):
        assert abs(float(res[key].detach()) - float(gold["train/" + key])) < 1e-3 * abs(float(gold["train/" + key]))      # measured 1.6e-4
    params = dict(model.named_parameters())
    bad = []
    for key in [k[len("train/grad/"):] for k in gold.files if k.startswith("train/grad/")]:
        g, w = params[key].grad.float().cpu().flatten(), torch.from_numpy(gold[f"train/grad/{key}"]).flatten()
        cos = float(torch.dot(g, w) / (g.norm() * w.norm()).clamp_min(1e-20))
        rel = float((g - w).norm() / w.norm().clamp_min(1e-20))
        print(f"   grad {key}: cosine {cos:.4f}, relative L2 error {rel:.4f}")
        bad.append((key, cos, rel)) if not ((cos > 0.98 and rel < 0.25) or (key.startswith("normal_module") and cos > 0.85)) else None
    # measured: trunk / context / encoder gradients cosine >= 0.989, relative L2 <= 0.15; normal module cosine 0.905-0.956.
    # the cosine loss sits at its plateau (1.0006: unit normals against N(0,1) targets), so the normal module's gradients are tiny
    # (|g| ~ 5e-4) and dominated by the bf16 rounding of its input volume: looser bound there
    assert not bad, bad
