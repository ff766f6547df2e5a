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
    assert err.max().item() < 2e-2 * 16.0 and err.mean().item() < 2e-3 * 16.0   # 2e-2 (bf16 path) of the 16 px disparity span (mindisp..maxdisp)
    assert (got["prob_depth"].float().cpu() - want["prob_depth"]).abs().max().item() < 2e-2
    fe = (got["ref_feature"].cpu() - want["ref_feature"]).abs().max().item()
    assert fe < 2e-2 * want["ref_feature"].abs().max().item()
    if name == "stereodpnet":
        n_err = (got["pred_normal"].float().cpu() - want["pred_normal"]).abs()
        print(f"   normal max err {n_err.max():.4f} mean {n_err.mean():.5f}")
        assert n_err.mean().item() < 2e-2 and n_err.max().item() < 0.15


def test_cpu_input_fails_loudly():
    model = build("psmnet").eval()
    with pytest.raises(RuntimeError):
        model(synthetic_batch(1, 256, 256))


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
    assert err.mean().item() < 2e-2 * 16.0 / 4 and n_err.mean().item() < 2e-2
