"""GPU parity of the StereoDPNet training step (ASM volume + aggregation + regression, fwd AND bwd) against the CPU oracle."""
import json

import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from dualpixelface_b200.synthetic import synth_state, synthetic_batch
from oracle import dpf_oracle as O

pytestmark = pytest.mark.gpu


def rel2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def test_asm_volume_train_fwd_bwd():
    """CostVolumeSDP in train mode vs the oracle's autograd (same seeded features / weights)."""
    from test_gpu_models import build
    shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_stereodpnet.json").read_text()).items()}
    st = synth_state(shapes, seed=1)
    g = torch.Generator().manual_seed(31)
    ref = torch.relu(torch.randn(2, 32, 16, 24, generator=g)).to(torch.bfloat16)
    tgt = torch.relu(torch.randn(2, 32, 16, 24, generator=g)).to(torch.bfloat16)
    keys = ["cost_volume.attention_layer.mask_convs.0.weight", "cost_volume.attention_layer.mask_convs.1.weight",
            "cost_volume.attention_layer.mask_convs.1.bias", "cost_volume.attention_layer.mask_convs.3.0.weight",
            "cost_volume.attention_layer.mask_convs.3.1.weight", "cost_volume.attention_layer.mask_convs.3.1.bias"]
    so = dict(st)
    for k in keys:
        so[k] = st[k].clone().requires_grad_(True)
    r, t = ref.float().requires_grad_(True), tgt.float().requires_grad_(True)
    vol = O.sdp_cost_volume(r, t, so, "cost_volume", O.cost_range(-4, 12, 8), True)          # [B,2C,D,H,W]
    dvol = torch.randn(vol.shape, generator=g).to(torch.bfloat16)
    vol.backward(dvol.float())
    model = build("stereodpnet", predict_normal=False)
    model.load_state_dict(st, strict=False)
    cv = model.cost_volume.cuda().train()
    rg = ref.permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True)
    tg = tgt.permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True)
    out = cv(rg, tg)                                                                          # [B,D,H,W,2C]
    out.backward(dvol.permute(0, 2, 3, 4, 1).contiguous().cuda())
    torch.cuda.synchronize()
    assert rel2(out.permute(0, 4, 1, 2, 3), vol) < 2e-2
    assert rel2(rg.grad.permute(0, 3, 1, 2), r.grad) < 5e-2 and rel2(tg.grad.permute(0, 3, 1, 2), t.grad) < 5e-2
    params = dict(cv.named_parameters())
    for k in keys:
        name = k[len("cost_volume."):]
        name = name.replace("mask_convs.3.1.", "normalize.") if name not in params else name
        e = rel2(params[name].grad, so[k].grad)
        print(f"   {k}: relative L2 grad error {e:.4f}")
        assert e < 6e-2


def test_stereodpnet_training_step_depth_only():
    """STEREODPNET (predict_normal=false) fwd+bwd on the sm_100a path vs the oracle's autograd."""
    from test_gpu_models import build
    shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_stereodpnet.json").read_text()).items()}
    st = synth_state(shapes, seed=1)
    batch = synthetic_batch(2, 64, 96, training=True, seed=0)
    probe = ["aggregation.dres0.0.0.weight", "aggregation.classif3.2.weight", "cost_volume.attention_layer.mask_convs.0.weight",
             "feature_extraction.lastconv.2.0.weight"]
    so = dict(st)
    for k in probe:
        so[k] = st[k].clone().requires_grad_(True)
    want = O.stereodpnet_forward(dict(batch), so, True, predict_normal=False)
    want["final_loss"].backward()
    model = build("stereodpnet", predict_normal=False)
    model.load_state_dict(st, strict=False)
    model.cuda().train()
    model.encoder_autocast = False
    res = model({k: v.cuda() for k, v in batch.items()})
    res["final_loss"].backward()
    torch.cuda.synchronize()
    d_err = (res["pred_depth"].detach().float().cpu() - want["pred_depth"].detach()).abs()
    print(f"SDP train pred_depth max err {d_err.max():.4f} mean {d_err.mean():.5f}; loss {float(res['final_loss'].detach()):.5f} "
          f"vs {float(want['final_loss'].detach()):.5f}")
    assert res["pred_depth"].shape[1] == 3 and res["pred_normal"] is None
    assert d_err.max().item() < 2e-2 * 16.0 and d_err.mean().item() < 2e-3 * 16.0
    assert abs(float(res["final_loss"].detach()) - float(want["final_loss"].detach())) < 2e-2 * float(want["final_loss"].detach())
    params = dict(model.named_parameters())
    for k in probe:
        got, ref = params[k].grad.float().cpu(), so[k].grad
        cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        print(f"   grad {k}: cosine {cos:.4f}, relative L2 error {rel2(got, ref):.4f}")
        assert cos > 0.94


def test_normal_branch_training_fails_loudly():
    from test_gpu_models import build
    model = build("stereodpnet").cuda().train()
    with pytest.raises(NotImplementedError):
        model({k: v.cuda() for k, v in synthetic_batch(2, 64, 96, training=True).items()})
