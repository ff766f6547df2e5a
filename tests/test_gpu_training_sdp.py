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
    assert d_err.max().item() < 0.15 and d_err.mean().item() < 0.014             # measured on a B200: 0.0745 / 0.0070 px
    assert abs(float(res["final_loss"].detach()) - float(want["final_loss"].detach())) < 2e-2 * float(want["final_loss"].detach())
    params = dict(model.named_parameters())
    for k in probe:
        got, ref = params[k].grad.float().cpu(), so[k].grad
        cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        print(f"   grad {k}: cosine {cos:.4f}, relative L2 error {rel2(got, ref):.4f}")
        assert cos > 0.94


ANM_PROBE = ["normal_estimator.deform_conv1.weight", "normal_estimator.deform_conv1.conv_offset.weight",
             "normal_estimator.deform_conv1.conv_offset.bias", "normal_estimator.act1.0.weight", "normal_estimator.act1.0.bias",
             "normal_estimator.deform_conv2.weight", "normal_estimator.deform_conv2.conv_offset.weight",
             "normal_estimator.act2.0.weight", "normal_estimator.n_convs.0.0.weight", "normal_estimator.n_convs.5.0.weight"]


def test_anm_training_fwd_bwd():
    """Normal branch alone in train mode (gather, offset convs, D3D fwd/bwd, BN, n_convs, tail) vs the oracle's autograd."""
    from test_gpu_models import build
    shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_stereodpnet.json").read_text()).items()}
    st = synth_state(shapes, seed=1)
    batch = synthetic_batch(2, 64, 96, training=True, seed=0)
    g = torch.Generator().manual_seed(77)
    out3 = torch.relu(torch.randn(2, 32, 8, 16, 24, generator=g)).to(torch.bfloat16)
    disp = torch.rand(2, 64, 96, generator=g) * 14.0 - 3.0
    model = build("stereodpnet")
    model.load_state_dict(st, strict=False)
    anm = model.normal_estimator.cuda().train()
    so = dict(st)
    for k in ANM_PROBE:
        so[k] = st[k].clone().requires_grad_(True)
    o3 = out3.float().requires_grad_(True)
    want = O.anm_forward(o3, disp, batch["K"], batch["abvalue"], so, "normal_estimator", anm.levels, True, anm.k)
    dn = torch.randn(want.shape, generator=g)
    want.backward(dn)
    og = out3.permute(0, 2, 3, 4, 1).contiguous().cuda().requires_grad_(True)
    normals, off1, off2 = anm([og], [disp.cuda()], {k: v.cuda() for k, v in batch.items()})
    normals[0].backward(dn.cuda())
    torch.cuda.synchronize()
    err = (normals[0].detach().cpu() - want.detach()).abs()
    print(f"ANM train normal max err {err.max():.4f} mean {err.mean():.5f}")
    assert err.max().item() < 5e-2 and err.mean().item() < 5e-3
    # bf16 rounding flips a fraction of the LeakyReLU / ReLU masks of the six n_convs (tools/debug_anm_bwd.py: the gradient
    # already differs by 16 % relative L2 at f2, before any D3D kernel, and stays there down to out3), hence cosine floors
    e = rel2(og.grad.permute(0, 4, 1, 2, 3), o3.grad)
    cos = F.cosine_similarity(og.grad.permute(0, 4, 1, 2, 3).float().cpu().flatten(), o3.grad.flatten(), dim=0).item()
    print(f"   d out3: cosine {cos:.4f}, relative L2 error {e:.4f}")
    assert cos > 0.97
    params = dict(model.named_parameters())
    for k in ANM_PROBE:
        got, ref = params[k].grad.float().cpu(), so[k].grad
        cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        print(f"   grad {k}: cosine {cos:.4f}, relative L2 error {rel2(got, ref):.4f}")
        assert cos > 0.97


def test_stereodpnet_training_step_with_normals():
    """The shipped training configuration (depth + normal heads) fwd+bwd on the sm_100a path vs the oracle's autograd.

    The losses are compared directly.  For the GRADIENT comparison the cosine normal loss is replaced by a linear probe of the
    predicted normals: at random initialisation |pred_normal| ~ 0.07, and the normalisation inside the cosine loss turns the
    2e-3 bf16 forward error into a 34 % relative error of dL/dnormal itself (cosine 0.94, measured) before any backward kernel
    has run, which would test the conditioning of the loss rather than the kernels.
    """
    from test_gpu_models import build
    shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_stereodpnet.json").read_text()).items()}
    st = synth_state(shapes, seed=1)
    batch = synthetic_batch(2, 64, 96, training=True, seed=0)
    probe = ["aggregation.dres0.0.0.weight", "aggregation.dres4.conv6.0.weight", "normal_estimator.deform_conv1.weight",
             "normal_estimator.deform_conv2.conv_offset.weight", "normal_estimator.n_convs.0.0.weight",
             "cost_volume.attention_layer.mask_convs.0.weight"]
    so = dict(st)
    for k in probe:
        so[k] = st[k].clone().requires_grad_(True)
    lin = torch.randn(2, 3, 64, 96, generator=torch.Generator().manual_seed(9))
    want = O.stereodpnet_forward(dict(batch), so, True, predict_normal=True)
    (want["smoothL1_loss"] + 10.0 * (want["pred_normal"] * lin).mean()).backward()
    model = build("stereodpnet")
    model.load_state_dict(st, strict=False)
    model.cuda().train()
    model.encoder_autocast = False
    res = model({k: v.cuda() for k, v in batch.items()})
    (res["smoothL1_loss"] + 10.0 * (res["pred_normal"] * lin.cuda()).mean()).backward()
    torch.cuda.synchronize()
    d_err = (res["pred_depth"].detach().float().cpu() - want["pred_depth"].detach()).abs()
    n_err = (res["pred_normal"].detach().float().cpu() - want["pred_normal"].detach()).abs()
    print(f"SDP train (normals) depth max err {d_err.max():.4f}; normal max err {n_err.max():.4f} mean {n_err.mean():.5f}")
    assert d_err.max().item() < 0.15 and n_err.mean().item() < 4.5e-3           # measured on a B200: 0.071 px / 0.00215
    # the k sampled levels are a discrete function of the predicted disparity: where the two disparities straddle a
    # selection boundary the branch sees a different level set (inherent to bf16 vs fp32 forward, not to the backward kernels)
    crange = torch.tensor(model.normal_estimator.levels).view(1, -1, 1, 1)

    def level_set(d):
        dq = F.interpolate(d[:, 0].detach().float().cpu().unsqueeze(1), scale_factor=0.25, mode="nearest") * 0.25
        return O.anm_select_levels(dq, crange, model.normal_estimator.k).sort(1)[0]

    flipped = (level_set(res["pred_depth"]) != level_set(want["pred_depth"])).any(1).float().mean().item()
    print(f"   pixels whose sampled level set differs: {100 * flipped:.2f} %")
    for name in ("smoothL1_loss", "cosine_loss", "final_loss"):
        got, ref = float(res[name].detach()), float(want[name].detach())
        print(f"   {name}: {got:.5f} vs {ref:.5f}")
        assert abs(got - ref) < 1.5e-3 * abs(ref)                            # measured <= 5.2e-4 relative
    params = dict(model.named_parameters())
    for k in probe:
        got, ref = params[k].grad.float().cpu(), so[k].grad
        cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        print(f"   grad {k}: cosine {cos:.4f}, relative L2 error {rel2(got, ref):.4f}")
        # normal-branch gradients are ill-conditioned in the branch INPUT at random init: fp32 oracle vs fp32 oracle with out3
        # perturbed by 2 % already gives cosine 0.93 at a normal error of 1e-3 (tools/anm_grad_sensitivity.py); here out3 carries
        # the bf16 error of ~25 layers (normal error 2-3e-3).  test_anm_training_fwd_bwd holds the same kernels to 0.97 on
        # identical inputs; everything else keeps the depth-only floor
        assert cos > (0.75 if k.startswith("normal_estimator.") else 0.94)


def test_stereodpnet_shipped_loss_backward_runs():
    """final_loss (smooth-L1 + cosine) backward of the shipped config: every parameter on the path gets a finite gradient."""
    from test_gpu_models import build
    model = build("stereodpnet").cuda().train()
    res = model({k: v.cuda() for k, v in synthetic_batch(2, 64, 96, training=True).items()})
    res["final_loss"].backward()
    torch.cuda.synchronize()
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
