"""GPU parity of the training path (forward with batch-statistics BatchNorm + backward) against PyTorch autograd / the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from dualpixelface_b200.synthetic import synthetic_batch

pytestmark = pytest.mark.gpu


def ndhwc(x):
    return x.permute(0, 2, 3, 4, 1).contiguous()


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def rel2(a, b):
    """Relative L2 error: robust to the few ReLU-mask flips that bf16 rounding of a near-zero activation causes (each flips
    one O(1) gradient element, which max-abs metrics over-weight)."""
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("kind,cin,cout,shape,residual,relu", [
    (0, 32, 32, (2, 4, 18, 20), True, True), (0, 64, 32, (1, 8, 16, 24), False, True), (0, 32, 32, (2, 3, 9, 11), True, False),
    (1, 32, 64, (2, 8, 20, 24), False, True), (1, 64, 64, (1, 4, 12, 16), False, True),
    (2, 64, 64, (1, 2, 9, 13), True, True), (2, 64, 32, (2, 4, 10, 12), True, False)])
def test_conv_bn_act_fwd_bwd(kind, cin, cout, shape, residual, relu):
    from dualpixelface_b200.train_ops import ConvBNAct, LayerCfg
    g = torch.Generator().manual_seed(100 + kind)
    b, d, h, w = shape
    x = torch.randn(b, cin, d, h, w, generator=g).to(torch.bfloat16)
    wshape = (cin, cout, 3, 3, 3) if kind == 2 else (cout, cin, 3, 3, 3)
    wt = (torch.randn(*wshape, generator=g) * (2.0 / (27 * cin)) ** 0.5)
    gamma, beta = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.2
    # fp32 reference on the bf16-rounded input / weight
    xr = x.float().requires_grad_(True)
    wr = wt.to(torch.bfloat16).float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    if kind == 0:
        z = F.conv3d(xr, wr, padding=1)
    elif kind == 1:
        z = F.conv3d(xr, wr, stride=2, padding=1)
    else:
        z = F.conv_transpose3d(xr, wr, stride=2, padding=1, output_padding=1)
    res = torch.randn(z.shape, generator=g).to(torch.bfloat16) if residual else None
    rr = res.float().requires_grad_(True) if residual else None
    y = F.batch_norm(z, None, None, gr, br, True, 0.1, 1e-5)
    if residual:
        y = y + rr
    if relu:
        y = F.relu(y)
    dy = torch.randn(y.shape, generator=g).to(torch.bfloat16)
    y.backward(dy.float())
    # sm_100a path
    bn = torch.nn.BatchNorm3d(cout).cuda()
    xg = ndhwc(x).cuda().requires_grad_(True)
    wg = wt.cuda().requires_grad_(True)
    gg, bg = gamma.cuda().requires_grad_(True), beta.cuda().requires_grad_(True)
    rg = ndhwc(res).cuda().requires_grad_(True) if residual else None
    out = ConvBNAct.apply(xg, wg, gg, bg, rg, LayerCfg(kind, relu, bn))
    out.backward(ndhwc(dy).cuda())
    torch.cuda.synchronize()
    assert rel(out.permute(0, 4, 1, 2, 3), y) < 2e-2
    assert rel2(xg.grad.permute(0, 4, 1, 2, 3), xr.grad) < 3e-2
    assert rel2(wg.grad, wr.grad) < 3e-2
    assert rel2(gg.grad, gr.grad) < 3e-2 and rel2(bg.grad, br.grad) < 3e-2
    if residual:
        assert rel2(rg.grad.permute(0, 4, 1, 2, 3), rr.grad) < 3e-2
    assert bn.num_batches_tracked.item() == 1 and bn.running_mean.abs().sum().item() > 0


def test_head_conv_fwd_bwd():
    from dualpixelface_b200.train_ops import HeadConv
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 32, 8, 12, 14, generator=g).to(torch.bfloat16)
    wt = torch.randn(1, 32, 3, 3, 3, generator=g) * 0.05
    prev = torch.randn(2, 1, 8, 12, 14, generator=g)
    xr, wr, pr = x.float().requires_grad_(True), wt.to(torch.bfloat16).float().requires_grad_(True), prev.clone().requires_grad_(True)
    y = F.conv3d(xr, wr, padding=1) + pr
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xg, wg = ndhwc(x).cuda().requires_grad_(True), wt.cuda().requires_grad_(True)
    pg = ndhwc(prev).cuda().requires_grad_(True)
    out = HeadConv.apply(xg, wg, pg)
    out.backward(ndhwc(dy).cuda())
    assert rel(out.permute(0, 4, 1, 2, 3), y) < 2e-2
    assert rel2(xg.grad.permute(0, 4, 1, 2, 3), xr.grad) < 3e-2
    assert rel2(wg.grad, wr.grad) < 3e-2
    assert rel(pg.grad.permute(0, 4, 1, 2, 3), pr.grad) < 1e-6


def test_psmnet_training_step_matches_reference():
    """One fwd+bwd of PSMNET on the sm_100a path vs the golden fixture produced by the unmodified reference."""
    from test_gpu_models import build
    import json
    from dualpixelface_b200.synthetic import synth_state
    gold = np.load(GOLDEN / "model_psmnet.npz")
    shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_psmnet.json").read_text()).items()}
    model = build("psmnet")
    model.load_state_dict(synth_state(shapes, seed=1), strict=False)
    model.cuda().train()
    model.encoder_autocast = False
    batch = {k: v.cuda() for k, v in synthetic_batch(2, 256, 256, training=True, seed=0).items()}
    res = model(batch)
    res["final_loss"].backward()
    torch.cuda.synchronize()
    d_err = (res["pred_depth"].detach().float().cpu() - torch.as_tensor(gold["train/pred_depth"])).abs()
    print(f"train pred_depth max err {d_err.max():.4f} mean {d_err.mean():.5f}; loss {float(res['final_loss'].detach()):.5f} vs {float(gold['train/final_loss']):.5f}")
    assert res["pred_depth"].shape[1] == 3
    assert d_err.max().item() < 0.17 and d_err.mean().item() < 0.015          # measured on a B200: 0.085 / 0.0073 px
    assert abs(float(res["final_loss"].detach()) - float(gold["train/final_loss"])) < 1.5e-3 * float(gold["train/final_loss"])   # measured 4.3e-4
    params = dict(model.named_parameters())
    # bf16 activations AND bf16 activation-gradients through 28 conv+BN layers: the error grows with depth from the loss
    # (measured cosine 1.0000 / 0.9999 / 0.994 / 0.958); thresholds are per depth.
    floor = {"aggregation.classif3.2.weight": 0.999, "aggregation.dres4.conv6.0.weight": 0.999,
             "aggregation.dres0.0.0.weight": 0.99, "feature_extraction.firstconv.0.0.weight": 0.94}
    for key in floor:
        want = torch.as_tensor(gold[f"train/grad/{key}"])
        got = params[key].grad.float().cpu()
        cos = F.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
        r = ((got - want).norm() / want.norm()).item()
        print(f"   grad {key}: cosine {cos:.4f}, relative L2 error {r:.4f}")
        assert cos > floor[key] and r < 0.35


@pytest.mark.parametrize("c,hw", [(32, (37, 50)), (64, (20, 28)), (128, (9, 13))])
def test_encoder_batchnorm_train_matches_torch(c, hw):
    """modules.EncoderBatchNorm2d (train mode, bf16 channels-last: the repository's BatchNorm kernels) == nn.BatchNorm2d in fp32 on
    the same bf16 input: output, running statistics, input / weight / bias gradients."""
    from dualpixelface_b200.modules import EncoderBatchNorm2d
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.randn(3, c, *hw, device="cuda", generator=g) * 1.5 + 0.3).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(3, c, *hw, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ours, ref = EncoderBatchNorm2d(c).cuda().train(), torch.nn.BatchNorm2d(c).cuda().train()
    ours.min_pixels = 0                                   # the size threshold is a performance heuristic; test the kernels
    with torch.no_grad():
        for m in (ours, ref):
            m.weight.copy_(torch.linspace(0.5, 1.5, c)); m.bias.copy_(torch.linspace(-0.2, 0.2, c))
    xo = x.clone().requires_grad_(True)
    yo = ours(xo)
    assert yo.dtype == torch.bfloat16 and yo.is_contiguous(memory_format=torch.channels_last)
    yo.backward(dy)
    xr = x.float().requires_grad_(True)
    yr = ref(xr)
    yr.backward(dy.float())
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-9))
    print(f"BN2d C={c}: y {rel(yo, yr):.2e}, dx {rel(xo.grad, xr.grad):.2e}, dgamma {rel(ours.weight.grad, ref.weight.grad):.2e}, "
          f"dbeta {rel(ours.bias.grad, ref.bias.grad):.2e}, running var {rel(ours.running_var, ref.running_var):.2e}")
    assert rel(yo, yr) < 5e-3 and rel(xo.grad, xr.grad) < 8e-3                       # bf16 rounding of y / dx
    assert rel(ours.weight.grad, ref.weight.grad) < 2e-3 and rel(ours.bias.grad, ref.bias.grad) < 2e-3
    assert rel(ours.running_mean, ref.running_mean) < 1e-4 and rel(ours.running_var, ref.running_var) < 1e-4
    assert int(ours.num_batches_tracked) == 1
