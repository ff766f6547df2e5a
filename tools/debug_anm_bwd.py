"""Stage-wise gradient comparison of the ANM training path against the oracle (debug aid)."""
import json, sys
import torch, torch.nn.functional as F
sys.path.insert(0, "tests")
from conftest import GOLDEN
from test_gpu_models import build
from test_gpu_training_sdp import ANM_PROBE, rel2
from dualpixelface_b200.synthetic import synth_state, synthetic_batch
from oracle import dpf_oracle as O

shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_stereodpnet.json").read_text()).items()}
st = synth_state(shapes, seed=1)
batch = synthetic_batch(2, 64, 96, training=True, seed=0)
g = torch.Generator().manual_seed(77)
out3 = torch.relu(torch.randn(2, 32, 8, 16, 24, generator=g)).to(torch.bfloat16)
disp = torch.rand(2, 64, 96, generator=g) * 14.0 - 3.0
model = build("stereodpnet"); model.load_state_dict(st, strict=False)
anm = model.normal_estimator.cuda().train()
so = dict(st)
for k in ANM_PROBE:
    so[k] = st[k].clone().requires_grad_(True)
o3 = out3.float().requires_grad_(True)
want, aux = O.anm_forward(o3, disp, batch["K"], batch["abvalue"], so, "normal_estimator", anm.levels, True, anm.k, return_aux=True)
for k in ("fv", "off1", "off2", "f1", "f2"):
    aux[k].retain_grad()
dn = torch.randn(want.shape, generator=g)
want.backward(dn)

import dualpixelface_b200.train_anm as TA
saved = {}
orig = TA.anm_train
def hooked(anm_, o, d, b):
    # re-implementation with retained intermediate grads
    from dualpixelface_b200 import ops
    bsz = o.shape[0]
    kq = b["K"].float().clone(); kq[:, :2, :] = kq[:, :2, :] / 4.0
    kinv = torch.inverse(kq).contiguous()
    idx, coord, minmax = ops.anm_select(d.detach().contiguous(), kinv, b["abvalue"].float().contiguous(), anm_.levels, anm_.k)
    fv = TA.GatherFn.apply(o, idx, coord, minmax); fv.retain_grad(); saved["fv"] = fv
    x = fv
    for i, (dc, act) in enumerate(((anm_.deform_conv1, anm_.act1), (anm_.deform_conv2, anm_.act2)), 1):
        off = TA.OffsetConvFn.apply(x, dc.conv_offset.weight, dc.conv_offset.bias); off.retain_grad(); saved[f"off{i}"] = off
        z = TA.DCNFn.apply(x, off, dc.weight)
        x = TA.BNActFn.apply(z, act[0].weight, act[0].bias, dc.bias, act[0]); x.retain_grad(); saved[f"f{i}"] = x
    f = x.view(bsz * anm_.k, x.shape[2], x.shape[3], x.shape[4]).permute(0, 3, 1, 2)
    for m in anm_.n_convs:
        conv = m[0]
        f = F.leaky_relu(F.conv2d(f, conv.weight.to(torch.bfloat16), None, 1, conv.dilation, conv.dilation), 0.1)
    return TA.TailFn.apply(f.permute(0, 2, 3, 1), bsz, anm_.k), saved["off1"], saved["off2"]
og = out3.permute(0, 2, 3, 4, 1).contiguous().cuda().requires_grad_(True)
normal, _, _ = hooked(anm, og, disp.cuda(), {k: v.cuda() for k, v in batch.items()})
normal.backward(dn.cuda())
torch.cuda.synchronize()
def cmp(name, got, ref):
    got, ref = got.float().cpu(), ref.float()
    cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    print(f"{name:28s} cos {cos:.4f} relL2 {rel2(got, ref):.4f}  |ref| {ref.norm():.4g}")
print("normal max err", (normal.detach().cpu() - want.detach()).abs().max().item())
for k in ("f2", "off2", "f1", "off1", "fv"):
    ref = aux[k]                                  # [b c k h w]
    got = saved[k]
    c = ref.shape[1]
    cmp("fwd " + k, got.detach()[..., :c].permute(0, 4, 1, 2, 3), ref.detach())
    cmp("grad " + k, got.grad[..., :c].permute(0, 4, 1, 2, 3), ref.grad)
cmp("grad out3", og.grad.permute(0, 4, 1, 2, 3), o3.grad)
params = dict(model.named_parameters())
for k in ANM_PROBE:
    cmp(k.replace("normal_estimator.", ""), params[k].grad, so[k].grad)
