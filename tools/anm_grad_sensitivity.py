"""Oracle-vs-oracle (fp32, CPU): how much the ANM parameter gradients move when the branch input out3 is perturbed by eps relative
noise.  Output on this container: eps 0.002 -> cosine 0.993, 0.01 -> 0.969, 0.02 -> 0.933 (normal mean err 0.0001/0.0005/0.0010):
the justification of the cosine floors in tests/test_gpu_training_sdp.py."""
import json, sys, torch, torch.nn.functional as F
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import GOLDEN
from dualpixelface_b200.synthetic import synth_state, synthetic_batch
from oracle import dpf_oracle as O
shapes = {k: tuple(v) for k, v in json.loads((GOLDEN / "state_keys_stereodpnet.json").read_text()).items()}
st = synth_state(shapes, seed=1)
batch = synthetic_batch(2, 64, 96, training=True, seed=0)
g = torch.Generator().manual_seed(77)
out3 = torch.relu(torch.randn(2, 32, 8, 16, 24, generator=g))
disp = torch.rand(2, 64, 96, generator=g) * 14.0 - 3.0
dn = torch.randn(2, 3, 64, 96, generator=g)
keys = ["normal_estimator.deform_conv1.weight", "normal_estimator.deform_conv2.conv_offset.weight", "normal_estimator.n_convs.0.0.weight"]
lv = [float(v) for v in O.cost_range(-4, 12, 8)]
def run(x):
    so = dict(st)
    for k in keys: so[k] = st[k].clone().requires_grad_(True)
    o3 = x.clone().requires_grad_(True)
    n = O.anm_forward(o3, disp, batch["K"], batch["abvalue"], so, "normal_estimator", lv, True, 4)
    n.backward(dn)
    return n.detach(), [so[k].grad for k in keys] + [o3.grad]
n0, g0 = run(out3)
for eps in (0.002, 0.01, 0.02):
    n1, g1 = run(out3 * (1 + eps * torch.randn(out3.shape, generator=g)))
    print(eps, "normal mean err %.5f" % (n1 - n0).abs().mean().item(), ["%.4f" % F.cosine_similarity(a.flatten(), b.flatten(), dim=0).item() for a, b in zip(g0, g1)])
