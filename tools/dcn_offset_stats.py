"""Offset statistics of the two D3D layers inside the bench model (what the gather's locality depends on)."""
import sys
sys.path.insert(0, ".")
import torch
import bench
from dualpixelface_b200 import ops
from dualpixelface_b200.synthetic import synthetic_batch
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
batch = {k: v.to(dev) for k, v in synthetic_batch(2, 1120, 1680, seed=0).items()}
seen = []
orig = ops.dcn3d
def spy(x, off, *a, **k):
    seen.append(off)
    return orig(x, off, *a, **k)
ops.dcn3d = spy
import dualpixelface_b200.modules as M
with torch.no_grad():
    model(batch)
for i, off in enumerate(seen):
    o = off[..., :81].float()
    q = torch.quantile(o.abs().flatten()[:: 97], torch.tensor([0.5, 0.9, 0.99], device=dev))
    print(f"D3D layer {i + 1}: offsets shape {tuple(off.shape)} mean {o.mean():.3f} std {o.std():.3f} |o| median {q[0]:.3f} p90 {q[1]:.3f} p99 {q[2]:.3f} max {o.abs().max():.2f}")
    od = o.view(*o.shape[:-1], 27, 3)
    print("   per-axis std (d,h,w):", [round(float(od[..., a].std()), 3) for a in range(3)], " per-axis mean:", [round(float(od[..., a].mean()), 3) for a in range(3)])
