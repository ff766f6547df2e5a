"""Time the four strided hourglass layers (stride-2 conv1/conv3, transposed conv5/conv6) at the BASELINE config-2 shape."""
import argparse, sys
sys.path.insert(0, ".")
import torch
from dualpixelface_b200.layers import KIND_S2, KIND_T2, TCConv3d
ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=4); ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--only", type=str, default="")
a = ap.parse_args()
cases = [("conv1_s2_32to64", KIND_S2, 32, 64, (8, 280, 420)), ("conv3_s2_64to64", KIND_S2, 64, 64, (4, 140, 210)),
         ("conv5_t2_64to64", KIND_T2, 64, 64, (2, 70, 105)), ("conv6_t2_64to32", KIND_T2, 64, 32, (4, 140, 210))]
for name, kind, cin, cout, (d, h, w) in cases:
    if a.only and a.only not in name:
        continue
    x = torch.randn(a.b, d, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
    layer = TCConv3d(wt.transpose(0, 1).contiguous() if kind == KIND_T2 else wt, kind, transposed=kind == KIND_T2)
    sc, sh = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    for _ in range(3):
        y = layer(x, sc, sh, relu=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        y = layer(x, sc, sh, relu=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    vox_in = a.b * d * h * w
    fl = 2 * 27 * cin * cout * (vox_in if kind == KIND_T2 else vox_in // 8)
    print(f"{name}: {len(layer.plan)} launches, {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s, out {tuple(y.shape)}")
