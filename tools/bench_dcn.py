"""Time the D3D kernel alone at the bench shape (ANM: B=4, K=4 planes, 280x420)."""
import argparse, sys
sys.path.insert(0, ".")
import torch
from dualpixelface_b200 import ops
ap = argparse.ArgumentParser()
ap.add_argument("--cin", type=int, default=64); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--b", type=int, default=4); ap.add_argument("--offscale", type=float, default=0.5)
ap.add_argument("--offc", type=int, default=96, help="offset channel pitch: 96 = staged path (the model), 81 = direct loads")
a = ap.parse_args()
b, d, h, w = a.b, 4, 280, 420
cpad = 64
x = torch.randn(b, d, h, w, 64, device="cuda").to(torch.bfloat16)
off = (torch.randn(b, d, h, w, a.offc, device="cuda") * a.offscale).contiguous()
wp = ops.pack_conv_weight(torch.randn(64, a.cin, 3, 3, 3, device="cuda") * 0.05, cin_pad=cpad)
for _ in range(2):
    y = ops.dcn3d(x, off, wp, cpad, relu=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    y = ops.dcn3d(x, off, wp, cpad, relu=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
fl = 2 * b * d * h * w * 27 * cpad * 64
print(f"dcn3d offc={a.offc} cin_pad={cpad} {b}x{d}x{h}x{w}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (gather L1 traffic {b*d*h*w*27*8*cpad*2/ms/1e9:.1f} TB/s)")
