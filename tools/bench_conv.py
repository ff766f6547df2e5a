"""Time one tcgen05 conv layer at a BASELINE shape (for ncu captures and quick A/B runs)."""
import argparse, sys
sys.path.insert(0, ".")
import torch
from dualpixelface_b200 import ops
ap = argparse.ArgumentParser()
ap.add_argument("--cin", type=int, default=32); ap.add_argument("--cout", type=int, default=32)
ap.add_argument("--b", type=int, default=1); ap.add_argument("--d", type=int, default=8)
ap.add_argument("--h", type=int, default=280); ap.add_argument("--w", type=int, default=420)
ap.add_argument("--iters", type=int, default=10); ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--kind", type=int, default=0)
a = ap.parse_args()
x = torch.randn(a.b, a.d, a.h, a.w, a.cin, device="cuda").to(torch.bfloat16)
ks = {0: (3, 3, 3), 3: (1, 3, 3), 4: (1, 1, 1)}[a.kind]
wp = ops.pack_conv_weight(torch.randn(a.cout, a.cin, *ks, device="cuda") * 0.05)
for _ in range(a.warmup):
    y = ops.conv3d(x, wp, a.kind, a.cout, relu=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    y = ops.conv3d(x, wp, a.kind, a.cout, relu=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
fl = 2 * a.b * a.d * a.h * a.w * ks[0] * ks[1] * ks[2] * a.cin * a.cout
print(f"conv kind={a.kind} {a.cin}->{a.cout} {a.b}x{a.d}x{a.h}x{a.w}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
