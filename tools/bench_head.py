"""Time the 32 -> 1 head conv at the BASELINE config-2 shape: dedicated kernel vs the generic engine (N = 16 padded MMAs)."""
import sys
sys.path.insert(0, ".")
import torch
from dualpixelface_b200 import ops
from dualpixelface_b200.layers import KIND_3x3x3, TCConv3d
b, d, h, w = 4, 8, 280, 420
x = torch.randn(b, d, h, w, 32, device="cuda").to(torch.bfloat16)
wt = torch.randn(1, 32, 3, 3, 3, device="cuda") * 0.05
res = torch.randn(b, d, h, w, 1, device="cuda")
layer = TCConv3d(wt, KIND_3x3x3)
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
by = b * d * h * w * (64 + 4 + 4)
ms = timeit(lambda: layer(x, residual=res, out_f32=True))
print(f"head kernel : {ms:.4f} ms  {by / ms / 1e6:.0f} GB/s algorithmic ({by / 1e6:.0f} MB)")
head, layer.head = layer.head, None
ms = timeit(lambda: layer(x, residual=res, out_f32=True))
print(f"generic path: {ms:.4f} ms")
