"""Time the tcgen05 weight-gradient kernel at the aggregation shapes."""
import sys
sys.path.insert(0, ".")
import torch
from dualpixelface_b200.ops_wgrad import conv3d_wgrad
for kind, cin, cout, shape in ((0, 32, 32, (2, 8, 280, 420)), (0, 64, 32, (2, 8, 280, 420)), (1, 32, 64, (2, 8, 280, 420)), (1, 64, 64, (2, 4, 140, 210))):
    b, d, h, w = shape
    x = torch.randn(b, d, h, w, cin, device="cuda").to(torch.bfloat16)
    zs = (b, (d + 1) // 2, (h + 1) // 2, (w + 1) // 2) if kind == 1 else shape
    dz = torch.randn(*zs, cout, device="cuda").to(torch.bfloat16)
    for _ in range(2):
        conv3d_wgrad(x, dz, kind)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        conv3d_wgrad(x, dz, kind)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2 * 27 * cin * cout * zs[0] * zs[1] * zs[2] * zs[3]
    print(f"wgrad kind={kind} {cin}->{cout} {shape}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
