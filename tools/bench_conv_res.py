"""kd-fused 3x3x3 conv with / without a bf16 residual at the config-2 shape (dres1.2 of the aggregation)."""
import sys
sys.path.insert(0, ".")
import torch
from dualpixelface_b200.layers import KIND_3x3x3, TCConv3d
b, d, h, w = 4, 8, 280, 420
for cin in (32, 64):
    x = torch.randn(b, d, h, w, cin, device="cuda").to(torch.bfloat16)
    res = torch.randn(b, d, h, w, 32, device="cuda").to(torch.bfloat16)
    layer = TCConv3d(torch.randn(32, cin, 3, 3, 3, device="cuda") * 0.05, KIND_3x3x3)
    sc, sh = torch.ones(32, device="cuda"), torch.zeros(32, device="cuda")
    for name, r in (("no residual", None), ("residual", res)):
        for _ in range(3): layer(x, sc, sh, residual=r, relu=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): layer(x, sc, sh, residual=r, relu=True)
        e1.record(); torch.cuda.synchronize()
        print(f"{cin}->32 {name}: {e0.elapsed_time(e1) / 10:.4f} ms")
