"""HBM-bound kernels at the BASELINE c2 shape: achieved GB/s on the algorithmic bytes."""
import sys
sys.path.insert(0, ".")
import torch
from dualpixelface_b200 import ops
from dualpixelface_b200.shift_tables import build_tables
B, H4, W4, C, D = 4, 280, 420, 32, 8
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
ref = torch.randn(B, H4, W4, C, device="cuda").to(torch.bfloat16); tgt = torch.randn_like(ref)
shifts = [-1, 0, 0, 0, 1, 1, 2, 2]
for mode, g, cv in (("concat", 0, 64), ("diff", 0, 32), ("gwc", 8, 8)):
    ms = timeit(lambda: ops.costvol_fwd(ref, tgt, shifts, mode, g))
    by = B * H4 * W4 * (2 * C * 2 + D * cv * 2)
    print(f"costvol_fwd {mode:6s}: {ms:.4f} ms  {by / ms / 1e6:8.1f} GB/s  ({by / 1e6:.1f} MB algorithmic)")
dvol = torch.randn(B, D, H4, W4, 64, device="cuda").to(torch.bfloat16)
ms = timeit(lambda: ops.costvol_bwd(ref, tgt, dvol, shifts, "concat"))
by = B * H4 * W4 * (D * 64 * 2 + 2 * C * 2)
print(f"costvol_bwd concat: {ms:.4f} ms  {by / ms / 1e6:8.1f} GB/s")
cost = torch.randn(B, D, H4, W4, device="cuda")
ms = timeit(lambda: ops.regress_fwd(cost, -4.0, 0.5))
by = B * (D * H4 * W4 * 4 + 16 * H4 * W4 * 4)
print(f"regress_fwd       : {ms:.4f} ms  {by / ms / 1e6:8.1f} GB/s  ({by / 1e6:.1f} MB algorithmic)")
dd = torch.randn(B, 4 * H4, 4 * W4, device="cuda")
ms = timeit(lambda: ops.regress_bwd(cost, dd, -4.0, 0.5))
print(f"regress_bwd       : {ms:.4f} ms  {by / ms / 1e6:8.1f} GB/s")
tb = {k: v.cuda() for k, v in build_tables(H4, W4, -1.0, "forward").items()}
ms = timeit(lambda: ops.asm_sample(ref, tb))
by = B * H4 * W4 * (C * 2 + 3 * C * 2)
print(f"asm_sample        : {ms:.4f} ms  {by / ms / 1e6:8.1f} GB/s")
smp = torch.randn(B, 3, H4, W4, C, device="cuda").to(torch.bfloat16); lg = torch.randn_like(smp)
a = torch.rand(B, C, device="cuda"); d = torch.rand(B, C, device="cuda")
vol = torch.empty(B, D, H4, W4, 2 * C, device="cuda", dtype=torch.bfloat16)
ms = timeit(lambda: ops.asm_blend(smp, lg, a, d, vol, 0, D, 0))
by = B * H4 * W4 * (2 * 3 * C * 2 + D * C * 2)
print(f"asm_blend (8 lvl) : {ms:.4f} ms  {by / ms / 1e6:8.1f} GB/s  ({by / 1e6:.1f} MB algorithmic)")
ms = timeit(lambda: ops.channel_stats(lg))
print(f"channel_stats     : {ms:.4f} ms  {lg.numel() * 2 / ms / 1e6:8.1f} GB/s")
