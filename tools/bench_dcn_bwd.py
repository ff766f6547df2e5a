"""Time the D3D backward kernels at the training shape (B pairs x k=4 planes x 280 x 420, 64 channels)."""
import argparse, sys
sys.path.insert(0, ".")
import torch
from dualpixelface_b200 import ops
from dualpixelface_b200.ops_dcn_bwd import dcn3d_bwd_data, dcn3d_bwd_weight
ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=2); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(a.b, 4, 280, 420, 64, device="cuda", generator=g).to(torch.bfloat16)
off = (torch.rand(a.b, 4, 280, 420, 81, device="cuda", generator=g) - 0.5) * 2.0
dy = torch.randn(a.b, 4, 280, 420, 64, device="cuda", generator=g).to(torch.bfloat16)
w = torch.randn(64, 64, 3, 3, 3, device="cuda", generator=g) * 0.03
def timed(f):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters
vox = a.b * 4 * 280 * 420
fl = 2 * 27 * 64 * 64 * vox
wp = ops.pack_conv_weight(w, cin_pad=64)
t = timed(lambda: ops.dcn3d(x, off, wp, 64)); print(f"dcn3d fwd        {t:7.3f} ms  {fl / t / 1e9:6.1f} TFLOP/s")
t = timed(lambda: dcn3d_bwd_data(x, off, dy, w, 64)); print(f"dcn3d bwd data64 {t:7.3f} ms  (includes the dx memset)")
t = timed(lambda: dcn3d_bwd_data(x, off, dy, w, 32)); print(f"dcn3d bwd data32 {t:7.3f} ms")
t = timed(lambda: dcn3d_bwd_weight(x, off, dy, 64)); print(f"dcn3d bwd weight {t:7.3f} ms  {fl / t / 1e9:6.1f} TFLOP/s")
