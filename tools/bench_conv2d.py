"""Time the ANM n_convs stack (six dilated 3x3 convs + LeakyReLU) on dpf_conv2d_tc_fwd at the config-2 shape (16 x 280 x 420)
against cuDNN (bf16 channels-last F.conv2d + leaky_relu), layer by layer."""
import sys
sys.path.insert(0, ".")
import torch
import torch.nn.functional as F
from dualpixelface_b200 import ops

def timed(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

N, H, W = 16, 280, 420
tot_a = tot_b = 0.0
for cin, cout, dil in ((64, 96, 1), (96, 96, 2), (96, 64, 4), (64, 64, 8), (64, 32, 1), (32, 3, 1)):
    x = torch.randn(N, H, W, cin, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(cout, cin, 3, 3, device="cuda") * 0.05)
    wp = ops.pack_conv2d_tc_weight(wt)
    a = timed(lambda: ops.conv2d_tc(x, wp, cout, dil, relu=True, slope=0.1))
    xc = x.permute(0, 3, 1, 2)
    wc = wt.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = timed(lambda: F.leaky_relu(F.conv2d(xc, wc, None, 1, dil, dil), 0.1))
    fl = 2 * 9 * cin * cout * N * H * W
    tot_a += a; tot_b += b
    print(f"conv2d {cin:3d}->{cout:3d} d{dil}: tcgen05 {a:.3f} ms ({fl / a / 1e9:7.1f} TFLOP/s)   cuDNN+leaky {b:.3f} ms ({fl / b / 1e9:7.1f} TFLOP/s)")
print(f"n_convs stack: tcgen05 {tot_a:.3f} ms, cuDNN {tot_b:.3f} ms")
