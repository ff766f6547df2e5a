"""GPU-side diagnostics for the tcgen05 convolution (run under gpurun; prints, asserts nothing)."""
import sys, time
sys.path.insert(0, ".")
import torch, torch.nn.functional as F
from dualpixelface_b200 import ops

def run(cin, cout, kind, shape, seed=0, delta_tap=None):
    b, d, h, w = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, cin, d, h, w, generator=g).to(torch.bfloat16)
    ks = {0: (3, 3, 3), 3: (1, 3, 3), 4: (1, 1, 1)}[kind]
    wt = torch.randn(cout, cin, *ks, generator=g) * 0.05
    if delta_tap is not None:
        wt = torch.zeros(cout, cin, *ks)
        for c in range(min(cin, cout)):
            wt[(c,c) + delta_tap] = 1.0
    wt = wt.to(torch.bfloat16)
    want = F.conv3d(x.float(), wt.float(), padding=tuple(k // 2 for k in ks))
    got = ops.conv3d(x.permute(0, 2, 3, 4, 1).contiguous().cuda(), ops.pack_conv_weight(wt.cuda()), kind, cout)
    torch.cuda.synchronize()
    got = got.permute(0, 4, 1, 2, 3).float().cpu()
    err = (got - want).abs()
    print(f"cin={cin} cout={cout} kind={kind} shape={shape} delta={delta_tap}: max err {err.max():.4f} (ref max {want.abs().max():.3f}) "
          f"frac bad {(err > 0.05 * want.abs().max()).float().mean():.4f}", flush=True)
    if err.max() > 0.05 * want.abs().max():
        bad = (err > 0.05 * want.abs().max())
        print("   bad per channel :", bad.float().mean(dim=(0, 2, 3, 4))[:8].tolist())
        print("   bad per d       :", bad.float().mean(dim=(0, 1, 3, 4)).tolist())
        print("   bad per h (0:20):", [round(v, 2) for v in bad.float().mean(dim=(0, 1, 2, 4))[:20].tolist()])
        print("   bad per w (0:30):", [round(v, 2) for v in bad.float().mean(dim=(0, 1, 2, 3))[:30].tolist()])
        print("   got[0,0,0,:3,:6]", got[0, 0, 0, :3, :6].tolist())
        print("   want[0,0,0,:3,:6]", want[0, 0, 0, :3, :6].tolist())

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    ops.lib()
    run(32, 32, 4, (1, 1, 16, 8))
    run(32, 32, 4, (1, 2, 16, 24))
    run(32, 32, 0, (1, 1, 16, 8), delta_tap=(1, 1, 1))
    run(32, 32, 0, (1, 1, 16, 8), delta_tap=(1, 0, 1))
    run(32, 32, 0, (1, 1, 16, 8), delta_tap=(1, 1, 2))
    run(32, 32, 0, (1, 3, 16, 24), delta_tap=(0, 1, 1))
    run(32, 32, 0, (1, 4, 16, 24))
    run(64, 32, 0, (1, 4, 16, 24))
    run(32, 64, 0, (1, 4, 16, 24))
    run(32, 16, 0, (1, 4, 16, 24))
    run(32, 32, 0, (2, 8, 70, 105))
    # quick timing at the BASELINE c2 shape, one sample
    for cin, cout in ((32, 32), (64, 32)):
        x = torch.randn(1, 8, 280, 420, cin, device="cuda").to(torch.bfloat16)
        wp = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05)
        for _ in range(3):
            y = ops.conv3d(x, wp, 0, cout, relu=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = ops.conv3d(x, wp, 0, cout, relu=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2 * 8 * 280 * 420 * 27 * cin * cout
        print(f"conv {cin}->{cout} 8x280x420: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
