import sys; sys.path.insert(0, '.')
import torch, torch.nn.functional as F
from dualpixelface_b200.ops_wgrad import conv3d_wgrad
def case(kind, cin, cout, shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    b, d, h, w = shape
    x = torch.randn(b, cin, d, h, w, generator=g).to(torch.bfloat16)
    dz = torch.randn(b, cout, d, h, w, generator=g).to(torch.bfloat16)
    ks = {0: (3,3,3), 3: (1,3,3), 4: (1,1,1)}[kind]
    wt = torch.zeros(cout, cin, *ks, requires_grad=True)
    y = F.conv3d(x.float(), wt, padding=tuple(k//2 for k in ks))
    y.backward(dz.float())
    got = conv3d_wgrad(x.permute(0,2,3,4,1).contiguous().cuda(), dz.permute(0,2,3,4,1).contiguous().cuda(), kind).cpu()
    err = ((got - wt.grad).norm() / wt.grad.norm()).item()
    print(f"kind={kind} {cin}->{cout} {shape}: rel L2 err {err:.5f}  max|ref| {wt.grad.abs().max():.3f}", flush=True)
    if err > 0.02:
        print("   got[0,0]", got[0,0].flatten()[:6].tolist(), "\n   ref[0,0]", wt.grad[0,0].flatten()[:6].tolist())
        print("   per-tap err", [round(((got[:,:,i//9 if kind==0 else 0,(i//3)%3 if kind!=4 else 0,i%3 if kind!=4 else 0]-wt.grad[:,:,i//9 if kind==0 else 0,(i//3)%3 if kind!=4 else 0,i%3 if kind!=4 else 0]).norm()/wt.grad.norm()).item(),3) for i in range(min(27, ks[0]*ks[1]*ks[2]))])
case(4, 64, 32, (1, 1, 16, 16))
case(4, 32, 32, (1, 1, 16, 16))
case(4, 32, 32, (2, 3, 20, 37))
case(3, 32, 32, (1, 2, 16, 16))
case(0, 32, 32, (1, 4, 16, 16))
case(0, 32, 32, (2, 8, 37, 53))
case(0, 64, 32, (1, 4, 18, 26))
case(0, 32, 16, (1, 4, 18, 26))
case(0, 32, 1, (2, 8, 20, 30))
case(0, 64, 64, (1, 4, 18, 26))
# timing at the BASELINE c2 shape
import torch
from dualpixelface_b200 import ops
for cin, cout in ((32, 32), (64, 32)):
    x = torch.randn(4, 8, 280, 420, cin, device="cuda").to(torch.bfloat16)
    dz = torch.randn(4, 8, 280, 420, cout, device="cuda").to(torch.bfloat16)
    for _ in range(2): conv3d_wgrad(x, dz, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): conv3d_wgrad(x, dz, 0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"wgrad {cin}->{cout} 4x8x280x420: {ms:.3f} ms  {2*4*8*280*420*27*cin*cout/ms/1e9:.1f} TFLOP/s")
    w = torch.zeros(cout, cin, 3, 3, 3, device="cuda", dtype=torch.bfloat16)
    xn, zn = x.permute(0, 4, 1, 2, 3), dz.permute(0, 4, 1, 2, 3)
    f = lambda: torch.ops.aten.convolution_backward(zn, xn, w, None, [1,1,1], [1,1,1], [1,1,1], False, [0,0,0], 1, [False, True, False])
    for _ in range(2): f()
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    print(f"   cuDNN wgrad (aten.convolution_backward): {e0.elapsed_time(e1)/5:.3f} ms")
