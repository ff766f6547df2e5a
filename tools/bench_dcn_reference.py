"""Reference D3D CUDA kernel (oracle/_ref/DCN.so, fp32 im2col + GEMM) vs dpf_dcn3d_fwd at the bench shape."""
import importlib.machinery, importlib.util, sys
sys.path.insert(0, ".")
import torch
from dualpixelface_b200 import ops
loader = importlib.machinery.ExtensionFileLoader("DCN", "oracle/_ref/DCN.so")
DCN = importlib.util.module_from_spec(importlib.util.spec_from_loader("DCN", loader)); loader.exec_module(DCN)
b, d, h, w = 4, 4, 280, 420
x = torch.randn(b, 64, d, h, w, device="cuda")
off = torch.randn(b, 81, d, h, w, device="cuda") * 0.5
wt = torch.randn(64, 64, 3, 3, 3, device="cuda") * 0.05
bias = torch.zeros(64, device="cuda")
args = (3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1)
def timed(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_ref = timed(lambda: DCN.deform_conv_forward(x, wt, bias, off, *args))
xp = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16).contiguous()
offp = off.permute(0, 2, 3, 4, 1).contiguous()
wp = ops.pack_conv_weight(wt, cin_pad=64)
t_new = timed(lambda: ops.dcn3d(xp, offp, wp, 64))
print(f"D3D forward {b}x64x{d}x{h}x{w}: reference CUDA kernels (fp32, im2col_step=1) {t_ref:.2f} ms, dpf_dcn3d_fwd (bf16) {t_new:.2f} ms, x{t_ref / t_new:.1f}")
dy = torch.randn(b, 64, d, h, w, device="cuda")
t_refb = timed(lambda: DCN.deform_conv_backward(x, wt, bias, off, dy, *args), 2)
from dualpixelface_b200.ops_dcn_bwd import dcn3d_bwd_data, dcn3d_bwd_weight
dyp = dy.permute(0, 2, 3, 4, 1).to(torch.bfloat16).contiguous()
t_newb = timed(lambda: (dcn3d_bwd_data(xp, offp, dyp, wt), dcn3d_bwd_weight(xp, offp, dyp, 64)), 2)
print(f"D3D backward: reference {t_refb:.2f} ms, dpf_dcn3d_bwd_data + _bwd_weight {t_newb:.2f} ms, x{t_refb / t_newb:.1f}")
