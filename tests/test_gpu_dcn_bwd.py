"""GPU parity of the D3D backward kernels (dpf_dcn3d_bwd_data / dpf_dcn3d_bwd_weight) against autograd through the oracle's
deform_conv3d (the restatement of deform_conv_cuda_backward, src/module/dcn3d/src/cuda/deform_conv_cuda.cu:128-285)."""
import pytest
import torch

from oracle import dpf_oracle as O

pytestmark = pytest.mark.gpu


def rel_l2(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    return float((got - want).norm() / want.norm().clamp_min(1e-12))


def make_case(cin, shape, seed):
    g = torch.Generator().manual_seed(seed)
    b, d, h, w = shape
    x = torch.randn(b, cin, d, h, w, generator=g).to(torch.bfloat16)
    off = (torch.rand(b, 81, d, h, w, generator=g) - 0.5) * 3.0                       # crosses the borders
    wt = (torch.randn(64, cin, 3, 3, 3, generator=g) * (2.0 / (cin * 27)) ** 0.5).to(torch.bfloat16)
    dy = torch.randn(b, 64, d, h, w, generator=g).to(torch.bfloat16)
    return x, off, wt, dy


def oracle_grads(x, off, wt, dy):
    xf, of, wf = x.float().requires_grad_(), off.clone().requires_grad_(), wt.float().requires_grad_()
    y = O.deform_conv3d(xf, of, wf, None)
    y.backward(dy.float())
    return xf.grad, of.grad, wf.grad


@pytest.mark.parametrize("cin,shape", [(35, (1, 2, 16, 16)), (64, (2, 4, 10, 13)), (64, (1, 3, 33, 20))])
def test_dcn3d_backward(cin, shape):
    from dualpixelface_b200 import ops
    from dualpixelface_b200.ops_dcn_bwd import dcn3d_bwd_data, dcn3d_bwd_weight
    x, off, wt, dy = make_case(cin, shape, 21)
    want_dx, want_doff, want_dw = oracle_grads(x, off, wt, dy)
    b, d, h, w = shape
    xp = torch.zeros(b, d, h, w, 64, dtype=torch.bfloat16)
    xp[..., :cin] = x.permute(0, 2, 3, 4, 1)
    offp = off.permute(0, 2, 3, 4, 1).contiguous().cuda()
    dyp = dy.permute(0, 2, 3, 4, 1).contiguous().cuda()
    dx, doff = dcn3d_bwd_data(xp.cuda(), offp, dyp, wt.cuda())
    dw = dcn3d_bwd_weight(xp.cuda(), offp, dyp, cin)
    torch.cuda.synchronize()
    assert rel_l2(dx[..., :cin].permute(0, 4, 1, 2, 3), want_dx) < 1e-2
    assert rel_l2(doff.permute(0, 4, 1, 2, 3), want_doff) < 1e-2
    assert rel_l2(dw, want_dw) < 1e-2                   # the sampled tile is rounded to bf16 before the MMA
    dx32, doff32 = dcn3d_bwd_data(xp.cuda(), offp, dyp, wt.cuda(), dx_channels=32)
    torch.cuda.synchronize()
    c32 = min(cin, 32)
    assert rel_l2(dx32[..., :c32].permute(0, 4, 1, 2, 3), want_dx[:, :c32]) < 1e-2 and float(dx32[..., 32:].abs().max()) == 0.0
    assert rel_l2(doff32.permute(0, 4, 1, 2, 3), want_doff) < 1e-2


def test_dcn_autograd_function():
    from dualpixelface_b200.ops_dcn_bwd import DCNFn
    x, off, wt, dy = make_case(64, (1, 2, 12, 18), 5)
    want_dx, want_doff, want_dw = oracle_grads(x, off, wt, dy)
    xp = x.permute(0, 2, 3, 4, 1).contiguous().cuda().requires_grad_()
    offp = off.permute(0, 2, 3, 4, 1).contiguous().cuda().requires_grad_()
    wp = wt.float().cuda().requires_grad_()
    z = DCNFn.apply(xp, offp, wp)
    z.backward(dy.permute(0, 2, 3, 4, 1).contiguous().cuda())
    torch.cuda.synchronize()
    assert rel_l2(xp.grad.permute(0, 4, 1, 2, 3), want_dx) < 1e-2
    assert rel_l2(offp.grad.permute(0, 4, 1, 2, 3), want_doff) < 1e-2
    assert rel_l2(wp.grad, want_dw) < 1e-2
