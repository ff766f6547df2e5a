"""GPU parity of the tcgen05 weight-gradient kernel (dpf_conv3d_wgrad) against PyTorch autograd."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,cin,cout,shape", [
    (4, 64, 32, (1, 1, 16, 16)), (4, 32, 32, (2, 3, 20, 37)), (3, 32, 32, (1, 2, 16, 16)), (0, 32, 32, (2, 8, 37, 53)),
    (0, 64, 32, (1, 4, 18, 26)), (0, 32, 16, (1, 4, 18, 26)), (0, 32, 1, (2, 8, 20, 30)), (0, 64, 64, (1, 4, 18, 26)),
    (0, 32, 32, (3, 1, 5, 7))])
def test_wgrad_matches_autograd(kind, cin, cout, shape):
    from dualpixelface_b200.ops_wgrad import conv3d_wgrad
    g = torch.Generator().manual_seed(5)
    b, d, h, w = shape
    x = torch.randn(b, cin, d, h, w, generator=g).to(torch.bfloat16)
    dz = torch.randn(b, cout, d, h, w, generator=g).to(torch.bfloat16)
    ks = {0: (3, 3, 3), 3: (1, 3, 3), 4: (1, 1, 1)}[kind]
    wt = torch.zeros(cout, cin, *ks, requires_grad=True)
    F.conv3d(x.float(), wt, padding=tuple(k // 2 for k in ks)).backward(dz.float())
    got = conv3d_wgrad(x.permute(0, 2, 3, 4, 1).contiguous().cuda(), dz.permute(0, 2, 3, 4, 1).contiguous().cuda(), kind).cpu()
    assert got.shape == wt.grad.shape
    assert ((got - wt.grad).norm() / wt.grad.norm()).item() < 1e-3          # bf16 operands are exact products, fp32 accumulation


@pytest.mark.parametrize("cin,cout,shape", [(32, 32, (1, 2, 16, 32)), (32, 64, (2, 8, 36, 52)), (64, 64, (1, 4, 18, 34)),
                                            (32, 32, (1, 3, 9, 21))])
def test_wgrad_stride2_matches_autograd(cin, cout, shape):
    """kind 1: dz lives on the ceil(x/2) grid (odd sizes included: the reference only uses even ones)."""
    from dualpixelface_b200.layers import KIND_S2
    from dualpixelface_b200.ops_wgrad import conv3d_wgrad
    g = torch.Generator().manual_seed(6)
    b, d, h, w = shape
    x = torch.randn(b, cin, d, h, w, generator=g).to(torch.bfloat16)
    wt = torch.zeros(cout, cin, 3, 3, 3, requires_grad=True)
    y = F.conv3d(x.float(), wt, stride=2, padding=1)
    dz = torch.randn(y.shape, generator=g).to(torch.bfloat16)
    y.backward(dz.float())
    got = conv3d_wgrad(x.permute(0, 2, 3, 4, 1).contiguous().cuda(), dz.permute(0, 2, 3, 4, 1).contiguous().cuda(), KIND_S2).cpu()
    assert got.shape == wt.grad.shape
    assert ((got - wt.grad).norm() / wt.grad.norm()).item() < 1e-3


@pytest.mark.parametrize("cin,cout,shape", [(64, 64, (1, 2, 9, 13)), (64, 32, (2, 4, 18, 26)), (32, 32, (1, 1, 8, 16))])
def test_wgrad_transposed_matches_autograd(cin, cout, shape):
    """kind 2 (ConvTranspose3d k3 s2 p1 op1): the stride-2 kernel with x and dz swapped; result in the [Cin,Cout,3,3,3] layout."""
    from dualpixelface_b200.layers import KIND_T2
    from dualpixelface_b200.ops_wgrad import conv3d_wgrad
    g = torch.Generator().manual_seed(7)
    b, d, h, w = shape
    x = torch.randn(b, cin, d, h, w, generator=g).to(torch.bfloat16)
    wt = torch.zeros(cin, cout, 3, 3, 3, requires_grad=True)
    y = F.conv_transpose3d(x.float(), wt, stride=2, padding=1, output_padding=1)
    dz = torch.randn(y.shape, generator=g).to(torch.bfloat16)
    y.backward(dz.float())
    got = conv3d_wgrad(x.permute(0, 2, 3, 4, 1).contiguous().cuda(), dz.permute(0, 2, 3, 4, 1).contiguous().cuda(), KIND_T2).cpu()
    assert got.shape == wt.grad.shape
    assert ((got - wt.grad).norm() / wt.grad.norm()).item() < 1e-3
