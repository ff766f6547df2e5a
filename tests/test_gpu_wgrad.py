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
