"""GPU parity of the dedicated 2-D tcgen05 convolution (csrc/conv2d_tc.cu, dpf_conv2d_tc_fwd) against torch's fp32 conv2d of the
same bf16-rounded operands (TF32 off): every (Cin, Cout) class the ANM `n_convs` stack uses (normal_module.py:59-66 of the
reference: 64->96->96->64->64->32->3 with dilation 1,2,4,8,1,1), ragged sizes, the fused epilogue, and bit-exact delta weights."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from dualpixelface_b200 import ops as _ops
    _ops.lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _ops


def run_case(ops, cin, cout, dil, shape, seed, bias=False, residual=False, slope=0.1, relu=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    n, h, w = shape
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5).to(torch.bfloat16)
    sc = (torch.rand(cout, device="cuda", generator=g) + 0.5) if bias else None
    sh = (torch.randn(cout, device="cuda", generator=g) * 0.1) if bias else None
    cst = (cout + 7) // 8 * 8
    res = torch.randn(n, h, w, cst, device="cuda", generator=g).to(torch.bfloat16) if residual else None
    want = F.conv2d(x.permute(0, 3, 1, 2).float(), wt.float(), None, 1, dil, dil)
    if bias:
        want = want * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    if residual:
        want = want + res[..., :cout].permute(0, 3, 1, 2).float()
    if relu:
        want = F.leaky_relu(want, slope)
    got = ops.conv2d_tc(x, ops.pack_conv2d_tc_weight(wt), cout, dil, sc, sh, res, relu, slope)
    torch.cuda.synchronize()
    assert got.shape == (n, h, w, cst)
    if cst > cout:
        pad = got[..., cout:].float()
        if residual:
            pad = pad - F.leaky_relu(res[..., cout:].float(), slope) if relu else pad - res[..., cout:].float()
        assert float(pad.abs().max()) == 0.0                       # pad channels: exact zeros (+ the residual, if any)
    err = (got[..., :cout].permute(0, 3, 1, 2).float() - want).abs()
    scale = want.abs().max().item()
    return err.max().item() / scale, err.mean().item() / scale, got


@pytest.mark.parametrize("cin,cout,dil", [(64, 96, 1), (96, 96, 2), (96, 64, 4), (64, 64, 8), (64, 32, 1), (32, 3, 1)])
@pytest.mark.parametrize("shape", [(2, 35, 53), (1, 70, 105)])
def test_anm_nconv_classes(ops, cin, cout, dil, shape):
    emax, emean, _ = run_case(ops, cin, cout, dil, shape, 40 + cin + cout + dil)
    print(f"conv2d_tc {cin}->{cout} d{dil} {shape}: max err {emax:.4f}, mean {emean:.5f} (relative to the output max)")
    assert emax < 6e-3 and emean < 1e-3                            # bf16 output rounding: 2^-9 of the value


@pytest.mark.parametrize("cin,cout,dil,shape", [(32, 32, 3, (2, 33, 47)), (32, 64, 5, (1, 40, 24)), (64, 16, 2, (3, 17, 9)),
                                                (96, 32, 1, (1, 16, 8)), (32, 48, 1, (1, 50, 70)), (32, 96, 7, (1, 29, 31))])
def test_epilogue_and_other_classes(ops, cin, cout, dil, shape):
    emax, emean, _ = run_case(ops, cin, cout, dil, shape, 7, bias=True, residual=True, slope=0.05)
    assert emax < 8e-3 and emean < 1.5e-3
    emax, _, _ = run_case(ops, cin, cout, dil, shape, 8, relu=False)
    assert emax < 6e-3


@pytest.mark.parametrize("dil", [1, 2, 4, 8])
def test_fullsize_delta_weights_exact_and_deterministic(ops, dil):
    """16 x 280 x 420 (the ANM shape at config 2): a weight tensor whose only non-zero tap copies channel c -> c reproduces the
    input shifted by the tap's dilated offset, zero-filled at the border -- bit exact over every tile, residue class and border;
    random weights launched twice give bit-identical outputs (single MMA issuer)."""
    g = torch.Generator(device="cuda").manual_seed(dil)
    x = torch.randn(16, 280, 420, 64, device="cuda", generator=g).to(torch.bfloat16)
    for tap in ((1, 1), (0, 2), (2, 0)):
        wt = torch.zeros(64, 64, 3, 3, device="cuda")
        wt[torch.arange(64), torch.arange(64), tap[0], tap[1]] = 1.0
        y = ops.conv2d_tc(x, ops.pack_conv2d_tc_weight(wt), 64, dil)
        dh, dw = (tap[0] - 1) * dil, (tap[1] - 1) * dil                   # y[h, w] = x[h + dh, w + dw]
        want = torch.zeros_like(y)
        src = x[:, max(dh, 0):280 + min(dh, 0), max(dw, 0):420 + min(dw, 0)]
        want[:, max(-dh, 0):280 + min(-dh, 0), max(-dw, 0):420 + min(-dw, 0)] = src
        assert torch.equal(y, want), (dil, tap)
    wt = (torch.randn(96, 64, 3, 3, device="cuda", generator=g) * 0.05)
    wp = ops.pack_conv2d_tc_weight(wt)
    y1 = ops.conv2d_tc(x, wp, 96, dil, relu=True, slope=0.1)
    y2 = ops.conv2d_tc(x, wp, 96, dil, relu=True, slope=0.1)
    assert torch.equal(y1, y2)
    want = F.leaky_relu(F.conv2d(x[:2].permute(0, 3, 1, 2).float(), wt.to(torch.bfloat16).float(), None, 1, dil, dil), 0.1)
    err = (y1[:2].permute(0, 3, 1, 2).float() - want).abs()
    assert err.max().item() < 6e-3 * want.abs().max().item()


def test_channel_windows(ops):
    """x_coff / y_coff: read 32 channels out of a 96-channel tensor, write 32 channels into a 64-channel tensor."""
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(2, 20, 30, 96, device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn(32, 32, 3, 3, device="cuda", generator=g) * 0.08).to(torch.bfloat16)
    out = torch.full((2, 20, 30, 64), 3.0, device="cuda", dtype=torch.bfloat16)
    ops.conv2d_tc(x, ops.pack_conv2d_tc_weight(wt), 32, 2, out=out, y_coff=32, x_coff=64)
    want = F.conv2d(x[..., 64:].permute(0, 3, 1, 2).float(), wt.float(), None, 1, 2, 2)
    err = (out[..., 32:].permute(0, 3, 1, 2).float() - want).abs().max().item()
    assert err < 6e-3 * want.abs().max().item() and float((out[..., :32].float() - 3.0).abs().max()) == 0.0
