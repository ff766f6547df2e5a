"""GPU checks at BASELINE.json's FULL sizes (config 2: 1120x1680, quarter resolution 280x420, D = 8).

The small-size parity tests compare against the CPU oracle; at full size the CPU oracle would take minutes, so this file uses
  * size-independent exact properties (delta-weight convolutions reproduce / shift their input bit-exactly over every tile
    and border; the shifted volume equals row-rolled features; zero-offset D3D with delta weights is the identity), and
  * the oracle's own torch code executed on the GPU in fp32 (TF32 off) as the checker for the floating-point stages and for
    one whole-model pass.
"""
import pytest
import torch

from dualpixelface_b200.synthetic import synthetic_batch
from oracle import dpf_oracle as O

pytestmark = pytest.mark.gpu

H4, W4, D = 280, 420, 8


@pytest.fixture(scope="module")
def ops():
    from dualpixelface_b200 import ops as _ops
    _ops.lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _ops


def _rand(shape, seed, relu=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randn(*shape, device="cuda", generator=g)
    return (torch.relu(t) if relu else t).to(torch.bfloat16)


def test_fullsize_costvol_is_row_rolled_features(ops):
    """concat volume: level i holds ref[h] and tgt[h + s_i] on the rows where both exist, zero elsewhere -- bit-exact."""
    shifts = [int(c) for c in O.cost_range(-4, 12, D)]
    ref, tgt = _rand((4, H4, W4, 32), 1), _rand((4, H4, W4, 32), 2)
    # the formula itself is pinned against the oracle at a small size first
    small = O.psm_concat_volume(ref[:1, :24, :16].permute(0, 3, 1, 2).float().cpu(), tgt[:1, :24, :16].permute(0, 3, 1, 2).float().cpu(),
                                O.cost_range(-4, 12, D))

    def expected(r, t):
        vol = torch.zeros(r.shape[0], D, r.shape[1], r.shape[2], 64, device=r.device, dtype=r.dtype)
        hh = r.shape[1]
        for i, s in enumerate(shifts):
            lo, hi = max(-s, 0), hh - max(s, 0)                               # rows h with 0 <= h + s < H
            vol[:, i, lo:hi, :, :32] = r[:, lo:hi]
            vol[:, i, lo:hi, :, 32:] = t[:, lo + s:hi + s]
        return vol

    assert torch.equal(expected(ref[:1, :24, :16], tgt[:1, :24, :16]).permute(0, 4, 1, 2, 3).float().cpu(), small)
    got = ops.costvol_fwd(ref, tgt, shifts, "concat")
    assert torch.equal(got, expected(ref, tgt))


@pytest.mark.parametrize("cin,cout", [(32, 32), (64, 32), (64, 64)])
def test_fullsize_conv_delta_weights_are_exact(ops, cin, cout):
    """3x3x3 conv whose only non-zero tap copies channel c -> c: centre tap = identity, corner tap = a (d,h,w) shift with
    zero fill.  Exact in bf16; exercises every tile, plane and border of the kd-fused kernel at 4x8x280x420."""
    from dualpixelface_b200.layers import KIND_3x3x3, TCConv3d
    x = _rand((4, D, H4, W4, cin), 3, relu=False)
    n = min(cin, cout)
    for tap in ((1, 1, 1), (0, 0, 0), (2, 2, 0)):
        w = torch.zeros(cout, cin, 3, 3, 3, device="cuda")
        w[torch.arange(n), torch.arange(n), tap[0], tap[1], tap[2]] = 1.0
        y = TCConv3d(w, KIND_3x3x3)(x)
        dd, dh, dw = tap[0] - 1, tap[1] - 1, tap[2] - 1                     # y[v] = x[v + (dd,dh,dw)]
        want = torch.zeros_like(y)
        src = x[:, max(dd, 0):D + min(dd, 0), max(dh, 0):H4 + min(dh, 0), max(dw, 0):W4 + min(dw, 0), :n]
        want[:, max(-dd, 0):D + min(-dd, 0), max(-dh, 0):H4 + min(-dh, 0), max(-dw, 0):W4 + min(-dw, 0), :n] = src
        assert torch.equal(y, want), tap


def test_fullsize_strided_conv_pair_delta(ops):
    """stride-2 conv with the centre delta picks the even voxels; the transposed conv with the centre delta puts them back
    on the even voxels of the fine grid (zeros elsewhere) -- exact, full aggregation size."""
    from dualpixelface_b200.layers import KIND_S2, KIND_T2, TCConv3d
    x = _rand((4, D, H4, W4, 32), 4, relu=False)
    w = torch.zeros(64, 32, 3, 3, 3, device="cuda")
    w[torch.arange(32), torch.arange(32), 1, 1, 1] = 1.0
    y = TCConv3d(w, KIND_S2)(x)                                              # [4,4,140,210,64]
    assert torch.equal(y[..., :32], x[:, ::2, ::2, ::2]) and float(y[..., 32:].abs().max()) == 0.0
    wt = torch.zeros(64, 32, 3, 3, 3, device="cuda")                         # ConvTranspose3d layout [Cin, Cout, k]
    wt[torch.arange(32), torch.arange(32), 1, 1, 1] = 1.0
    z = TCConv3d(wt, KIND_T2, transposed=True)(y)                            # [4,8,280,420,32]
    want = torch.zeros_like(z)
    want[:, ::2, ::2, ::2] = x[:, ::2, ::2, ::2]
    assert torch.equal(z, want)


def test_fullsize_dcn_zero_offsets_delta_is_identity(ops):
    x = _rand((4, 4, H4, W4, 64), 5, relu=False)
    off = torch.zeros(4, 4, H4, W4, 96, device="cuda")
    w = torch.zeros(64, 64, 3, 3, 3, device="cuda")
    w[torch.arange(64), torch.arange(64), 1, 1, 1] = 1.0
    y = ops.dcn3d(x, off, ops.pack_conv_weight(w, cin_pad=64), 64)
    assert torch.equal(y, x)
    off[..., 13 * 3 + 2] = 1.0                                               # centre tap samples (d, h, w + 1): integer shift
    y = ops.dcn3d(x, off, ops.pack_conv_weight(w, cin_pad=64), 64)
    want = torch.zeros_like(x)
    want[:, :, :, :-1] = x[:, :, :, 1:]
    assert torch.equal(y, want)


def test_fullsize_regression_vs_oracle_code_on_gpu(ops):
    g = torch.Generator(device="cuda").manual_seed(6)
    cost = torch.randn(4, D, H4, W4, device="cuda", generator=g) * 3.0
    disp, _ = ops.regress_fwd(cost, -4.0, 0.5, False)
    up = O.upsample_cost(cost.unsqueeze(1))                                  # oracle code, executed on the GPU in fp32
    want, _ = O.regression(up, O.disparity_bins(-4, 12, D))
    assert (disp - want).abs().max().item() < 1e-3


def test_fullsize_model_vs_oracle_code_on_gpu(ops):
    """One 1120x1680 pair through the whole StereoDPNet path vs the oracle's torch code run on the GPU in fp32."""
    from test_gpu_models import build, calibrated_state
    st, fwd = calibrated_state("stereodpnet", synthetic_batch(2, 64, 96, training=True, seed=0))   # stats from a CPU pass
    batch = {k: v.cuda() for k, v in synthetic_batch(1, 1120, 1680, training=True, seed=3).items()}
    with torch.no_grad():
        want = fwd(dict(batch), {k: v.cuda() for k, v in st.items()}, False)
    model = build("stereodpnet")
    model.load_state_dict(st, strict=False)
    model.cuda().eval()
    model.encoder_autocast = False
    with torch.no_grad():
        got = model(batch)
    err = (got["pred_depth"].float() - want["pred_depth"]).abs()
    n_err = (got["pred_normal"].float() - want["pred_normal"]).abs()
    print(f"full size: disparity max err {err.max():.4f} mean {err.mean():.5f}; normal max err {n_err.max():.4f} mean {n_err.mean():.5f}")
    assert err.max().item() < 0.32 and err.mean().item() < 0.032             # measured 0.177 / 0.0182 px (2e-2 of 16 px = 0.32)
    assert n_err.mean().item() < 0.03                                        # measured 0.0149


@pytest.mark.parametrize("kind,cin,cout", [("s1", 32, 32), ("s1", 64, 32), ("s2", 32, 64), ("t2", 64, 32)])
def test_fullsize_conv_random_vs_torch_fp32(ops, kind, cin, cout):
    """Random weights at the full aggregation size against torch's fp32 convolution of the same bf16-rounded operands (TF32
    off).  Several MMA-issuing warps accumulate into the same TMEM accumulators: a lost update would show up here as an error
    of a whole tap's contribution (percent level), far above the bf16 output rounding (2^-9 of the value)."""
    import torch.nn.functional as F
    from dualpixelface_b200.layers import KIND_3x3x3, KIND_S2, KIND_T2, TCConv3d
    shape = (2, 4, 140, 210) if kind == "t2" else (2, D, H4, W4)
    x = _rand((*shape, cin), 11, relu=False)
    g = torch.Generator(device="cuda").manual_seed(12)
    xc = x.permute(0, 4, 1, 2, 3).float()
    if kind == "t2":
        w = (torch.randn(cin, cout, 3, 3, 3, device="cuda", generator=g) * 0.05).to(torch.bfloat16).float()
        want = F.conv_transpose3d(xc, w, stride=2, padding=1, output_padding=1)
        got = TCConv3d(w, KIND_T2, transposed=True)(x)
    else:
        w = (torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) * 0.05).to(torch.bfloat16).float()
        want = F.conv3d(xc, w, stride=2 if kind == "s2" else 1, padding=1)
        got = TCConv3d(w, KIND_S2 if kind == "s2" else KIND_3x3x3)(x)
    err = (got.permute(0, 4, 1, 2, 3).float() - want).abs()
    scale = want.abs().max().item()
    print(f"{kind} {cin}->{cout}: max abs err {err.max().item():.4g} (output max {scale:.3g}), mean {err.mean().item():.3g}")
    assert err.max().item() < 6e-3 * scale and err.mean().item() < 1e-3 * scale
