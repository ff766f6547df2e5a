"""GPU parity: each sm_100a kernel, called through the C ABI, against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dpf_oracle as O

pytestmark = pytest.mark.gpu

CR = O.cost_range(-4, 12, 8)
SHIFTS = [int(c) for c in CR]


@pytest.fixture(scope="module")
def ops():
    from dualpixelface_b200 import ops as _ops
    _ops.lib()
    return _ops


def feat(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(*shape, generator=g)).to(torch.bfloat16)


def nhwc(x):      # [B,C,H,W] -> [B,H,W,C] contiguous
    return x.permute(0, 2, 3, 1).contiguous()


def ndhwc(x):     # [B,C,D,H,W] -> [B,D,H,W,C]
    return x.permute(0, 2, 3, 4, 1).contiguous()


def from_ndhwc(y):
    return y.permute(0, 4, 1, 2, 3).contiguous()


def rel_err(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-6)).item()


# ---------------------------------------------------------------------------------------------------- volume
@pytest.mark.parametrize("shape", [(2, 32, 12, 10), (1, 32, 37, 75), (2, 32, 112, 112)])
def test_costvol_concat_diff_exact(ops, shape):
    ref, tgt = feat(shape, 21), feat(shape, 22)
    want_c = O.psm_concat_volume(ref.float(), tgt.float(), CR)
    want_d = O.diff_volume(ref.float(), tgt.float(), CR).to(torch.bfloat16)
    got_c = ops.costvol_fwd(nhwc(ref).cuda(), nhwc(tgt).cuda(), SHIFTS, "concat")
    got_d = ops.costvol_fwd(nhwc(ref).cuda(), nhwc(tgt).cuda(), SHIFTS, "diff")
    assert torch.equal(from_ndhwc(got_c).float().cpu(), want_c)                      # bit-exact copy semantics
    assert torch.equal(from_ndhwc(got_d).cpu(), want_d)                               # one bf16 rounding of an fp32 difference


@pytest.mark.parametrize("c,groups", [(32, 8), (32, 16), (32, 32), (64, 8), (64, 32), (128, 8), (16, 8)])
def test_costvol_gwc(ops, c, groups):
    """every lanes-per-group / groups-per-lane class of the warp-shuffle correlation: 4, 2, 1 channels per group (several group
    sums per lane), 8 (one lane per group) and 16 (xor-shuffle reduction over two lanes); ragged widths (33 = partial strip)."""
    ref, tgt = feat((2, c, 20, 33), 23), feat((2, c, 20, 33), 24)
    want = O.psm_gwc_volume(ref.float(), tgt.float(), CR, groups)
    got = ops.costvol_fwd(nhwc(ref).cuda(), nhwc(tgt).cuda(), SHIFTS, "gwc", groups)
    assert rel_err(from_ndhwc(got), want) < 1e-2                                      # 1 bf16 ulp of the output


@pytest.mark.parametrize("mode,groups", [("concat", 0), ("diff", 0), ("gwc", 8)])
def test_costvol_bwd(ops, mode, groups):
    ref, tgt = feat((2, 32, 14, 19), 25), feat((2, 32, 14, 19), 26)
    r, t = ref.float().requires_grad_(True), tgt.float().requires_grad_(True)
    vol = {"concat": lambda: O.psm_concat_volume(r, t, CR), "diff": lambda: O.diff_volume(r, t, CR),
           "gwc": lambda: O.psm_gwc_volume(r, t, CR, groups)}[mode]()
    g = torch.Generator().manual_seed(27)
    dvol = torch.randn(vol.shape, generator=g).to(torch.bfloat16)
    vol.backward(dvol.float())
    dref, dtgt = ops.costvol_bwd(nhwc(ref).cuda(), nhwc(tgt).cuda(), ndhwc(dvol).cuda(), SHIFTS, mode, groups)
    assert rel_err(dref.permute(0, 3, 1, 2), r.grad) < 1e-2
    assert rel_err(dtgt.permute(0, 3, 1, 2), t.grad) < 1e-2


# ------------------------------------------------------------------------------------------------- regression
@pytest.mark.parametrize("shape", [(2, 8, 20, 28), (1, 8, 70, 105)])
def test_regress_fwd_bwd(ops, shape):
    g = torch.Generator().manual_seed(51)
    cost = (torch.randn(*shape, generator=g) * 2.0).requires_grad_(True)
    bins = O.disparity_bins(-4, 12, 8)
    full = O.upsample_cost(cost.unsqueeze(1))
    want_d, want_p = O.regression(full, bins)
    disp, prob = ops.regress_fwd(cost.detach().cuda(), -4.0, 0.5, want_prob=True)
    assert (disp.cpu() - want_d).abs().max().item() < 1e-3 * 11.5
    assert (prob.cpu() - want_p).abs().max().item() < 1e-4
    gd = torch.randn(want_d.shape, generator=g)
    want_d.backward(gd)
    dcost = ops.regress_bwd(cost.detach().cuda(), gd.cuda(), -4.0, 0.5)
    assert rel_err(dcost, cost.grad) < 1e-3


@pytest.mark.parametrize("shape", [(1, 8, 16, 24), (2, 8, 37, 53)])
def test_regress_fwd_halfpixel(ops, shape):
    """dpf_regress_fwd_halfpixel == F.interpolate(scale_factor=4, 'trilinear', align_corners=False) + softmax + expectation (NNet,
    src/model/nnet/mainmodel.py:149-152)."""
    g = torch.Generator().manual_seed(52)
    cost = torch.randn(*shape, generator=g) * 2.0
    bins = O.disparity_bins(-4, 12, 8)
    full = F.interpolate(cost.unsqueeze(1), scale_factor=4, mode="trilinear", align_corners=False).squeeze(1)
    want_d, want_p = O.regression(full, bins)
    disp, prob = ops.regress_fwd(cost.cuda(), -4.0, 0.5, want_prob=True, align_corners=False)
    assert (disp.cpu() - want_d).abs().max().item() < 1e-3 * 11.5
    assert (prob.cpu() - want_p).abs().max().item() < 1e-4


@pytest.mark.parametrize("align", [True, False])
def test_regress_fwd_degenerate_costs(ops, align):
    """Adjacent levels thousands apart (an uncalibrated network): no bin lands exactly on an interior knot, so every bin underflows
    against the maximum knot -- the kernel then redoes its sums against the maximum bin, as torch's softmax does (finite output)."""
    g = torch.Generator().manual_seed(53)
    cost = torch.randn(1, 8, 9, 12, generator=g) * 3.0e4
    bins = O.disparity_bins(-4, 12, 8)
    full = F.interpolate(cost.unsqueeze(1), scale_factor=4, mode="trilinear", align_corners=align).squeeze(1)
    want_d, want_p = O.regression(full, bins)
    disp, prob = ops.regress_fwd(cost.cuda(), -4.0, 0.5, want_prob=True, align_corners=align)
    assert torch.isfinite(disp).all() and torch.isfinite(prob).all()
    # interpolated bins of such costs carry ~1e-2 absolute fp32 round-off, which the softmax turns into percent-level probability
    # differences where two bins tie; away from ties the winner-takes-all result is exact
    assert (disp.cpu() - want_d).abs().median().item() < 1e-3 and (disp.cpu() - want_d).abs().max().item() < 0.5


# ------------------------------------------------------------------------------------------------ convolution
def conv_case(ops, cin, cout, kind, shape, seed, with_affine=True, residual=False, relu=True, out_f32=False):
    b, d, h, w = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, cin, d, h, w, generator=g).to(torch.bfloat16)
    ks = {0: (3, 3, 3), 3: (1, 3, 3), 4: (1, 1, 1)}[kind]
    wt = (torch.randn(cout, cin, *ks, generator=g) * (2.0 / (cin * ks[0] * ks[1] * ks[2])) ** 0.5).to(torch.bfloat16)
    scale = (torch.rand(cout, generator=g) + 0.5) if with_affine else None
    shift = (torch.randn(cout, generator=g) * 0.2) if with_affine else None
    res = torch.randn(b, cout, d, h, w, generator=g) if residual else None
    if res is not None and not out_f32:
        res = res.to(torch.bfloat16)
    want = F.conv3d(x.float(), wt.float(), padding=tuple(k // 2 for k in ks))
    if with_affine:
        want = want * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    if residual:
        want = want + res.float()
    if relu:
        want = F.relu(want)
    got = ops.conv3d(ndhwc(x).cuda(), ops.pack_conv_weight(wt.cuda(), kind=kind), kind, cout,
                     scale.cuda() if with_affine else None, shift.cuda() if with_affine else None,
                     ndhwc(res).cuda() if residual else None, relu, out_f32=out_f32)
    torch.cuda.synchronize()
    return rel_err(from_ndhwc(got), want)


def test_conv_pointwise_is_plain_gemm(ops):
    """1x1x1 conv == one GEMM: isolates the UMMA descriptor / TMEM plumbing from the tap addressing."""
    assert conv_case(ops, 32, 32, 4, (1, 2, 16, 24), 1, with_affine=False, relu=False) < 1e-2


@pytest.mark.parametrize("cin,cout", [(32, 32), (64, 32), (32, 64), (32, 16), (64, 16)])
def test_conv_3x3x3(ops, cin, cout):
    assert conv_case(ops, cin, cout, 0, (1, 4, 16, 24), 2) < 1e-2


@pytest.mark.parametrize("shape", [(2, 8, 37, 53), (1, 2, 70, 105), (1, 1, 5, 7), (3, 3, 16, 8)])
def test_conv_ragged_shapes(ops, shape):
    assert conv_case(ops, 32, 32, 0, shape, 3, residual=True) < 1e-2
    assert conv_case(ops, 64, 32, 0, shape, 4, residual=False) < 1e-2


def test_conv_head_fp32_out(ops):
    assert conv_case(ops, 32, 1, 0, (2, 8, 20, 30), 5, with_affine=False, residual=True, relu=False, out_f32=True) < 1e-2


@pytest.mark.parametrize("kind", [3, 4])
def test_conv_planar_pointwise(ops, kind):
    assert conv_case(ops, 32, 32, kind, (2, 3, 21, 35), 6) < 1e-2


def test_conv_many_tiles_persistent(ops):
    """More tiles than SMs: exercises the ring / TMEM phase wrap-around of the persistent CTAs."""
    assert conv_case(ops, 32, 32, 0, (2, 8, 112, 112), 7, residual=True) < 1e-2


# --------------------------------------------------------------------------------------------------------- ASM
@pytest.mark.parametrize("direction", ["forward", "backward"])
def test_asm_sample(ops, direction):
    from dualpixelface_b200.shift_tables import build_tables
    x = feat((2, 32, 16, 24), 11)
    want = torch.stack(O.subpixel_samples(x.float(), -1.0, direction), 1)           # [B,S,C,H,W]
    tb = {k: v.cuda() for k, v in build_tables(16, 24, -1.0, direction).items()}
    got = ops.asm_sample(nhwc(x).cuda(), tb)                                          # [B,S,H,W,C]
    err = (got.permute(0, 1, 4, 2, 3).float().cpu() - want).abs().max().item()
    assert err < 2e-2 * want.abs().max().item()


def test_channel_stats_and_blend(ops):
    g = torch.Generator().manual_seed(91)
    b, s, h, w, c = 2, 3, 13, 17, 32
    smp = torch.randn(b, s, h, w, c, generator=g).to(torch.bfloat16)
    lg = torch.randn(b, s, h, w, c, generator=g).to(torch.bfloat16)
    st = ops.channel_stats(lg.cuda()).cpu()
    lf = lg.float().reshape(b, -1, c)
    assert torch.allclose(st[..., 0], lf.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[..., 1], (lf * lf).sum(1), rtol=1e-4, atol=1e-2)
    a = torch.rand(b, c, generator=g) + 0.5
    dd = torch.randn(b, c, generator=g) * 0.3
    vol = torch.zeros(b, 8, h, w, 64, dtype=torch.bfloat16, device="cuda")
    ops.asm_blend(smp.cuda(), lg.cuda(), a.cuda(), dd.cuda(), vol, 0, 8, 32)
    gate = torch.sigmoid(lg.float() * a.view(b, 1, 1, 1, c) + dd.view(b, 1, 1, 1, c))
    want = (smp.float() * torch.softmax(gate, dim=1)).mean(1)
    for d in (0, 7):
        assert (vol[:, d, :, :, 32:].float().cpu() - want).abs().max().item() < 2e-2
    assert vol[..., :32].abs().max().item() == 0


# --------------------------------------------------------------------------------------------------------- ANM
def test_anm_select_gather(ops):
    g = torch.Generator().manual_seed(61)
    b, h4, w4, c = 2, 10, 12, 32
    out3 = torch.randn(b, c, 8, h4, w4, generator=g).to(torch.bfloat16)
    disp = torch.rand(b, 4 * h4, 4 * w4, generator=g) * 14.0 - 4.5
    from dualpixelface_b200.synthetic import synthetic_batch
    batch = synthetic_batch(b, 4 * h4, 4 * w4, seed=3)
    kq = batch["K"].clone()
    kq[:, :2] = kq[:, :2] / 4.0
    kinv = torch.inverse(kq).contiguous()
    idx, coord, minmax = ops.anm_select(disp.cuda(), kinv.cuda(), batch["abvalue"].cuda(), [float(v) for v in CR], 4)
    fv = ops.anm_gather(ndhwc(out3).cuda(), idx, coord, minmax, 64)
    dq = F.interpolate(disp.unsqueeze(1), scale_factor=0.25, mode="nearest") * 0.25
    crt = torch.as_tensor(CR, dtype=torch.float32).view(1, -1, 1, 1)
    want_idx = O.anm_select_levels(dq, crt, 4)
    assert torch.equal(idx.cpu().long(), want_idx)                                    # generic (off-level) disparities: exact
    sel_disp = torch.gather(crt.expand(b, 8, h4, w4), 1, want_idx)
    want_coord = O.anm_coord_volume(sel_disp, batch["K"], batch["abvalue"])          # [B,K,3,H,W]
    got_coord = fv[..., 32:35].float().cpu().permute(0, 1, 4, 2, 3)
    assert (got_coord - want_coord).abs().max().item() < 1e-2
    want_cost = torch.gather(out3.permute(0, 2, 1, 3, 4), 1, want_idx.unsqueeze(2).expand(-1, -1, c, -1, -1))
    assert torch.equal(fv[..., :32].cpu().permute(0, 1, 4, 2, 3), want_cost)
    assert fv[..., 35:].abs().max().item() == 0


def test_anm_select_tie_rule(ops):
    """d exactly on a level: the K-th pick is a two-way tie; the documented rule is 'lower level index wins'."""
    b, h4, w4 = 1, 4, 4
    disp = torch.full((b, 16, 16), 4.0)            # quarter-res 1.0 == level 4 -> candidates {3,4,5} + tie {2,6}
    kinv = torch.eye(3).unsqueeze(0).contiguous()
    ab = torch.tensor([[32.98, -26996.49]])
    idx, _, _ = ops.anm_select(disp.cuda(), kinv.cuda(), ab.cuda(), [float(v) for v in CR], 4)
    assert idx[0, :, 0, 0].cpu().tolist() == [2, 3, 4, 5]


# ------------------------------------------------------------------------------ stride-2 / transposed / D3D
def strided_case(ops, cin, cout, shape, seed, transposed, residual=False):
    from dualpixelface_b200.layers import KIND_S2, KIND_T2, TCConv3d
    b, d, h, w = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, cin, d, h, w, generator=g).to(torch.bfloat16)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.2
    if transposed:
        wt = (torch.randn(cin, cout, 3, 3, 3, generator=g) * (2.0 / (cin * 27 / 8)) ** 0.5).to(torch.bfloat16)
        want = F.conv_transpose3d(x.float(), wt.float(), stride=2, padding=1, output_padding=1)
    else:
        wt = (torch.randn(cout, cin, 3, 3, 3, generator=g) * (2.0 / (cin * 27)) ** 0.5).to(torch.bfloat16)
        want = F.conv3d(x.float(), wt.float(), stride=2, padding=1)
    want = want * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    res = None
    if residual:
        res = torch.randn(want.shape, generator=g).to(torch.bfloat16)
        want = want + res.float()
    want = F.relu(want)
    layer = TCConv3d(wt.cuda(), KIND_T2 if transposed else KIND_S2, transposed=transposed)
    got = layer(ndhwc(x).cuda(), scale.cuda(), shift.cuda(), residual=ndhwc(res).cuda() if residual else None, relu=True)
    torch.cuda.synchronize()
    assert tuple(got.shape[1:4]) == tuple(want.shape[2:])
    return rel_err(from_ndhwc(got), want)


@pytest.mark.parametrize("cin,cout,shape", [(32, 64, (1, 8, 32, 16)), (32, 64, (2, 8, 36, 52)), (64, 64, (1, 4, 18, 26)),
                                            (32, 32, (1, 2, 70, 106)), (32, 64, (1, 5, 35, 53)), (64, 64, (2, 7, 21, 19)),
                                            (64, 32, (1, 1, 17, 9)), (32, 16, (1, 3, 16, 16))])
def test_conv_stride2(ops, cin, cout, shape):
    """plane-streamed stride-2 kernel (dpf_conv3d_s2_fwd): even and ODD depths / heights / widths (the last output plane then has
    no kd = 2 plane), one and two 32-channel k-parts, 16 / 32 / 64 output channels."""
    assert strided_case(ops, cin, cout, shape, 11, transposed=False) < 1e-2


@pytest.mark.parametrize("cin,cout,shape", [(64, 64, (1, 2, 16, 8)), (64, 32, (2, 4, 18, 26)), (64, 64, (1, 2, 35, 53)),
                                            (32, 32, (1, 1, 9, 13))])
def test_conv_transposed(ops, cin, cout, shape):
    assert strided_case(ops, cin, cout, shape, 12, transposed=True, residual=True) < 1e-2


@pytest.mark.parametrize("cin,shape,off_c", [(35, (2, 4, 10, 13), 81), (64, (2, 4, 10, 13), 81), (35, (2, 4, 10, 13), 96),
                                             (64, (2, 4, 70, 69), 96), (35, (1, 4, 150, 101), 84)])
def test_dcn3d(ops, cin, shape, off_c):
    """off_c = 81: offsets read straight from global memory; off_c = 96 / 84 (16-byte aligned voxel rows): the staged path --
    offsets double-buffered in shared memory by halves; the larger shapes give every CTA several work units, which exercises
    the refill-while-in-use protocol of the two halves."""
    g = torch.Generator().manual_seed(13)
    b, d, h, w = shape
    cpad, cs = 64, 64
    x = torch.randn(b, cin, d, h, w, generator=g).to(torch.bfloat16)
    off = (torch.rand(b, 81, d, h, w, generator=g) - 0.5) * 3.0                       # up to +-1.5 voxels, crosses borders
    wt = (torch.randn(64, cin, 3, 3, 3, generator=g) * (2.0 / (cin * 27)) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(64, generator=g) * 0.1
    want = F.relu(O.deform_conv3d(x.float(), off, wt.float(), bias))
    xp = torch.zeros(b, d, h, w, cs, dtype=torch.bfloat16)
    xp[..., :cin] = ndhwc(x)
    offp = torch.full((b, d, h, w, off_c), 7.0)                                       # pad channels: garbage that must not be read
    offp[..., :81] = off.permute(0, 2, 3, 4, 1)
    got = ops.dcn3d(xp.cuda(), offp.cuda(), ops.pack_conv_weight(wt.cuda(), cin_pad=cpad), cpad,
                    torch.ones(64).cuda(), bias.cuda(), relu=True)
    torch.cuda.synchronize()
    assert rel_err(from_ndhwc(got), want) < 2e-2        # the gathered A tile is rounded to bf16 before the MMA


def test_pyramid_cat_matches_interpolate():
    """dpf_pyramid_cat == cat([f1, bilinear x2, bilinear x4]) with align_corners=True (bf16 rounding of the fp32 blend)."""
    from dualpixelface_b200.encoder_fused import pyramid_cat
    g = torch.Generator().manual_seed(61)
    f1, f2, f3 = (torch.randn(2, 32, 20 // s, 28 // s, generator=g).to(torch.bfloat16) for s in (1, 2, 4))
    up = lambda t, s: F.interpolate(t.float(), scale_factor=s, mode="bilinear", align_corners=True)
    want = torch.cat([f1.float(), up(f2, 2), up(f3, 4)], 1)
    cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
    got = pyramid_cat(cl(f1), cl(f2), cl(f3)).float().cpu()
    assert got.shape == want.shape
    assert torch.equal(got[:, :32], f1.float())
    assert (got - want).abs().max().item() < 2e-2


def test_fpn_merge_matches_torch():
    """dpf_fpn_merge == (lateral + bias) + F.interpolate(top, size, 'nearest'), bf16."""
    from dualpixelface_b200.encoder_fused import fpn_merge
    g = torch.Generator().manual_seed(62)
    lat = torch.randn(2, 32, 18, 26, generator=g).to(torch.bfloat16)
    top = torch.randn(2, 32, 9, 13, generator=g).to(torch.bfloat16)
    bias = torch.randn(32, generator=g)
    want = ((lat.float() + bias.view(1, -1, 1, 1)).to(torch.bfloat16).float() + F.interpolate(top.float(), size=(18, 26), mode="nearest")).to(torch.bfloat16)
    cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
    got = fpn_merge(cl(lat), bias.cuda(), cl(top)).cpu()
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape,cout", [((2, 37, 53), 32), ((1, 16, 24), 32), ((3, 70, 105), 16), ((2, 280, 420), 32)])
def test_conv2d_rows_matches_torch(ops, shape, cout):
    """dpf_conv2d_fwd (row-streamed 2-D mode of the kd-fused kernel: 16 row streams, halo planes, centre-row taps) against
    torch's fp32 conv2d of the same bf16 operands, with bias, residual and LeakyReLU; exact delta-weight check included."""
    n, h, w = shape
    g = torch.Generator().manual_seed(71)
    x = torch.randn(n, 32, h, w, generator=g).to(torch.bfloat16)
    wt = (torch.randn(cout, 32, 3, 3, generator=g) * 0.06).to(torch.bfloat16)
    bias = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(n, cout, h, w, generator=g).to(torch.bfloat16)
    want = F.leaky_relu(F.conv2d(x.float(), wt.float(), bias, padding=1) + res.float(), 0.05)
    xc = x.permute(0, 2, 3, 1).contiguous().cuda()
    got = ops.conv2d_rows(xc, ops.pack_conv2d_weight(wt.cuda()), cout, None, bias.cuda(), res.permute(0, 2, 3, 1).contiguous().cuda(),
                          relu=True, slope=0.05)
    err = (got.permute(0, 3, 1, 2).float().cpu() - want).abs()
    assert err.max().item() < 6e-3 * want.abs().max().item(), err.max().item()
    # delta weights: the (kh, kw) = (0, 2) tap copies channel c of pixel (y-1, x+1): exact, exercises every stream boundary
    wd = torch.zeros(cout, 32, 3, 3)
    wd[torch.arange(cout), torch.arange(cout), 0, 2] = 1.0
    got = ops.conv2d_rows(xc, ops.pack_conv2d_weight(wd.cuda()), cout).permute(0, 3, 1, 2).cpu()
    want = torch.zeros(n, cout, h, w, dtype=torch.bfloat16)
    want[:, :, 1:, :-1] = x[:, :cout, :-1, 1:]
    assert torch.equal(got, want)


def test_conv2d_rows_cin64_multi_chunk(ops):
    """64 input channels and a 96-channel output (three 32-wide launches), LeakyReLU(0.1): the ANM n_convs[0] shape class."""
    g = torch.Generator().manual_seed(72)
    x = torch.randn(2, 64, 33, 45, generator=g).to(torch.bfloat16)
    wt = (torch.randn(96, 64, 3, 3, generator=g) * 0.04).to(torch.bfloat16)
    want = F.leaky_relu(F.conv2d(x.float(), wt.float(), None, padding=1), 0.1)
    got = ops.conv2d_rows_multi(x.permute(0, 2, 3, 1).contiguous().cuda(), ops.conv2d_rows_plan(wt.float().cuda()), relu=True, slope=0.1)
    err = (got.permute(0, 3, 1, 2).float().cpu() - want).abs()
    assert err.max().item() < 6e-3 * want.abs().max().item(), err.max().item()


@pytest.mark.parametrize("dil,shape", [(3, (2, 37, 53)), (5, (1, 70, 105)), (3, (2, 280, 420)), (5, (1, 16, 24))])
def test_conv2d_rows_dilated(ops, dil, shape):
    """Dilated 3x3 conv (DPBlock branches, dilation 3 / 5): residue classes of rows as separate dilation-1 launches, column
    dilation as tap offsets; against torch fp32 and an exact delta-weight check across the class / stream boundaries."""
    n, h, w = shape
    g = torch.Generator().manual_seed(73)
    x = torch.randn(n, 32, h, w, generator=g).to(torch.bfloat16)
    wt = (torch.randn(32, 32, 3, 3, generator=g) * 0.06).to(torch.bfloat16)
    bias = torch.randn(32, generator=g) * 0.1
    want = F.conv2d(x.float(), wt.float(), bias, padding=dil, dilation=dil)
    xc = x.permute(0, 2, 3, 1).contiguous().cuda()
    got = ops.conv2d_rows(xc, ops.pack_conv2d_weight(wt.cuda()), 32, None, bias.cuda(), dil=dil)
    err = (got.permute(0, 3, 1, 2).float().cpu() - want).abs()
    assert err.max().item() < 6e-3 * want.abs().max().item(), err.max().item()
    wd = torch.zeros(32, 32, 3, 3)
    wd[torch.arange(32), torch.arange(32), 2, 0] = 1.0                   # y[r, c] = x[r + dil, c - dil]
    got = ops.conv2d_rows(xc, ops.pack_conv2d_weight(wd.cuda()), 32, dil=dil).permute(0, 3, 1, 2).cpu()
    want = torch.zeros(n, 32, h, w, dtype=torch.bfloat16)
    if h > dil and w > dil:
        want[:, :, :h - dil, dil:] = x[:, :, dil:, :w - dil]
    assert torch.equal(got, want)


def test_conv2d_rows_channel_windows_chain(ops):
    """96 -> 32 conv as three 32-channel input windows (x_coff) chained through the residual input: bias + skip enter with the
    first window, the activation leaves with the last (DPBlock conv3 over the concatenated branches)."""
    g = torch.Generator().manual_seed(74)
    x = torch.randn(2, 96, 35, 52, generator=g).to(torch.bfloat16)
    skip = torch.randn(2, 32, 35, 52, generator=g).to(torch.bfloat16)
    wt = (torch.randn(32, 96, 3, 3, generator=g) * 0.04).to(torch.bfloat16)
    bias = torch.randn(32, generator=g) * 0.1
    want = F.leaky_relu(F.conv2d(x.float(), wt.float(), bias, padding=1) + skip.float(), 0.05)
    xc = x.permute(0, 2, 3, 1).contiguous().cuda()
    wins = [ops.pack_conv2d_weight(wt.float().cuda()[:, k:k + 32]) for k in (0, 32, 64)]
    t = ops.conv2d_rows(xc, wins[0], 32, None, bias.cuda(), skip.permute(0, 2, 3, 1).contiguous().cuda(), x_coff=0)
    t = ops.conv2d_rows(xc, wins[1], 32, None, None, t, x_coff=32)
    t = ops.conv2d_rows(xc, wins[2], 32, None, None, t, relu=True, slope=0.05, x_coff=64)
    err = (t.permute(0, 3, 1, 2).float().cpu() - want).abs()
    assert err.max().item() < 1.5e-2 * want.abs().max().item(), err.max().item()      # two extra bf16 roundings of the partial sums


@pytest.mark.parametrize("n_heads,with_normal,with_mask", [(3, True, True), (3, False, True), (1, True, True), (3, True, False)])
def test_fused_losses(ops, n_heads, with_normal, with_mask):
    """dpf_fused_losses (value + gradient in one pass) == the PyTorch statement of the reference's losses (losses.smooth_l1 / cosine,
    themselves pinned against the oracle and the reference's golden loss values on CPU) and its autograd gradients; includes
    |d| > 1 (linear branch), fully masked-out pixels and a zero-length predicted normal (the 1e-6 clamps)."""
    from dualpixelface_b200 import losses
    g = torch.Generator().manual_seed(5 + n_heads)
    b, h, w = 2, 37, 53
    pred = (torch.randn(b, n_heads, h, w, generator=g) * 3.0).cuda().requires_grad_(True)
    disp = torch.randn(b, h, w, generator=g).cuda()
    mask = (torch.rand(b, h, w, generator=g) > 0.3).float().cuda() if with_mask else torch.ones(b, h, w).cuda()
    pn = torch.randn(b, 1, 3, h, w, generator=g)
    pn[0, 0, :, 3, 4] = 0.0                                                   # zero-length prediction: clamp path
    pn = pn.cuda().requires_grad_(True)
    nrm = torch.randn(b, 3, h, w, generator=g).cuda()
    batch = {"disp": disp, "mask": mask, "normal": nrm}
    wts = (1.0, 0.7, 0.5)
    want_l1 = losses.smooth_l1(pred, batch, wts)
    want_lc = losses.cosine(pn, batch) if with_normal else None
    (want_l1 * 1.3 + (want_lc * 0.7 if with_normal else 0.0)).backward()
    gp, gn = pred.grad.clone(), pn.grad.clone() if with_normal else None
    p2, n2 = pred.detach().clone().requires_grad_(True), pn.detach().clone().requires_grad_(True)
    l1, lc = losses.FusedLossFn.apply(p2, n2 if with_normal else None, disp, mask, nrm if with_normal else None, wts)
    (l1 * 1.3 + (lc * 0.7 if with_normal else 0.0)).backward()
    assert abs(float(l1) - float(want_l1)) < 1e-5 * max(1.0, abs(float(want_l1)))
    assert (p2.grad - gp).abs().max().item() < 1e-6 * max(1.0, gp.abs().max().item()) + 1e-9
    if with_normal:
        assert abs(float(lc) - float(want_lc)) < 1e-5
        assert (n2.grad - gn).abs().max().item() < 2e-4 * gn.abs().max().item()
    # determinism: the partial sums are combined in a fixed order
    l1b, _ = losses.FusedLossFn.apply(p2.detach(), n2.detach() if with_normal else None, disp, mask, nrm if with_normal else None, wts)
    assert float(l1b) == float(l1)


@pytest.mark.parametrize("shape", [(2, 5, 37, 50), (1, 8, 28, 60), (1, 1, 14, 30), (2, 3, 16, 33)])
def test_conv3d_head_vs_torch(shape):
    """dpf_conv3d_head_fwd (taps as the GEMM N dimension + shifted sums) == F.conv3d in fp32 on the same bf16 operands; ragged
    sizes, tiles that end exactly at / one past the image border, shift and fp32 residual."""
    import torch.nn.functional as F
    from dualpixelface_b200.layers import KIND_3x3x3, TCConv3d
    b, d, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(b, d, h, w, 32, device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn(1, 32, 3, 3, 3, device="cuda", generator=g) * 0.1).to(torch.bfloat16).float()
    res = torch.randn(b, d, h, w, 1, device="cuda", generator=g)
    layer = TCConv3d(wt, KIND_3x3x3)
    assert layer.head is not None
    want = F.conv3d(x.float().permute(0, 4, 1, 2, 3), wt, padding=1).permute(0, 2, 3, 4, 1)
    got = layer(x, out_f32=True)
    assert got.shape == want.shape
    err = (got - want).abs().max().item()
    print(f"head {shape}: max abs err {err:.2e} (output max {want.abs().max():.2f})")
    assert err < 2e-4 * max(1.0, want.abs().max().item())
    got2 = layer(x, shift=torch.tensor([0.25], device="cuda"), residual=res, out_f32=True)
    assert (got2 - (want + 0.25 + res)).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item())
    assert torch.equal(layer(x, out_f32=True), got)                       # deterministic


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 37, 131), (3, 16, 16)])
def test_stem_conv_vs_torch(shape):
    """dpf_stem_conv_fwd (3x3 stride 2 pad 1, 8-channel padded input, bias + ReLU) == F.conv2d in fp32 on the same bf16 operands,
    incl. odd sizes and tiles that cross the image border."""
    import torch.nn.functional as F
    from dualpixelface_b200 import ops
    n, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.zeros(n, h, w, 8, device="cuda", dtype=torch.bfloat16)
    x[..., :3] = torch.randn(n, h, w, 3, device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn(32, 3, 3, 3, device="cuda", generator=g) * 0.2).to(torch.bfloat16).float()
    bias = torch.randn(32, device="cuda", generator=g) * 0.1
    got = ops.stem_conv(x, ops.pack_stem_weight(wt), bias, relu=True)
    want = F.relu(F.conv2d(x[..., :3].float().permute(0, 3, 1, 2), wt, bias, stride=2, padding=1)).permute(0, 2, 3, 1)
    assert got.shape == want.shape
    err = (got.float() - want).abs().max().item()
    print(f"stem {shape}: max abs err {err:.2e} (output max {want.max():.2f})")
    assert err < 8e-3 * max(1.0, want.max().item())              # bf16 rounding of the output
