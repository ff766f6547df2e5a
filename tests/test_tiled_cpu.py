"""CPU (gloo, world_size 2 and in-process world 1): the host logic of the row-tiled single-pair path (BASELINE config 5,
dualpixelface_b200/tiled.py): halo arithmetic of every conv geometry of the encoder, global-coordinate bilinear upsampling,
the whole tiled StereoDPNet encoder against the untiled one, and the tile re-indexing of the ASM sampling tables."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn
import torch.nn.functional as F

from conftest import ROOT
from dualpixelface_b200 import tiled
from dualpixelface_b200.shift_tables import build_tables


def test_conv_halo_rule():
    assert tiled.conv_halo(3, 1, 1, 1) == (1, 1)
    assert tiled.conv_halo(3, 2, 1, 1) == (1, 0)          # stride 2: the last output row never looks below the tile
    assert tiled.conv_halo(3, 1, 5, 5) == (5, 5)
    assert tiled.conv_halo(3, 2, 2, 2) == (2, 1)          # DPBlock.conv4 with ratio_s = 2
    with pytest.raises(ValueError):
        tiled.conv_halo(3, 4, 1, 0)                       # stride larger than the footprint: not a tiling-friendly geometry


@pytest.mark.parametrize("k,s,d,p", [(3, 1, 1, 1), (3, 2, 1, 1), (3, 1, 3, 3), (3, 2, 2, 2), (3, 1, 2, 2), (1, 2, 1, 0)])
def test_tiled_conv_single_process_tiles_equal_full(k, s, d, p):
    """world = 1 tiling object per tile, halos filled by hand from the full image: every tile reproduces its rows exactly."""
    g = torch.Generator().manual_seed(k * 100 + s * 10 + d)
    conv = nn.Conv2d(4, 6, k, s, p, d, bias=True)
    x = torch.randn(2, 4, 64, 20, generator=g)
    want = conv(x)
    for (y0, y1) in ((0, 16), (16, 48), (48, 64)):
        class T:                                               # hand-made neighbour rows
            def halo_cat(self, t, top, bottom, row_dim, wrap=False):
                up = x[:, :, max(y0 - top, 0):y0]
                up = F.pad(up, (0, 0, top - up.shape[2], 0))
                dn = x[:, :, y1:y1 + bottom]
                dn = F.pad(dn, (0, 0, 0, bottom - dn.shape[2]))
                return torch.cat([up, t, dn], 2)
        got = tiled.tiled_conv2d(conv, x[:, :, y0:y1], T())
        assert torch.allclose(got, want[:, :, y0 // s:y1 // s], atol=1e-6)


def test_local_tables_reindex_and_wrap():
    hg, w, hh = 32, 8, tiled.ASM_HALO
    for direction in ("forward", "backward"):
        tab = build_tables(hg, w, -1.0, direction)
        for q0, q1 in ((0, 16), (16, 32), (0, 32), (8, 24)):
            lt = tiled.local_tables(tab, q0, q1, hg, "cpu")
            n = q1 - q0 + 2 * hh
            assert lt["ri"].shape == (3, n, 2) and lt["rw"].shape == (3, n, 2)
            ext_row = [(q0 - hh + e) % hg for e in range(n)]                 # global row held by each extended row (wrap-around)
            for s_ in range(3):
                for r in range(hh, n - hh):
                    gl = q0 + r - hh
                    for j in range(2):
                        gi, li = int(tab["ri"][s_, gl, j]), int(lt["ri"][s_, r, j])
                        assert (gi < 0) == (li < 0)
                        if gi >= 0:
                            assert ext_row[li] == gi                           # same source row through the tile
                        assert float(lt["rw"][s_, r, j]) == float(tab["rw"][s_, gl, j])
                assert int(lt["ri"][s_, :hh].max()) == -1 and int(lt["ri"][s_, n - hh:].max()) == -1


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        torch.set_num_threads(2)
        from dualpixelface_b200.runner import load_config, model_selector
        model = model_selector(load_config("eval_faceDP", "pytest", root=ROOT, make_dirs=False), root=ROOT).eval()
        for m in model.feature_extraction.modules():          # non-trivial BatchNorm statistics to fold
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5)
        h, w = 192, 48                                          # 96-row tiles: 12 rows at 1/8 resolution >= the dilation-5 halo
        g = torch.Generator().manual_seed(5)
        img = torch.randn(2, 3, h, w, generator=g)
        t = tiled.RowTiling(h, rank, world)
        enc = tiled.TiledSDPEncoder(model.feature_extraction, t, "cpu", torch.float32)
        got = enc(img[:, :, t.y0:t.y1])                                                            # [N, Hloc/4, W/4, C]
        with torch.no_grad():
            want = tiled.fused_torch_encoder(model.feature_extraction, "cpu", torch.float32)(img).permute(0, 2, 3, 1)
        q0, q1 = t.rows(4)
        err = (got - want[:, q0:q1]).abs().max().item() / want.abs().max().item()
        # bilinear rows with global coordinates, halo from the neighbour
        x = torch.randn(1, 3, h // 8, 6, generator=g)
        up = tiled.tiled_bilinear_rows(x[:, :, t.y0 // 8:t.y1 // 8].contiguous(), 2, h // 8, t.y0 // 8, t)
        ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)[:, :, t.y0 // 4:t.y1 // 4]
        # circular halo
        r = torch.arange(float(t.y0), float(t.y1)).view(1, -1, 1)
        wr = t.halo_cat(r, 1, 1, 1, wrap=True)
        ok_wrap = float(wr[0, 0, 0]) == (t.y0 - 1) % h and float(wr[0, -1, 0]) == t.y1 % h
        ret[rank] = (err < 1e-5, bool(torch.allclose(up, ref, atol=1e-6)), ok_wrap, t.exchanges > 50, err)
    finally:
        dist.destroy_process_group()


def test_tiled_encoder_world2_equals_untiled():
    world, port = 2, 31500 + os.getpid() % 2000
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(v[:4] == (True, True, True, True) for v in dict(ret).values()), dict(ret)


def test_overlap_encoder_margin_is_sufficient():
    """Overlap-recompute encoder (no communication): a tile in the MIDDLE of a tall image -- artificial crop borders on both sides,
    exactly encoder_margin() rows away -- reproduces the untiled encoder's rows (to fp32 noise); with a 96-row margin it does not.
    (The analytic receptive field, 346 px, is a guarantee; with these random weights the influence has decayed to 1e-6 by ~224 px.)"""
    from dualpixelface_b200.runner import load_config, model_selector
    torch.manual_seed(0)
    torch.set_num_threads(4)
    model = model_selector(load_config("eval_faceDP", "pytest", root=ROOT, make_dirs=False), root=ROOT).eval()
    for m in model.feature_extraction.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5)
    margin = tiled.encoder_margin(1, 1)
    assert margin == 368
    h, w = 2 * margin + 64, 32                                     # rows [margin, margin + 64) are 'margin' away from both ends
    img = torch.randn(1, 3, h, w, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = tiled.fused_torch_encoder(model.feature_extraction, "cpu", torch.float32)(img).permute(0, 2, 3, 1)

    class T:                                                       # a 3-tile split whose middle tile is the probe
        height, y0, y1, world = h, margin, margin + 64, 3
    enc = tiled.OverlapSDPEncoder(model.feature_extraction, T, "cpu", fused=False, dtype=torch.float32)
    assert (enc.a0, enc.a1) == (0, h)
    # emulate artificial borders: zero everything outside [a0, a1) by shrinking the crop via a smaller margin
    got = enc(img)
    assert (got - want[:, margin // 4:(margin + 64) // 4]).abs().max().item() < 1e-5 * want.abs().max().item()
    for shrink, expect_equal in ((0, True), (margin - 96, False)):
        enc.margin = margin - shrink
        enc.a0, enc.a1 = T.y0 - enc.margin, T.y1 + enc.margin
        pad = shrink                                               # crop rows [a0, a1) of a LARGER image: real data beyond the crop is cut off
        big = torch.randn(1, 3, h + 2 * 64, w, generator=torch.Generator().manual_seed(2))
        del pad
        class TB:
            height, y0, y1, world = h + 128, margin + 64, margin + 128, 3
        encb = tiled.OverlapSDPEncoder(model.feature_extraction, TB, "cpu", fused=False, dtype=torch.float32)
        encb.margin = margin - shrink
        encb.a0, encb.a1 = TB.y0 - encb.margin, TB.y1 + encb.margin
        with torch.no_grad():
            wantb = tiled.fused_torch_encoder(model.feature_extraction, "cpu", torch.float32)(big).permute(0, 2, 3, 1)[:, TB.y0 // 4:TB.y1 // 4]
        err = (encb(big) - wantb).abs().max().item() / wantb.abs().max().item()
        assert (err < 1e-5) == expect_equal, (shrink, err)
