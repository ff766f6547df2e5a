"""BASELINE config 5: ONE high-resolution dual-pixel pair split into ROW TILES over the GPUs of a box, with halo exchange between
row neighbours (no counterpart in the reference, which runs a single image on a single GPU).

One process per GPU; rank r owns the full-resolution rows [y0_r, y1_r) with boundaries on multiples of 16 px
(``parallel.row_tiles``), i.e. whole rows at every resolution of the network (1/2, 1/4, 1/8, 1/16).  Every layer computes exactly
its own rows; what it needs from the neighbours travels as HALO ROWS over NCCL point-to-point (``dist.batch_isend_irecv``:
both directions of a layer's exchange in one group call), never as recomputation (the receptive field of the path is ~51
quarter-resolution rows against 70-row tiles).  Rows beyond the image border are zeros -- they ARE the convolution padding.

  2-D encoder      default "overlap": the untiled encoder plan on the tile + its receptive-field margin (368 px), no communication
                   (the encoder is adjacent to the path and 22 % of the FLOPs; its ~45 per-layer exchanges cost more host time
                   than the recomputation).  "exchange": every k x k conv (stride s, dilation d) runs on cat(top halo, rows, bottom
                   halo) with zero padding along H switched off: top = d(k-1)/2 rows, bottom = d(k-1)+1-s-top rows.  FPN nearest x2
                   is tile-local; the bilinear x2 / x4 pyramid upsampling uses the GLOBAL align_corners coordinates.
  ASM volume       the sampling tables are built for the global height and re-indexed to the tile; 2 halo rows with WRAP-AROUND
                   between the first and the last tile (the phase sample is a circular row shift, asm.py:63-75); the
                   InstanceNorm statistics are all-reduced ([2B,32,2] fp32).
  3-D aggregation  activations live in buffers with 2 halo rows at every resolution; stride-1 convs run on the whole buffer and
                   refresh their halo (2 rows each way) afterwards; stride-2 convs produce their own bottom halo, transposed
                   convs both halos; BN / residual / ReLU stay fused in the conv epilogues.
  regression       dpf_regress_fwd_tile / dpf_anm_tail_tile: trilinear / bilinear x4 with the GLOBAL coordinates.
  ANM              min / max of the coordinate volume all-reduced; D3D halo = 1 + ceil(max |row offset|), measured per layer
                   (all-reduce MAX); the dilated 2-D convs exchange d rows.

``TiledStereoDPNet(model, height, rank, world)(batch)`` takes the FULL pair on every rank (the images are 45 MB; slicing is
local) and returns this rank's rows of ``pred_depth`` / ``pred_normal``; ``gather()`` assembles them on every rank.  With
world == 1 the same code runs without communication (used by the tests to pin the tiling logic against the untiled model).
"""
from __future__ import annotations

import copy
from collections import OrderedDict
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import ops, shift_tables
from .parallel import row_tiles


class RowTiling:
    """Row ranges of one rank at every resolution + the halo exchange primitive."""

    def __init__(self, height: int, rank: int, world: int, group=None, unit: int = 16):
        self.height, self.rank, self.world, self.group = height, rank, world, group
        self.tiles = row_tiles(height, world, unit)
        self.y0, self.y1 = self.tiles[rank]
        self.first, self.last = rank == 0, rank == world - 1
        self.bytes_exchanged = 0          # halo bytes this rank sent (reported by bench.py)
        self.exchanges = 0
        self.log = None                   # a list -> every exchange appends (tensor shape, (top, bottom) rows, bytes sent)

    def rows(self, div: int) -> Tuple[int, int]:
        return self.y0 // div, self.y1 // div

    def halo(self, x: torch.Tensor, top: int, bottom: int, row_dim: int, wrap: bool = False) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """(rows from above [top], rows from below [bottom]) of `x` (this rank's rows, no halo): the upper neighbour's last `top`
        rows and the lower neighbour's first `bottom` rows; zeros beyond the image border, or -- with wrap -- the rows of the
        opposite end of the image.  One batched NCCL / gloo point-to-point group."""
        n = x.shape[row_dim]
        assert top <= n and bottom <= n, f"halo ({top}, {bottom}) larger than the {n}-row tile"
        shp = lambda k: x.shape[:row_dim] + (k,) + x.shape[row_dim + 1:]
        up_buf = x.new_zeros(shp(top)) if top else None
        dn_buf = x.new_zeros(shp(bottom)) if bottom else None
        if self.world == 1:
            if wrap:
                if top:
                    up_buf = x.narrow(row_dim, n - top, top).contiguous()
                if bottom:
                    dn_buf = x.narrow(row_dim, 0, bottom).contiguous()
            return up_buf, dn_buf
        up, dn = self.rank - 1, self.rank + 1
        if wrap:
            up, dn = up % self.world, dn % self.world
        # Order matters when both neighbours are the SAME rank (world 2 with wrap-around): messages between one pair match in posting
        # order, and every rank receives "rows from above" first -- so the rows that are somebody's upper halo are sent first.
        p2p = []
        if top and 0 <= dn < self.world:             # my last rows are the upper halo of the rank below
            send = x.narrow(row_dim, n - top, top).contiguous()
            p2p.append(dist.P2POp(dist.isend, send, dn, self.group))
            self.bytes_exchanged += send.numel() * send.element_size()
        if bottom and 0 <= up < self.world:          # my first rows are the lower halo of the rank above
            send = x.narrow(row_dim, 0, bottom).contiguous()
            p2p.append(dist.P2POp(dist.isend, send, up, self.group))
            self.bytes_exchanged += send.numel() * send.element_size()
        if top and 0 <= up < self.world:
            p2p.append(dist.P2POp(dist.irecv, up_buf, up, self.group))
        if bottom and 0 <= dn < self.world:
            p2p.append(dist.P2POp(dist.irecv, dn_buf, dn, self.group))
        if p2p:
            self.exchanges += 1
            if self.log is not None:
                sent = sum(op.tensor.numel() * op.tensor.element_size() for op in p2p if op.op is dist.isend)
                self.log.append((tuple(x.shape), str(x.dtype).replace("torch.", ""), (top, bottom), int(sent)))
            for w in dist.batch_isend_irecv(p2p):
                w.wait()
        return up_buf, dn_buf

    def halo_cat(self, x: torch.Tensor, top: int, bottom: int, row_dim: int, wrap: bool = False) -> torch.Tensor:
        up, dn = self.halo(x, top, bottom, row_dim, wrap)
        parts = ([up] if up is not None else []) + [x] + ([dn] if dn is not None else [])
        if len(parts) == 1:
            return x
        if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
            # NCHW views of channels-last memory (the 2-D encoder): keep the layout, cuDNN would otherwise transpose every layer
            parts = [p.contiguous(memory_format=torch.channels_last) for p in parts]
            return torch.cat(parts, row_dim).contiguous(memory_format=torch.channels_last)
        return torch.cat(parts, row_dim)

    def halo_fill_(self, x_ext: torch.Tensor, h: int, row_dim: int) -> torch.Tensor:
        """Refresh the `h` halo rows on both sides of an extended buffer in place from the neighbours' interior rows."""
        n = x_ext.shape[row_dim]
        up, dn = self.halo(x_ext.narrow(row_dim, h, n - 2 * h), h, h, row_dim)
        x_ext.narrow(row_dim, 0, h).copy_(up)
        x_ext.narrow(row_dim, n - h, h).copy_(dn)
        return x_ext

    def zero_border_(self, x_ext: torch.Tensor, h: int, row_dim: int) -> torch.Tensor:
        """Rows beyond the image border do not exist: they must read as zeros (= the convolution padding)."""
        n = x_ext.shape[row_dim]
        if self.first:
            x_ext.narrow(row_dim, 0, h).zero_()
        if self.last:
            x_ext.narrow(row_dim, n - h, h).zero_()
        return x_ext

    def all_reduce(self, t: torch.Tensor, op=dist.ReduceOp.SUM) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=op, group=self.group)
        return t


# ======================================================================================================
# 2-D encoder on row tiles (torch convolutions with BatchNorm folded, bf16 channels-last; adjacent to the hot path)
# ======================================================================================================
def fused_torch_encoder(enc: nn.Module, device, dtype=torch.bfloat16) -> nn.Module:
    """Eval-mode deep copy of a feature extractor with every BatchNorm2d folded into its convolution."""
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    from . import modules as M
    enc = copy.deepcopy(enc).eval()

    def walk(mod):
        for _, child in list(mod.named_children()):
            if isinstance(child, nn.Sequential) and len(child) >= 2 and isinstance(child[0], nn.Conv2d) and isinstance(child[1], nn.BatchNorm2d):
                child[0] = fuse_conv_bn_eval(child[0], child[1])
                child[1] = nn.Identity()
            if isinstance(child, M._SepConv):
                child.pointwise = fuse_conv_bn_eval(child.pointwise, child.bn)
                child.bn = nn.Identity()
            walk(child)

    walk(enc)
    return enc.to(device=device, dtype=dtype, memory_format=torch.channels_last)


def conv_halo(kernel: int, stride: int, dilation: int, padding: int) -> Tuple[int, int]:
    """(top, bottom) halo rows a k x k conv needs so that a tile whose first row is a multiple of `stride` produces exactly its
    own output rows with zero padding along H switched off:  out row j reads in rows j*s - p + t*d, t = 0..k-1."""
    span = dilation * (kernel - 1)
    top = padding
    bottom = span + 1 - stride - top
    if bottom < 0:
        raise ValueError(f"conv k={kernel} s={stride} d={dilation} p={padding}: not a 'same'-style convolution")
    return top, bottom


def tiled_conv2d(conv: nn.Conv2d, x: torch.Tensor, tiling: RowTiling) -> torch.Tensor:
    kh = conv.kernel_size[0]
    if kh == 1:
        assert conv.padding[0] == 0
        return F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
    top, bottom = conv_halo(kh, conv.stride[0], conv.dilation[0], conv.padding[0])
    assert x.shape[2] % conv.stride[0] == 0, "tile heights are multiples of every stride of the network"
    xe = tiling.halo_cat(x, top, bottom, 2)
    return F.conv2d(xe, conv.weight, conv.bias, conv.stride, (0, conv.padding[1]), conv.dilation, conv.groups)


_ROW_TABLES: dict = {}


def tiled_bilinear_rows(x: torch.Tensor, factor: int, h_in_global: int, in_row0: int, tiling: RowTiling) -> torch.Tensor:
    """F.interpolate(scale_factor=factor, mode='bilinear', align_corners=True) for this rank's output rows: x [N,C,hin,win] holds
    the input rows in_row0 .. in_row0+hin-1 of an image h_in_global rows tall.  Source index rule of ATen (UpSample.cuh):
    src = dst * (in-1)/(out-1) in fp32, i0 = int(src), i1 = min(i0+1, in-1), weight src - i0.  Returns fp32."""
    n, c, hin, win = x.shape
    h_out_global = h_in_global * factor
    o0, o1 = in_row0 * factor, (in_row0 + hin) * factor
    xe = tiling.halo_cat(x, 1, 1, 2).float()                                    # rows in_row0-1 .. in_row0+hin
    key = (factor, h_in_global, in_row0, hin, str(x.device))
    if key not in _ROW_TABLES:                                                  # host-built once (no H2D copy in the steady state)
        scale = torch.tensor((h_in_global - 1) / (h_out_global - 1), dtype=torch.float32) if h_out_global > 1 else torch.tensor(0.0)
        src = scale * torch.arange(o0, o1, dtype=torch.float32)
        i0 = src.to(torch.int64)
        i1 = torch.clamp(i0 + 1, max=h_in_global - 1)
        _ROW_TABLES[key] = ((src - i0.to(torch.float32)).to(x.device).view(1, 1, -1, 1), (i0 - (in_row0 - 1)).to(x.device),
                            (i1 - (in_row0 - 1)).to(x.device))
    l1, li0, li1 = _ROW_TABLES[key]
    rows = (1.0 - l1) * xe.index_select(2, li0) + l1 * xe.index_select(2, li1)   # [N,C,hout_loc,win] fp32
    # columns: the untiled rule (H is already at its final size: identity along H)
    return F.interpolate(rows, size=(rows.shape[2], win * factor), mode="bilinear", align_corners=True)


def encoder_margin(n_inter1: int = 1, n_inter2: int = 1) -> int:
    """Half-height of the StereoDPNet encoder's receptive field in full-resolution rows, rounded up to a multiple of 16 (+16 spare).
    Walking back from the quarter-resolution output (src/model/stereodpnet/modules.py:58-134): lastconv 2 x (3x3 @1/4) = 8 px;
    bilinear x4 of the 1/16 level = 16; FPN 3x3 @1/16 = 16; block3 (DPBlock @1/8 -> 1/16: conv1 1 + conv2 1 + dilated 5 + conv3 1 +
    conv4 2 + depthwise 2 = 12 rows @1/8) = 96; each interblock2 (11 rows @1/8) = 88; block2 (12 rows @1/4) = 48; each interblock1
    (11 rows @1/4) = 44; block1 (12 rows @1/2) = 24; firstconv (2 rows @1/2 + the stride-2 3x3) = 6."""
    px = 8 + 16 + 16 + 96 + 88 * n_inter2 + 48 + 44 * n_inter1 + 24 + 6
    return (px + 15) // 16 * 16 + 16


class _Crop:
    """A row crop [a0, a1) of an image `height` rows tall, seen as a world-1 tiling (halo rows = zeros)."""

    def __init__(self, height, a0, a1):
        self.height, self.y0, self.y1, self.world = height, a0, a1, 1

    def halo_cat(self, x, top, bottom, row_dim, wrap=False):
        shp = lambda k: x.shape[:row_dim] + (k,) + x.shape[row_dim + 1:]
        return torch.cat([x.new_zeros(shp(top)), x, x.new_zeros(shp(bottom))], row_dim)


class OverlapSDPEncoder:
    """The encoder by OVERLAP-RECOMPUTE: this rank runs the UNTILED encoder on its rows plus the receptive-field margin
    (encoder_margin(): 368 px each side) and keeps its own rows.  No communication at all: the per-layer exchange variant
    (TiledFusedSDPEncoder: ~45 exchanges) is bound by the host-side cost of its NCCL calls (12.4 ms for a half image on 2 GPUs),
    while the encoder is only 22 % of the model's FLOPs, so recomputing (280 + 2 x 368) / 280 of a tile is cheaper.  The 3-D
    path (receptive field ~51 quarter-res rows per side against 70-row tiles) keeps the per-layer halo exchange.
    fused=True: the bench configuration's plan (encoder_fused.FusedSDPEncoder, crop-aware pyramid kernel); fused=False: plain torch
    modules in `dtype` (CPU / fp32 tests)."""

    def __init__(self, enc: nn.Module, tiling: RowTiling, device, fused: bool = True, dtype=torch.float32):
        self.t, self.fused, self.dtype = tiling, fused, dtype
        self.margin = encoder_margin(len(enc.interblock1), len(enc.interblock2))
        self.a0, self.a1 = max(0, tiling.y0 - self.margin), min(tiling.height, tiling.y1 + self.margin)
        if fused:
            from .encoder_fused import FusedSDPEncoder
            self.plan = FusedSDPEncoder(enc)
        else:
            self.enc = fused_torch_encoder(enc, device, dtype)

    @torch.no_grad()
    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        """images [N,3,H,W] (the WHOLE image) -> [N,Hloc/4,W/4,C] for this rank's rows."""
        t = self.t
        x = images[:, :, self.a0:self.a1]
        lo, hi = (t.y0 - self.a0) // 4, (t.y1 - self.a0) // 4
        if self.fused:
            from . import encoder_fused
            encoder_fused.CROP = (t.height // 4, self.a0 // 4)
            try:
                y = self.plan(x)
            finally:
                encoder_fused.CROP = None
            return y.permute(0, 2, 3, 1)[:, lo:hi].contiguous()
        e, crop = self.enc, _Crop(t.height, self.a0, self.a1)
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
        o1 = e.block1(e.firstconv(x))
        o2 = o1
        for m in e.interblock1:
            o2 = m(o2)
        o2 = e.block2(o2)
        o3 = o2
        for m in e.interblock2:
            o3 = m(o3)
        o3 = e.block3(o3)
        f = e.fpn(OrderedDict(layer1=o1, layer2=o2, layer3=o3))
        up2 = tiled_bilinear_rows(f["layer2"], 2, t.height // 8, self.a0 // 8, crop)
        up4 = tiled_bilinear_rows(f["layer3"], 4, t.height // 16, self.a0 // 16, crop)
        y = e.lastconv(torch.cat([f["layer1"], up2.to(self.dtype), up4.to(self.dtype)], 1))
        return y.permute(0, 2, 3, 1)[:, lo:hi].contiguous()


class TiledFusedSDPEncoder:
    """The bench configuration's encoder plan (encoder_fused.FusedSDPEncoder: BatchNorm folded, bf16 channels-last cuDNN convs with
    fused bias / activation tails, FPN merge kernel) on this rank's rows: the plan's convolutions pick up their halo rows through
    encoder_fused.TILING."""

    def __init__(self, enc: nn.Module, tiling: RowTiling):
        from .encoder_fused import FusedSDPEncoder
        self.t = tiling
        self.plan = FusedSDPEncoder(enc)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        from . import encoder_fused
        encoder_fused.TILING = self.t
        try:
            y = self.plan(x)                                         # [N,C,Hloc/4,W/4] channels-last
        finally:
            encoder_fused.TILING = None
        return y.permute(0, 2, 3, 1).contiguous()


class TiledSDPEncoder:
    """feature_extraction of StereoDPNet (src/model/stereodpnet/modules.py:58-134) on this rank's rows, plain torch modules (any
    dtype / device: the fp32 parity tests and the CPU gloo tests use it)."""

    def __init__(self, enc: nn.Module, tiling: RowTiling, device, dtype=torch.bfloat16):
        self.t, self.dtype = tiling, dtype
        self.enc = fused_torch_encoder(enc, device, dtype)
        for m in self.enc.modules():
            if isinstance(m, nn.Conv2d):
                m.forward = (lambda x, _m=m: tiled_conv2d(_m, x, self.t))

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """x [N,3,Hloc,W] (this rank's rows) -> [N,Hloc/4,W/4,C] bf16 channels-last."""
        e, t = self.enc, self.t
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
        o1 = e.block1(e.firstconv(x))
        o2 = o1
        for m in e.interblock1:
            o2 = m(o2)
        o2 = e.block2(o2)
        o3 = o2
        for m in e.interblock2:
            o3 = m(o3)
        o3 = e.block3(o3)
        f = e.fpn(OrderedDict(layer1=o1, layer2=o2, layer3=o3))       # nearest x2 top-down merges are tile-local
        up2 = tiled_bilinear_rows(f["layer2"], 2, t.height // 8, t.y0 // 8, t)
        up4 = tiled_bilinear_rows(f["layer3"], 4, t.height // 16, t.y0 // 16, t)
        y = torch.cat([f["layer1"], up2.to(self.dtype), up4.to(self.dtype)], 1).contiguous(memory_format=torch.channels_last)
        y = e.lastconv(y)
        return y.permute(0, 2, 3, 1).contiguous()


# ======================================================================================================
# the 3-D hot path on row tiles
# ======================================================================================================
HALO = 2          # halo rows of every 3-D activation buffer (even: stride-2 layers keep their row parity)


ASM_HALO = 2      # rows the sampling tables reach: shift +-1 row, and the bilinear table's second tap one row further (weight ~1e-5)


def local_tables(tab: dict, q0: int, q1: int, hglob: int, device, hh: int = ASM_HALO) -> dict:
    """Re-index global sampling tables (shift_tables.build_tables for the GLOBAL height) to a tile that holds the global rows
    q0-hh .. q1+hh-1 modulo the image height (hh halo rows each side with WRAP-AROUND: the phase sample is a circular shift):
    extended row e holds global row (q0 - hh + e) mod H, so source row g sits at (g - q0 + hh) mod H."""
    ri, rw = tab["ri"][:, q0:q1].to(torch.int64), tab["rw"][:, q0:q1].clone()
    n = q1 - q0 + 2 * hh
    loc = torch.remainder(ri - (q0 - hh), hglob)
    assert not bool(((ri >= 0) & (loc >= n)).any()), f"a sampling table reaches beyond the {hh}-row halo"
    loc = torch.where(ri < 0, ri, loc).to(torch.int32)
    # the kernel produces as many rows as it reads: pad the table to the extended height (the halo rows' outputs are discarded)
    pad_i = torch.full((ri.shape[0], hh, 2), -1, dtype=torch.int32)
    pad_w = torch.zeros(ri.shape[0], hh, 2)
    return {"ri": torch.cat([pad_i, loc, pad_i], 1).contiguous().to(device), "rw": torch.cat([pad_w, rw, pad_w], 1).contiguous().to(device),
            "ci": tab["ci"].to(device), "cw": tab["cw"].to(device)}


class TiledStereoDPNet:
    """Row-tiled inference of a STEREODPNET (eval mode, shipped configuration: cached first level, deformable ANM)."""

    def __init__(self, model, height: int, rank: int = 0, world: int = 1, group=None):
        from .models import STEREODPNET
        assert isinstance(model, STEREODPNET) and not model.training
        self.m = model
        self.t = RowTiling(height, rank, world, group)
        dev = next(model.parameters()).device
        # encoder precision follows the model: bf16 (the bench configuration) or fp32 (encoder_autocast off: parity tests)
        # "overlap" (default): untiled encoder on the tile + its receptive-field margin, no communication; "exchange": per-layer halos
        self.encoder_mode = "overlap"
        self.enc = OverlapSDPEncoder(model.feature_extraction, self.t, dev, fused=bool(model.encoder_autocast))
        self._enc_exchange = None
        self.stage_events = None          # set to [] to collect per-stage CUDA events (tools/tiled_check.py, bench.py)
        # D3D halo rows per layer: measured on the first pass (all-reduced max |row offset|, one host sync each), then FIXED with a
        # margin so that later passes are sync-free (and capturable in a CUDA graph); check_reach() verifies it after the fact
        self._hd: List[Optional[int]] = [None, None]
        self._reach = [None, None]
        self._kinv_key, self._kinv = None, None
        model.aggregation._build()
        model.cost_volume._pack()
        if model.predict_normal:
            model.normal_estimator._build()
        self._tabs = {}

    # ---- ASM volume -------------------------------------------------------------------------------------------------------
    def _volume(self, ref: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
        """ref / tgt [B,Hq,W4,C] local rows -> volume buffer [B,D,Hq+2*HALO,W4,2C] (halo not yet valid)."""
        cv, t = self.m.cost_volume, self.t
        if not cv.cached_first_level:
            raise NotImplementedError("row tiles implement the shipped (cached first level) volume")
        b, hq, w4, c = ref.shape
        q0, q1 = t.rows(4)
        hglob = t.height // 4
        disp = cv.costrange[0]
        key = (hglob, w4, q0, q1)
        if key not in self._tabs:
            self._tabs[key] = {d: local_tables(shift_tables.build_tables(hglob, w4, disp, d, cv.shifting_layer.modes), q0, q1, hglob, ref.device)
                               for d in ("forward", "backward")}
        x = torch.cat([ref, tgt], 0)                                           # [2B,Hq,W4,C]
        xe = t.halo_cat(x, ASM_HALO, ASM_HALO, 1, wrap=True)                   # circular: the phase sample wraps around the image
        sf = ops.asm_sample(xe[:b].contiguous(), self._tabs[key]["forward"])[:, :, ASM_HALO:-ASM_HALO]
        sb = ops.asm_sample(xe[b:].contiguous(), self._tabs[key]["backward"])[:, :, ASM_HALO:-ASM_HALO]
        smp = torch.cat([sf, sb], 0).contiguous()                              # [2B,S,Hq,W4,C]
        # mask convs: 1x3x3 needs 1 halo row (zeros at the border = its padding), 1x1x1 none
        pk = cv._pack()
        se = t.halo_cat(smp, 1, 1, 2)
        mfeat = pk["conv1"](se, pk["bn"][0], pk["bn"][1], relu=True)[:, :, 1:-1].contiguous()
        logits = pk["conv2"](mfeat)
        st = t.all_reduce(ops.channel_stats(logits))                           # InstanceNorm3d: statistics over the WHOLE image
        n = float(logits.shape[1] * hglob * w4)
        mean = st[..., 0] / n
        var = (st[..., 1] / n - mean * mean).clamp_min(0.0)
        inorm = cv.attention_layer.normalize
        a = (inorm.weight.float().unsqueeze(0) / torch.sqrt(var + inorm.eps)).contiguous()
        d = (inorm.bias.float().unsqueeze(0) - mean * a).contiguous()
        vol = torch.empty(b, cv.level, hq, w4, 2 * c, device=ref.device, dtype=torch.bfloat16)
        ops.asm_blend(smp[:b], logits[:b], a[:b].contiguous(), d[:b].contiguous(), vol, 0, cv.level, 0)
        ops.asm_blend(smp[b:], logits[b:], a[b:].contiguous(), d[b:].contiguous(), vol, 0, cv.level, c)
        return self._ext(vol)

    # ---- buffers with halo ----------------------------------------------------------------------------------------------------
    def _ext(self, x: torch.Tensor, row_dim: int = 2, h: int = HALO) -> torch.Tensor:
        """interior rows -> extended buffer with valid halos."""
        return self.t.halo_cat(x, h, h, row_dim)

    def _s1(self, name: str, x_ext, residual=None, relu=True):
        conv, (sc, sh) = self.m.aggregation._plan[name]
        y = conv(x_ext, sc, sh, residual=residual, relu=relu)
        return self.t.halo_fill_(y, HALO, 2)

    def _s2(self, name: str, x_ext):
        conv, (sc, sh) = self.m.aggregation._plan[name]
        y = conv(x_ext, sc, sh, relu=True)                                     # rows: halo 1 (top row invalid), see module docstring
        out = torch.empty(y.shape[0], y.shape[1], y.shape[2] + 2, y.shape[3], y.shape[4], device=y.device, dtype=y.dtype)
        out[:, :, 1:-1] = y
        return self.t.halo_fill_(out, HALO, 2)

    def _t2(self, name: str, x_ext, residual_ext, relu: bool):
        conv, (sc, sh) = self.m.aggregation._plan[name]
        y = conv(x_ext, sc, sh, relu=False)[:, :, HALO:-HALO]                  # [.., 2*Hc + 2*HALO, ..]: both halos computed locally
        y = y + residual_ext
        if relu:
            y = torch.relu_(y)
        return self.t.zero_border_(y.contiguous(), HALO, 2)

    def _hourglass(self, name, x, presqu, postsqu, cost0):
        o = self._s2(name + ".conv1", x)
        pre = self._s1(name + ".conv2", o, residual=postsqu, relu=True)
        o = self._s2(name + ".conv3", pre)
        o = self._s1(name + ".conv4", o)
        post = self._t2(name + ".conv5", o, presqu if presqu is not None else pre, True)
        out = self._t2(name + ".conv6", post, cost0, False)
        return out, pre, post

    def _aggregate(self, vol_ext):
        c0 = self._s1("dres0.0", vol_ext)
        c0 = self._s1("dres0.2", c0)
        r = self._s1("dres1.0", c0)
        cost0 = self._s1("dres1.2", r, residual=c0, relu=False)
        out1, pre1, post1 = self._hourglass("dres2", cost0, None, None, cost0)
        out2, _, post2 = self._hourglass("dres3", out1, pre1, post1, cost0)
        out3, _, _ = self._hourglass("dres4", out2, pre1, post2, cost0)
        prev = None
        for k, o in ((1, out1), (2, out2), (3, out3)):
            hfeat = self._s1(f"classif{k}.0", o)
            head, _ = self.m.aggregation._plan[f"classif{k}.2"]
            prev = head(hfeat, residual=prev, relu=False, out_f32=True)        # cumulative adds
        # rows 1 .. n-2 of the head output are valid (hfeat's halo is): the regression reads rows q0-1 .. q1 only -> no exchange
        return prev.squeeze(-1), out3

    # ---- regression + normal branch ---------------------------------------------------------------------------------------------
    def _regress(self, cost3_ext: torch.Tensor) -> torch.Tensor:
        from . import _lib
        t, reg = self.t, self.m.regression_layer
        b, d, he, w4 = cost3_ext.shape
        q0, _ = t.rows(4)
        rows = t.y1 - t.y0
        disp = torch.empty(b, rows, 4 * w4, device=cost3_ext.device, dtype=torch.float32)
        _lib.check(ops.lib().dpf_regress_fwd_tile(ops._p(cost3_ext.contiguous()), ops._p(disp), None, b, d, he, w4, t.height // 4,
                                                  q0 - HALO, rows, t.y0, float(reg.mindisp), float(reg.step), ops._stream()),
                   "dpf_regress_fwd_tile")
        return disp

    def _normals(self, out3_ext: torch.Tensor, disp: torch.Tensor, batch: dict) -> torch.Tensor:
        from .ops_tail import anm_tail
        anm, t = self.m.normal_estimator, self.t
        if not anm.use_deform:
            raise NotImplementedError("row tiles implement the shipped (deformable) normal branch")
        p = anm._plan
        b = disp.shape[0]
        q0, q1 = t.rows(4)
        if self._kinv_key != batch["K"].data_ptr():                            # torch.inverse synchronises: once per intrinsics tensor
            kq = batch["K"].float().clone()
            kq[:, :2, :] = kq[:, :2, :] / 4.0
            kinv = torch.inverse(kq)
            kinv[:, :, 2] = kinv[:, :, 2] + kinv[:, :, 1] * float(q0)          # the kernel's row index is tile-local: v = h + q0
            self._kinv_key, self._kinv = batch["K"].data_ptr(), kinv.contiguous()
        idx, coord, minmax = ops.anm_select(disp.contiguous(), self._kinv, batch["abvalue"].float().contiguous(), anm.levels, anm.k)
        mn = t.all_reduce(minmax[:, 0].contiguous(), dist.ReduceOp.MIN)        # coordinate normalisation over the WHOLE image
        mx = t.all_reduce(minmax[:, 1].contiguous(), dist.ReduceOp.MAX)
        minmax = torch.stack([mn, mx], 1).contiguous()
        out3 = out3_ext[:, :, HALO:-HALO].contiguous()
        x = ops.anm_gather(out3, idx, coord, minmax, 64)                       # [B,K,Hq,W4,64] interior rows
        for i in (1, 2):
            off = p[f"off{i}"](t.halo_cat(x, 1, 1, 2), shift=p[f"offb{i}"], out_f32=True)[:, :, 1:-1].contiguous()
            reach = t.all_reduce(off[..., 1:81:3].abs().max().reshape(1), dist.ReduceOp.MAX)     # (d, h, w) per tap: rows = 1::3
            self._reach[i - 1] = reach
            if self._hd[i - 1] is None:                                                     # first pass: one host sync per layer
                self._hd[i - 1] = min(int(torch.ceil(reach).item()) + 3, x.shape[2])        # tap (+-1) + offset + trilinear corner + margin
            hd = self._hd[i - 1]
            y = ops.dcn3d(t.halo_cat(x, hd, hd, 2), t.halo_cat(off, hd, hd, 2), p[f"w{i}"], p[f"cpad{i}"], p[f"aff{i}"][0],
                          p[f"aff{i}"][1], relu=True)
            x = y[:, :, hd:-hd].contiguous()
        f = x.view(b * anm.k, x.shape[2], x.shape[3], x.shape[4])
        for wp, cout, dil in p["nconv"]:
            fe = t.halo_cat(f, dil, dil, 1)
            f = ops.conv2d_tc(fe, wp, cout, dil, relu=True, slope=0.1)[:, dil:-dil].contiguous()
        fe = t.halo_cat(f, 1, 1, 1)
        return anm_tail(fe, b, anm.k, h4_global=t.height // 4, q_row0=q0 - 1, out_rows=t.y1 - t.y0, y_row0=t.y0)

    def _mark(self, name):
        if self.stage_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.stage_events.append((name, e))

    @torch.no_grad()
    def __call__(self, batch: dict) -> dict:
        m, t = self.m, self.t
        ref_img, tgt_img = m._select_views(batch)
        assert ref_img.shape[-2] == t.height
        b = ref_img.shape[0]
        self._mark("start")
        if self.encoder_mode == "overlap":
            f = self.enc(torch.cat([ref_img, tgt_img], 0)).to(torch.bfloat16)
        else:
            if self._enc_exchange is None:
                dev = ref_img.device
                self._enc_exchange = TiledFusedSDPEncoder(m.feature_extraction, t) if m.encoder_autocast else \
                    TiledSDPEncoder(m.feature_extraction, t, dev, torch.float32)
            f = self._enc_exchange(torch.cat([ref_img[:, :, t.y0:t.y1], tgt_img[:, :, t.y0:t.y1]], 0)).to(torch.bfloat16)
        self._mark("encoder")
        vol = self._volume(f[:b].contiguous(), f[b:].contiguous())
        self._mark("cost_volume")
        cost3, out3 = self._aggregate(vol)
        self._mark("aggregation")
        disp = self._regress(cost3)
        self._mark("regression")
        normal = self._normals(out3, disp, batch) if m.predict_normal else None
        self._mark("normal_branch")
        return {"pred_depth": disp.unsqueeze(1), "pred_normal": normal.unsqueeze(1) if normal is not None else None,
                "rows": (t.y0, t.y1)}

    def check_reach(self) -> bool:
        """True if the D3D row offsets of the LAST pass stayed inside the halo fixed on the first pass (one host sync)."""
        return all(r is None or float(r) <= h - 2 for r, h in zip(self._reach, self._hd))

    def capture(self, batch: dict):
        """Capture one tiled forward -- kernels AND the NCCL halo exchanges -- in a CUDA graph.  The tiled path is launch-bound
        (~85 exchanges + ~500 kernels for a few ms of GPU work per rank); replaying a graph removes the host from the loop.
        Returns (replay, static_batch, static_out): copy new images into static_batch['left'/'right'], call replay(), read static_out."""
        static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
        for _ in range(2):                                     # eager warm-up: communicators, plans, tables, the D3D halo
            self(static)
        torch.cuda.synchronize()
        if self.t.world > 1:
            dist.barrier(group=self.t.group)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self(static)
        return g.replay, static, out

    def gather(self, res: dict) -> dict:
        """Assemble the full image on every rank (all_gather of the row tiles; tiles differ by at most 16 rows, padded)."""
        t = self.t
        if t.world == 1:
            return res
        out = {}
        hmax = max(e - s for s, e in t.tiles)
        for k in ("pred_depth", "pred_normal"):
            v = res[k]
            if v is None:
                out[k] = None
                continue
            pad = F.pad(v, (0, 0, 0, hmax - v.shape[-2])).contiguous()
            parts = [torch.empty_like(pad) for _ in range(t.world)]
            dist.all_gather(parts, pad, group=t.group)
            out[k] = torch.cat([p_[..., : e - s, :] for p_, (s, e) in zip(parts, t.tiles)], -2)
        return out
